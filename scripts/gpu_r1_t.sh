#!/bin/bash
# session 2, call 5: 8-warp flash CTAs + cheaper tile loads + SFU exp2; dense VAE attention; GEMM tile sweep
mkdir -p gpurun_out
echo "== kernel tests"; timeout 1200 python -m pytest tests/test_gpu_kernels.py -q --timeout 600 2>&1 | tail -5 | cut -c1-250
echo "== kernel bench"; timeout 300 python scripts/kernel_bench.py --only self_attn 2>&1 | cut -c1-200
echo "== full-size parity"; timeout 1200 python -m pytest tests/test_gpu_pipeline.py -q -s -k "full" --timeout 1000 2>&1 | grep -E "full-size|passed|failed|Error|assert" | cut -c1-300 | tail -20
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/t_bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','e2e','early_exit_images_per_s_1gpu','gpu_launches')})
PY
echo "== gemm sweep"; timeout 900 python scripts/gemm_sweep.py --emit gpurun_out/skp_gemm_tuned.inc 2>&1 | cut -c1-260
