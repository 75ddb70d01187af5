#!/bin/bash
# session 2, call 2: tensor-core cross-attention parity, mma.sync ceiling, ncu --set full of the flash fwd and row attn-store kernels
mkdir -p gpurun_out
echo "== mma.sync peak"; timeout 60 scripts/bin/mma_peak
echo "== kernel tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -q -s -k "cross_attn or self_attn" --timeout 400 2>&1 | grep -E "cross-attn|passed|failed|Error|error|assert" | cut -c1-250 | tail -30
echo "== kernel bench"; timeout 400 python scripts/kernel_bench.py --only cross_attn 2>&1 | cut -c1-260
echo "== full-size parity"; timeout 1200 python -m pytest tests/test_gpu_pipeline.py -q -s -k "full" --timeout 1000 2>&1 | grep -E "full-size|passed|failed|Error|assert" | cut -c1-300 | tail -20
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/q_bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','e2e','early_exit_images_per_s_1gpu','roofline')})
PY
echo "== ncu set full"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:capture_store_row --launch-skip 3 -c 1 -o gpurun_out/q_store_row python scripts/kernel_bench.py --only capture_store_fwd --reps 1 > gpurun_out/q_ncu1.log 2>&1; tail -2 gpurun_out/q_ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sa_fwd_kernel --launch-skip 3 -c 1 -o gpurun_out/q_sa_fwd python scripts/kernel_bench.py --only self_attn_fwd --reps 1 > gpurun_out/q_ncu2.log 2>&1; tail -2 gpurun_out/q_ncu2.log
ls -la gpurun_out/*.ncu-rep
