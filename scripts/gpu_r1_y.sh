#!/bin/bash
# session 2, call 10: stride-2 dgrad on our kernel, capture_mean forward A/B, per-shape table incl. GroupNorm
mkdir -p gpurun_out
echo "== kernel tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -q -k "conv or capture" --timeout 400 2>&1 | tail -3 | cut -c1-250
echo "== kernel bench"; timeout 300 python scripts/kernel_bench.py --only capture_mean 2>&1 | cut -c1-260
echo "== bench (mean fwd = store+collect)"; SKP_CAPTURE_MEAN_FWD=store timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d.get(k) for k in ('value','ms_per_step','early_exit_images_per_s_1gpu')})"
echo "== bench (default)"; timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d.get(k) for k in ('value','ms_per_step','early_exit_images_per_s_1gpu')})"
echo "== shapes"; timeout 600 python scripts/profile_step.py --shapes gpurun_out/y_shapes.json --table gpurun_out/y_step_table.json 2>&1 | grep -E "gn_|split_bf16|ln_split|total GPU" | cut -c1-160 | head -70
