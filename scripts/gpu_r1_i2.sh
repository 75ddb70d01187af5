#!/bin/bash
# session 2, call 20: one-CTA-per-group fused GroupNorm
mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -x -q -k "groupnorm or conv or tiny or full or golden" --timeout 900 2>&1 | tail -3 | cut -c1-250
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d.get(k) for k in ('value','ms_per_step','early_exit_images_per_s_1gpu')})"
echo "== bench (group kernels off)"; SKP_GN_GROUP=0 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d.get(k) for k in ('value','ms_per_step')})"
echo "== table"; timeout 600 python scripts/profile_step.py --table gpurun_out/i2_step_table.json 2>&1 | grep -E "gn_|total GPU" | cut -c1-150
