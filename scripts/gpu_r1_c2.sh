#!/bin/bash
# session 2, call 14: im2col kernels (9-tap ILP / small-channel variant)
mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -x -q -k "conv or groupnorm or full or tiny" --timeout 900 2>&1 | tail -3 | cut -c1-250
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d.get(k) for k in ('value','ms_per_step','early_exit_images_per_s_1gpu')})"
echo "== table"; timeout 600 python scripts/profile_step.py --table gpurun_out/c2_step_table.json 2>&1 | grep -E "im2col|total GPU" | cut -c1-150
