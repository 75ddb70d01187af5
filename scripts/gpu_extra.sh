#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -x -q -k "capture_mean or capture_kernel_variants or tiny_stage1 or full" --timeout 600 2>&1 | tail -4 | cut -c1-300
echo "== kernel bench"; timeout 200 python scripts/kernel_bench.py --only capture_mean_bwd 2>&1 | cut -c1-200
SKP_CAPTURE_BWD_ROW=0 timeout 200 python scripts/kernel_bench.py --only capture_mean_bwd 2>&1 | cut -c1-200
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d.get(k) for k in ('value','ms_per_step','early_exit_images_per_s_1gpu')})"
