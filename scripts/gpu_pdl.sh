#!/bin/bash
# PDL experiment: GEMM / conv / split / split-K reduce launched with programmatic stream serialization
mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -x -q -k "gemm or conv or capture_store or tiny or full" --timeout 900 2>&1 | tail -3 | cut -c1-250
echo "== bench PDL on"; timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d.get(k) for k in ('value','ms_per_step','early_exit_images_per_s_1gpu')})"
echo "== bench PDL off"; SKP_PDL=0 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d.get(k) for k in ('value','ms_per_step')})"
