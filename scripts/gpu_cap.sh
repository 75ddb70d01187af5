#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k capture --timeout 200 2>&1 | tail -3
echo "== cache"; timeout 200 python scripts/kernel_bench.py --only capture_ 2>&1 | grep -v bwd | cut -c1-200
echo "== nocache"; SKP_CAPTURE_CACHE=0 timeout 200 python scripts/kernel_bench.py --only capture_ 2>&1 | grep -v bwd | cut -c1-200
