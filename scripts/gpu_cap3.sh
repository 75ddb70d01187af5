#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k capture --timeout 200 2>&1 | tail -2
echo "== nocache(default)"; timeout 200 python scripts/kernel_bench.py --only capture_ 2>&1 | grep -v bwd | cut -c1-200
echo "== cache"; SKP_CAPTURE_CACHE=1 timeout 200 python scripts/kernel_bench.py --only capture_ 2>&1 | grep -v bwd | cut -c1-200
