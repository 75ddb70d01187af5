#!/bin/bash
mkdir -p gpurun_out
echo "== token-count tests + capture shapes"; timeout 1500 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_kernels.py -x -q -s -k "other_token_counts or capture_store" --timeout 900 2>&1 | grep -E "\[tiny|passed|failed|Error|assert" | cut -c1-250 | tail
echo "== kernel bench cfg5"; timeout 300 python scripts/kernel_bench.py --only "capture_store_fwd (cfg5" 2>&1 | cut -c1-260
echo "== bench N=500"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --tokens 500 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d.get(k) for k in ('value','ms_per_step','early_exit_images_per_s_1gpu')}); print(d['roofline'])"
echo "== bench N=100"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --tokens 100 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d.get(k) for k in ('value','ms_per_step')}); print({k:d['roofline'].get(k) for k in ('achieved','frac','ms_per_launch')})"
