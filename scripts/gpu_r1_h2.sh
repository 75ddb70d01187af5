#!/bin/bash
# session 2, call 19: parallel loss reductions / FPS kernel
mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -x -q -k "loss or fps or furthest or select or sharpen or equiv or tiny or full or golden" --timeout 900 2>&1 | tail -3 | cut -c1-250
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d.get(k) for k in ('value','ms_per_step','early_exit_images_per_s_1gpu')})"
