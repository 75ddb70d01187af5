#!/bin/bash
# session 2, call 13: vectorised split / split-K reduce, GroupNorm sizing + replicated accumulators, register-cached LayerNorm
mkdir -p gpurun_out
echo "== all gpu tests"; timeout 1500 python -m pytest tests/ -x -q -m gpu --timeout 900 > gpurun_out/b2_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/b2_gpu_tests.log | cut -c1-300
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d.get(k) for k in ('value','ms_per_step','early_exit_images_per_s_1gpu')})"
echo "== table"; timeout 600 python scripts/profile_step.py --table gpurun_out/b2_step_table.json 2>&1 | cut -c1-150 | head -34
