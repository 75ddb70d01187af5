#!/bin/bash
# session 2, call 9: VAE-prefetch pipeline in the step graph
mkdir -p gpurun_out
echo "== pipeline tests"; timeout 1500 python -m pytest tests/test_gpu_pipeline.py -x -q --timeout 900 2>&1 | tail -4 | cut -c1-300
echo "== bench prefetch"; timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err; tail -2 gpurun_out/x_bench.err | cut -c1-200; python - <<'PY'
import json
d=json.loads(open('gpurun_out/x_bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','e2e','early_exit_images_per_s_1gpu','gpu_launches')})
PY
echo "== bench no prefetch"; SKP_VAE_PREFETCH=0 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d.get(k) for k in ('value','ms_per_step','early_exit_images_per_s_1gpu')})"
