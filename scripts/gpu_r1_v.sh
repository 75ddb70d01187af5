#!/bin/bash
# session 2, call 7: tuned GEMM plans + two-stream step graph: full GPU test-suite, smoke, bench (two-stream on / off)
mkdir -p gpurun_out
echo "== all gpu tests"; timeout 1500 python -m pytest tests/ -x -q -m gpu --timeout 900 > gpurun_out/v_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/v_gpu_tests.log | cut -c1-300
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2 | cut -c1-300
echo "== bench two-stream"; timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; tail -3 gpurun_out/v_bench.err | cut -c1-300; python - <<'PY'
import json
d=json.loads(open('gpurun_out/v_bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','e2e','early_exit_images_per_s_1gpu','gpu_launches','roofline_gemm')})
PY
echo "== bench one-stream"; SKP_TWO_STREAM=0 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d.get(k) for k in ('value','ms_per_step','early_exit_images_per_s_1gpu')})"
