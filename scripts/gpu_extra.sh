#!/bin/bash
mkdir -p gpurun_out
echo "== capture tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -q -k "capture or argsort or top_k or select" --timeout 300 2>&1 | tail -3
echo "== sanitizer"; timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python scripts/san_n500.py 2>&1 | grep -E "store|fused|ERROR SUMMARY|Invalid" | head -12
echo "== bench N=500 (graph)"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --tokens 500 > gpurun_out/n500g.json 2> gpurun_out/n500g.err; echo rc=$?; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/n500g.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('value','ms_per_step','early_exit_images_per_s_1gpu')}); print({k:d['roofline'].get(k) for k in ('achieved','frac','ms_per_launch','algorithmic_bytes')})
except Exception as e: print('no json', e)
PY
