#!/bin/bash
# final verification of the round: full GPU suite + smoke, default bench (with CPU baseline), reference arm, step table
mkdir -p gpurun_out
bash scripts/gpu_tests.sh
echo "== default bench"; timeout 1200 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','e2e','early_exit_images_per_s_1gpu','gpu_launches','clocks')})
print('roofline', {k:d['roofline'].get(k) for k in ('achieved','peak','frac','ms_per_launch','traffic')})
print('roofline_gemm', {k:d['roofline_gemm'].get(k) for k in ('achieved','achieved_issued','peak','frac','frac_issued','ms_per_launch')})
print('cpu_baseline', {k:d['cpu_baseline'].get(k) for k in ('value','cores','kind')})
PY
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
echo "== table"; timeout 600 python scripts/profile_step.py --table gpurun_out/final_step_table.json --shapes gpurun_out/final_shapes.json > gpurun_out/final_table.log 2>&1; tail -1 gpurun_out/final_table.log
