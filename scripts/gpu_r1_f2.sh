#!/bin/bash
# session 2, call 17: full suite + default bench (with CPU baseline) + step table + ncu of the tcgen05 attention kernel
mkdir -p gpurun_out
bash scripts/gpu_tests.sh
echo "== default bench"; timeout 1200 python bench.py > gpurun_out/f2_bench.json 2> gpurun_out/f2_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/f2_bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','e2e','early_exit_images_per_s_1gpu','gpu_launches','roofline','roofline_gemm','cpu_baseline','clocks')})
PY
echo "== table"; timeout 600 python scripts/profile_step.py --table gpurun_out/f2_step_table.json 2>&1 | grep -E "sa_tc|total GPU" | cut -c1-150
echo "== ncu"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:sa_tc_fwd_kernel --launch-skip 3 -c 1 -o gpurun_out/f2_sa_tc python scripts/kernel_bench.py --only self_attn_fwd --reps 1 > gpurun_out/f2_ncu.log 2>&1; tail -2 gpurun_out/f2_ncu.log
