#!/bin/bash
# session 2, call 4: MMA issue-order refactor of the flash kernels, 2-slice row attn-store kernel; ncu of the S=4096 flash fwd
mkdir -p gpurun_out
echo "== kernel tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -q -k "self_attn or cross_attn or capture" --timeout 400 2>&1 | tail -5 | cut -c1-250
echo "== kernel bench"; timeout 300 python scripts/kernel_bench.py --only capture_store_fwd 2>&1 | cut -c1-260
timeout 300 python scripts/kernel_bench.py --only self_attn 2>&1 | cut -c1-260
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/s_bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','e2e','early_exit_images_per_s_1gpu','roofline','gpu_launches')})
PY
echo "== ncu set full"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:sa_fwd_kernel<\(int\)48' --launch-skip 3 -c 1 -o gpurun_out/s_sa_fwd python scripts/kernel_bench.py --only self_attn_fwd --reps 1 > gpurun_out/s_ncu2.log 2>&1; tail -2 gpurun_out/s_ncu2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:capture_store_row --launch-skip 3 -c 1 -o gpurun_out/s_store_row python scripts/kernel_bench.py --only capture_store_fwd --reps 1 > gpurun_out/s_ncu1.log 2>&1; tail -2 gpurun_out/s_ncu1.log
ls -la gpurun_out/s_*.ncu-rep
