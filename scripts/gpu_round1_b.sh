#!/bin/bash
# Per-kernel roofline table, fp32-strict bench, ncu launch list + full captures of the top kernels.
mkdir -p gpurun_out
echo "== kernel bench"; timeout 600 python scripts/kernel_bench.py --json gpurun_out/b_kernels.json 2>&1 | tail -60 | tee gpurun_out/b_kernels.log
echo "== bench fp32"; timeout 900 python bench.py --steps 5 --warmup 3 --precision fp32 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/b_bench_fp32.log
echo "== ncu launch list"; timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/b_launches.csv python bench.py --steps 1 --warmup 1 --precision fp32 --no-cpu-baseline > gpurun_out/b_ncu_bench.log 2>&1
tail -2 gpurun_out/b_ncu_bench.log
echo "== ncu full: capture + gemm + attn"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"capture_fwd_kernel|gemm_nt_tc_kernel|cross_attn_fwd_kernel|cross_attn_bwd_dq|capture_bwd_kernel" -c 24 -o gpurun_out/b_prof python scripts/kernel_bench.py --reps 1 > gpurun_out/b_ncu_full.log 2>&1
ls -la gpurun_out/ | tail -12
