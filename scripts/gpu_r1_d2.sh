#!/bin/bash
# session 2, call 15: ncu launch list of one eager optimizer step with the final kernels + step table + kernel bench table
mkdir -p gpurun_out
timeout 1200 ncu --nvtx --nvtx-include "skp_step" --metrics gpu__time_duration.sum,launch__grid_size --clock-control none --csv --log-file gpurun_out/d2_launches.csv python scripts/profile_step.py > gpurun_out/d2_ncu.log 2>&1; echo "rc=$?"; wc -l gpurun_out/d2_launches.csv
timeout 600 python scripts/profile_step.py --table gpurun_out/d2_step_table.json --shapes gpurun_out/d2_shapes.json > gpurun_out/d2_table.log 2>&1; tail -1 gpurun_out/d2_table.log
timeout 900 python scripts/kernel_bench.py --json gpurun_out/d2_kernel_bench.json > gpurun_out/d2_kernel_bench.log 2>&1; tail -3 gpurun_out/d2_kernel_bench.log | cut -c1-200
