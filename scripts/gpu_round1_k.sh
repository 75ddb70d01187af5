#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/profile_step.py --table gpurun_out/k_step_table.json 2>&1 | tail -36 | tee gpurun_out/k_table.log
timeout 1500 ncu --nvtx --nvtx-include "skp_step/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/k_launches.csv python scripts/profile_step.py > gpurun_out/k_ncu.log 2>&1
ls -la gpurun_out/k_launches.csv; wc -l gpurun_out/k_launches.csv
