#!/bin/bash
# session 2, call 12: ncu launch list of one eager optimizer step (per-launch durations + grid sizes)
mkdir -p gpurun_out
timeout 1200 ncu --nvtx --nvtx-include "skp_step" --metrics gpu__time_duration.sum,launch__grid_size --clock-control none --csv --log-file gpurun_out/a2_launches.csv python scripts/profile_step.py > gpurun_out/a2_ncu.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/a2_ncu.log; wc -l gpurun_out/a2_launches.csv
