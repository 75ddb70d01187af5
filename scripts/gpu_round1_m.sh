#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu --timeout 900 > gpurun_out/t_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/t_gpu_tests.log
timeout 1500 ncu --nvtx --nvtx-include "skp_step" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/m_launches.csv python scripts/profile_step.py > gpurun_out/m_ncu.log 2>&1
grep -c gpu__time_duration gpurun_out/m_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_nt_tc_kernel" --nvtx --nvtx-include "skp_step" -c 40 -o gpurun_out/m_gemm python scripts/profile_step.py > gpurun_out/m_ncu2.log 2>&1
ls -la gpurun_out/m_*
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err; tail -c 1500 gpurun_out/m_bench.json
