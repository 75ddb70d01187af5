#!/bin/bash
# the exact commands the driver runs at round end, with logs kept
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu --timeout 900 > gpurun_out/t_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t_gpu_tests.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/t_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/t_smoke.log | cut -c1-300
