#!/bin/bash
mkdir -p gpurun_out
echo "== kernels"; timeout 900 python -m pytest tests/test_gpu_kernels.py -q --timeout 300 2>&1 | tail -8 | tee gpurun_out/j_kernels.log
echo "== pipeline"; timeout 900 python -m pytest tests/test_gpu_pipeline.py -q -s --timeout 800 2>&1 | grep -E "full-size|candidate|passed|failed|Error" | tee gpurun_out/j_pipe.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d.get(k) for k in ('value','ms_per_step','e2e','gpu_launches','early_exit_images_per_s_1gpu','skp_kernel_ms_total','skp_share_of_timed_step','roofline','skp_kernel_ms_in_one_profiled_step')})" | tee gpurun_out/j_bench.log
