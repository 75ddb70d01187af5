#!/bin/bash
mkdir -p gpurun_out
echo "== pipeline tiny"; timeout 600 python -m pytest tests/test_gpu_pipeline.py -q -k "tiny or engine" --timeout 300 2>&1 | tail -15 | tee gpurun_out/e_tiny.log
echo "== kernels"; timeout 600 python -m pytest tests/test_gpu_kernels.py -q --timeout 300 2>&1 | tail -5 | tee gpurun_out/e_kernels.log
echo "== full parity"; timeout 900 python -m pytest tests/test_gpu_pipeline.py -q -s -k "full" --timeout 800 2>&1 | grep "full-size\|candidate\|passed\|failed\|Error" | tee gpurun_out/e_full.log
for mode in "--precision fp32" "--precision fp32 --no-graph"; do
echo "== bench $mode"; timeout 900 python bench.py --steps 10 --warmup 3 $mode --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d.get(k) for k in ('value','ms_per_step','e2e','gpu_launches','early_exit_images_per_s_1gpu','skp_kernel_ms_total','skp_share_of_timed_step','skp_kernel_ms_in_one_profiled_step')})" | tee -a gpurun_out/e_bench.log
done
