#!/bin/bash
mkdir -p gpurun_out
echo "== gemm + conv unit"; timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "gemm or conv" --timeout 300 2>&1 | tail -15 | tee gpurun_out/d_gemm.log
echo "== pipeline tiny"; timeout 600 python -m pytest tests/test_gpu_pipeline.py -q -k "tiny or engine" --timeout 300 2>&1 | tail -15 | tee gpurun_out/d_tiny.log
echo "== full parity"; timeout 900 python -m pytest tests/test_gpu_pipeline.py -q -s -k "full" --timeout 800 2>&1 | grep "full-size\|candidate\|passed\|failed\|Error" | tee gpurun_out/d_full.log
for prec in fp32 reference; do
echo "== bench trunk=tc precision=$prec"; timeout 900 python bench.py --steps 5 --warmup 4 --precision $prec --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d.get(k) for k in ('value','ms_per_step','e2e','early_exit_images_per_s_1gpu','skp_share_of_step','skp_kernel_ms_in_one_profiled_step')})" | tee gpurun_out/d_bench_$prec.log
done
