#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k capture --timeout 200 2>&1 | tail -2
echo "== cache"; timeout 200 python scripts/kernel_bench.py --only capture_ 2>&1 | grep -v bwd | cut -c1-200
echo "== nocache"; SKP_CAPTURE_CACHE=0 timeout 200 python scripts/kernel_bench.py --only capture_ 2>&1 | grep -v bwd | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"capture_fwd_quad" -s 6 -c 2 -o gpurun_out/h_capq python scripts/kernel_bench.py --only capture_store_fwd --reps 2 > gpurun_out/h_ncu.log 2>&1
ls -la gpurun_out/h_capq* 
