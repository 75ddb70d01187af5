#!/bin/bash
# session 2, call 6: 2-rows-per-CTA row attn-store kernel; graph-timed GEMM tile sweep
mkdir -p gpurun_out
echo "== kernel tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -q -k "capture or dense" --timeout 400 2>&1 | tail -3 | cut -c1-250
echo "== kernel bench"; timeout 300 python scripts/kernel_bench.py --only capture_store_fwd 2>&1 | cut -c1-260
echo "== gemm sweep"; timeout 1500 python scripts/gemm_sweep.py --top 48 --emit gpurun_out/skp_gemm_tuned.inc > gpurun_out/u_sweep.log 2>&1; tail -60 gpurun_out/u_sweep.log | cut -c1-250
