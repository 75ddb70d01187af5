#!/bin/bash
# First GPU contact: kernel parity (SIMT paths first, tcgen05 GEMM isolated behind timeouts), pipeline parity, smoke, bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/a_gpu.txt 2>&1
lscpu | head -20 >> gpurun_out/a_gpu.txt
echo "== kernels (no tc)"; timeout 900 python -m pytest tests/test_gpu_kernels.py -q -k "not tc" --timeout 300 2>&1 | tail -40 | tee gpurun_out/a_kernels.log
echo "== tcgen05 gemm"; timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "tc" --timeout 120 2>&1 | tail -40 | tee gpurun_out/a_tc.log
echo "== pipeline tiny simt"; SKP_GEMM_IMPL=simt timeout 600 python -m pytest tests/test_gpu_pipeline.py -q -k "tiny and not tc" --timeout 300 2>&1 | tail -30 | tee gpurun_out/a_pipe_simt.log
echo "== pipeline tiny tc"; timeout 600 python -m pytest tests/test_gpu_pipeline.py -q -k "tiny and tc or engine" --timeout 300 2>&1 | tail -30 | tee gpurun_out/a_pipe_tc.log
echo "== pipeline full"; timeout 900 python -m pytest tests/test_gpu_pipeline.py -q -s -k "full" --timeout 800 2>&1 | tail -30 | tee gpurun_out/a_pipe_full.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/a_smoke.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -5 | tee gpurun_out/a_bench.log
