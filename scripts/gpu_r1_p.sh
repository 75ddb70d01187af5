#!/bin/bash
# session 2, call 1: new self-attention + row attn-store kernels: parity, per-kernel timing, full-size parity, bench
mkdir -p gpurun_out
echo "== kernel tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -q -s -k "self_attn or capture_store or capture_matches" --timeout 400 2>&1 | grep -E "self-attn|passed|failed|Error|error|assert" | cut -c1-250 | tail -40
echo "== kernel bench"; timeout 400 python scripts/kernel_bench.py --only self_attn 2>&1 | cut -c1-260
timeout 300 python scripts/kernel_bench.py --only sdpa 2>&1 | cut -c1-260
timeout 300 python scripts/kernel_bench.py --only capture_store_fwd 2>&1 | cut -c1-260
SKP_CAPTURE_ROW=0 timeout 300 python scripts/kernel_bench.py --only capture_store_fwd 2>&1 | cut -c1-260
echo "== full-size parity"; timeout 1200 python -m pytest tests/test_gpu_pipeline.py -q -s -k "full or stage1" --timeout 1000 2>&1 | grep -E "full-size|passed|failed|Error|assert" | cut -c1-300 | tail -20
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err; tail -c 3000 gpurun_out/p_bench.json; tail -5 gpurun_out/p_bench.err
