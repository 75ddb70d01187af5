#!/bin/bash
# session 2, call 16: tcgen05 self-attention forward
mkdir -p gpurun_out
echo "== tc attention tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -s -k "self_attn" --timeout 120 2>&1 | grep -E "self-attn|passed|failed|Error|error|assert|Timeout" | cut -c1-220 | tail -40
echo "== kernel bench"; timeout 300 python scripts/kernel_bench.py --only self_attn_fwd 2>&1 | cut -c1-200
echo "== full parity"; timeout 900 python -m pytest tests/test_gpu_pipeline.py -q -s -k "full" --timeout 600 2>&1 | grep -E "full-size|passed|failed|Error|assert" | cut -c1-300 | tail
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d.get(k) for k in ('value','ms_per_step','early_exit_images_per_s_1gpu')})"
