timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "self_attn_fwd_bwd" -s 2>&1 | grep -E "tcgen05\]|passed|failed|Error|error" | cut -c1-250 | tail -30
timeout 300 python scripts/attn_bwd_bench.py 2>&1 | tail -8
