for i in 1 2 3 4 5 6; do python scripts/_fwd_dbg.py 2>&1 | grep -vE "'[0-9.]+e-05', '[0-9.]+e-05', '[0-9.]+e-05'\] bad rows 0" | cut -c1-120; done; echo soak-done
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "self_attn" 2>&1 | tail -1
python scripts/attn_bwd_bench.py 2>&1 | grep forward
SKP_ATTN_FWD_PIPE=0 python scripts/attn_bwd_bench.py 2>&1 | grep forward | head -1
