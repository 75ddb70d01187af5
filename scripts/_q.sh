timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "self_attn_fwd_bwd and tcgen05" 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sa_tc_bwd -c 4 --csv python scripts/attn_bwd_bench.py --reps 1 2>/dev/null | grep -E "sa_tc_bwd" | awk -F'","' '{print substr($5,1,60), $NF}' | head -8
