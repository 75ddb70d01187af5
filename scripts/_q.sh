timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "self_attn_fwd_bwd and tcgen05" -s 2>&1 | grep -E "tcgen05\] S=(4096|128)|passed|failed|Error|error" | cut -c1-250 | tail -12
timeout 300 python scripts/attn_bwd_bench.py 2>&1 | grep tcgen05
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:sa_tc_bwd -c 4 --csv python scripts/attn_bwd_bench.py --reps 1 2>/dev/null | grep -E "sa_tc_bwd" | awk -F'","' '{print $1, $(NF-2), $NF}' | cut -c1-200 | head -8
