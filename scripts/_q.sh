python scripts/profile_step.py --table gpurun_out/r4f_step_table.json --shapes gpurun_out/r4f_step_shapes.json > gpurun_out/r4f_profile.log 2>&1; tail -2 gpurun_out/r4f_profile.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sa_tc_bwd_kernel -s 0 -c 1 -o gpurun_out/r4f_attn_bwd python scripts/attn_bwd_bench.py --reps 1 > gpurun_out/r4f_ncu.log 2>&1
timeout 300 python scripts/attn_bwd_bench.py 2>&1 | tail -6
python bench.py --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
