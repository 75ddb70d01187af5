mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests/ -x -q -m gpu --timeout 900 -s > gpurun_out/r5j_gpu_tests.log 2>&1; grep -E "parity|passed|failed|error" gpurun_out/r5j_gpu_tests.log | cut -c1-300 | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1 | cut -c1-300
( time timeout 1500 python bench.py > gpurun_out/r5j_bench.json 2> gpurun_out/r5j_bench.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r5j_bench.json').read().strip().splitlines()[-1])
print(round(d['value'],2), round(d['ms_per_step'],3), round(d['e2e']['value'],2), d['clocks'])
for k in ('tokens_100_images_per_s_1gpu','tokens_500_images_per_s_1gpu','batch4_accum_images_per_s_1gpu','full_forward_images_per_s_1gpu'): print(k, d.get(k))
print(d['roofline']['frac'], d['cfg5_sdxl_shaped_1gpu']['images_per_s'], d['gpu_launches'], d['roofline_gemm'])
PY

