timeout 1500 python -m pytest tests/ -x -q -m gpu --timeout 900 -s 2>&1 | grep -E "parity|passed|failed|error" | cut -c1-400 | tail -20
( time timeout 1500 python bench.py > gpurun_out/r3c_bench.json 2> gpurun_out/r3c_bench.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3c_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','gpu_launches','roofline','cfg5_sdxl_shaped_1gpu','tokens_100_images_per_s_1gpu','tokens_500_images_per_s_1gpu','batch4_accum_images_per_s_1gpu','full_forward_images_per_s_1gpu','clocks'):
    print(k, json.dumps(d.get(k))[:400])
PY
