timeout 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -x -k "tiny_stage1 or store_dump or full_sd15-77 or graph" 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['gpu_launches'])"
python scripts/profile_step.py --tokens 500 --table gpurun_out/r2m_step_table_n500.json 2>/dev/null | tail -34
