timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "cross_attn or self_attn" 2>&1 | tail -12
timeout 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -x -k "tiny_stage1 or store_dump or full_sd15" 2>&1 | tail -6
for v in 1 0; do echo "SKP_XATTN_TC=$v"; SKP_XATTN_TC=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['gpu_launches'])"; done
