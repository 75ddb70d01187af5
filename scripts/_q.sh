timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "groupnorm or conv3x3" 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -x -k "tiny_stage1 or store_dump or full_sd15 or graph" 2>&1 | tail -3
for v in 8 4 1; do echo "SKP_GN_CLUSTER=$v"; SKP_GN_CLUSTER=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['gpu_launches'])"; done
