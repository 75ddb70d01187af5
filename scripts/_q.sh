for v in 0 1; do echo "SKP_CTX_REPEAT=$v"; SKP_CTX_REPEAT=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['gpu_launches'])"; done
