timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "capture" 2>&1 | tail -2
for n in 77 500; do echo "tokens=$n"; timeout 600 python bench.py --tokens $n --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['gpu_launches'])"; done
python scripts/profile_step.py --tokens 500 --table gpurun_out/r2p_step_table_n500.json 2>/dev/null | tail -34 | head -24
