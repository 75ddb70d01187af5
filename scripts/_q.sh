timeout 600 python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu -k "two_rank" 2>&1 | tail -2 | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r4g_bench_2gpu.json 2> gpurun_out/r4g_bench_2gpu.err; echo rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r4g_bench_2gpu.json').read().strip().splitlines()[-1])
for k in ('value','n_gpus','ms_per_step','e2e','clocks'): print(k, json.dumps(d.get(k))[:300])
PY
python bench.py --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('1gpu', d['value'], d['ms_per_step'], d['e2e']['value'])"
