mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "fps or selection" 2>&1 | tail -3
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r5d_bench_2gpu.json 2> gpurun_out/r5d_bench_2gpu.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r5d_bench_2gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['value'],2), round(d['ms_per_step'],3), round(d['e2e']['value'],2), d['clocks'])
PY
tail -3 gpurun_out/r5d_bench_2gpu.err | cut -c1-300
