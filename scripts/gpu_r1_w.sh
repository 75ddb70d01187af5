#!/bin/bash
# session 2, call 8 (2 GPUs): weak-scaling bench with the two-stream step graph + captured NCCL all-reduce
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/w_bench2.json 2> gpurun_out/w_bench2.err; echo "rc=$?"; tail -3 gpurun_out/w_bench2.err | cut -c1-300
python - <<'PY'
import json
for l in open('gpurun_out/w_bench2.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({k:d.get(k) for k in ('value','n_gpus','ms_per_step','e2e','gpu_launches','clocks')})
PY
