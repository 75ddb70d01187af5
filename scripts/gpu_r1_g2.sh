#!/bin/bash
# session 2, call 18 (4 GPUs): weak-scaling sanity of the final build (driver-style launch)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/g2_bench4.json 2> gpurun_out/g2_bench4.err; echo "rc=$?"; tail -2 gpurun_out/g2_bench4.err | cut -c1-200
python - <<'PY'
import json
for l in open('gpurun_out/g2_bench4.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({k:d.get(k) for k in ('value','n_gpus','ms_per_step','e2e','clocks')})
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29520 bench.py --impl reference --gpus 4 --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
