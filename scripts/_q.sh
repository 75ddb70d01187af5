timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r4o_bench_2gpu.json 2> gpurun_out/r4o_bench_2gpu.err; echo rc=$?
( time timeout 1500 python bench.py > gpurun_out/r4o_bench.json 2> gpurun_out/r4o_bench.err ) 2>&1 | tail -3
python - <<'PY'
import json
for f in ('gpurun_out/r4o_bench_2gpu.json','gpurun_out/r4o_bench.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['n_gpus'], round(d['value'],2), round(d['ms_per_step'],3), round(d['e2e']['value'],2), d['clocks'])
d=json.loads(open('gpurun_out/r4o_bench.json').read().strip().splitlines()[-1])
for k in ('roofline','tokens_100_images_per_s_1gpu','tokens_500_images_per_s_1gpu','batch4_accum_images_per_s_1gpu','full_forward_images_per_s_1gpu'):
    print(k, json.dumps(d.get(k))[:300])
print(d['cfg5_sdxl_shaped_1gpu']['images_per_s'])
PY
