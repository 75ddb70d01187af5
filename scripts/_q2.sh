mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "equivariance or losses or loss_api" 2>&1 | tail -5
timeout 300 python scripts/capture_soak.py --reps 300 > gpurun_out/r5b_capture_soak.log 2>&1; echo "soak rc=$?"; cat gpurun_out/r5b_capture_soak.log | cut -c1-250
timeout 120 python - <<'PY'
import torch
from stablekeypoints_b200 import ops
maps=torch.rand(77,128,128,device='cuda',requires_grad=True); mt=torch.rand(77,128,128,device='cuda',requires_grad=True)
sel=torch.arange(10,device='cuda'); th=torch.tensor([[0.85,0.2,0.1],[-0.2,0.85,-0.15]],device='cuda')
def step():
    l=ops.equivariance_loss_op(maps,mt,sel,th); l.backward(); maps.grad=None; mt.grad=None
for _ in range(5): step()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as p:
    for _ in range(20): step()
    torch.cuda.synchronize()
for e in p.key_averages():
    if 'equiv' in e.key: print(e.key[:60], e.count, round(e.device_time_total/e.count,2),'us')
PY
