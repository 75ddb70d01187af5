timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "capture" 2>&1 | tail -2
python - <<'PY'
import torch, sys
sys.path.insert(0,'.')
from stablekeypoints_b200 import ops
from stablekeypoints_b200._lib import lib
flush=torch.empty(512*1024*1024//4,device='cuda')
for (h,s,n,r) in [(8,16,500,128),(8,32,500,128),(8,16,100,128),(8,16,77,128)]:
    lg=torch.randn(h,s*s,n,device='cuda')*3
    lib().skp_capture_tc(0)
    ts=[]
    for i in range(8):
        flush.fill_(float(i)); a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); ops.capture_store(lg,r); b.record(); torch.cuda.synchronize()
        if i>=3: ts.append(a.elapsed_time(b))
    ms=sum(ts)/len(ts); print((h,s,n,r), round(ms*1e3,1),'us', round((h*r*r*n*4)/ms/1e6,1),'GB/s')
lib().skp_capture_tc(1)
PY
