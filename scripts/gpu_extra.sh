#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -x -q -k "capture or tiny or full or other_token" --timeout 600 2>&1 | tail -4 | cut -c1-300
echo "== sanitizer (bwd)"; cat > /tmp/san_bwd.py <<'PY'
import sys; sys.path.insert(0, '.')
import torch
from stablekeypoints_b200 import ops
for n in (77, 500):
    lgs = [(torch.randn(8, s * s, n, device="cuda") * 3).requires_grad_(True) for s in (16, 16, 16, 32)]
    m = ops.capture_mean(lgs, 128)
    g = torch.autograd.grad(m, lgs, torch.randn_like(m))
    torch.cuda.synchronize()
    print(n, [float(x.abs().sum()) for x in g])
PY
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python /tmp/san_bwd.py 2>&1 | grep -E "^77|^500|ERROR SUMMARY|Invalid" | head
