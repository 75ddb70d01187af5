"""compute-sanitizer target: the attn-store / fused capture forwards at N=500 and N=77 (all four captured-layer shapes).
    compute-sanitizer --tool memcheck python scripts/sanitize_capture.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stablekeypoints_b200 import ops
dev = torch.device("cuda")
for n in (500, 77):
    lgs = [torch.randn(8, s * s, n, device=dev) * 3 for s in (16, 16, 16, 32)]
    for mode in ("store", "fused"):
        ops.CAPTURE_MEAN_FWD = mode
        m = ops.capture_mean(lgs, 128)
        torch.cuda.synchronize()
        print(n, mode, float(m.sum()))
    st = [ops.capture_store(l, 128) for l in lgs]
    torch.cuda.synchronize()
    print(n, "store ok", [float(s.sum()) for s in st])
