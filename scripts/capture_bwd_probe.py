#!/usr/bin/env python
"""capture_mean forward + backward on the SD1.5 captured layers (3 x 16x16 + 1 x 32x32, 8 heads, R = 128) for profilers / timing.
    python scripts/capture_bwd_probe.py [--tokens 500] [--reps 5]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from stablekeypoints_b200 import ops

n = int(sys.argv[sys.argv.index("--tokens") + 1]) if "--tokens" in sys.argv else 500
reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 5
dev = torch.device("cuda")
lg = [(torch.randn(8, s * s, n, device=dev) * 3).requires_grad_(True) for s in (16, 16, 16, 32)]
g = torch.randn(n, 128, 128, device=dev)
ts = []
for i in range(reps + 2):
    for l in lg:
        l.grad = None
    m = ops.capture_mean(lg, 128)
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.record()
    m.backward(g)
    en.record()
    torch.cuda.synchronize()
    if i >= 2:
        ts.append(st.elapsed_time(en))
print(json.dumps({"tokens": n, "capture_mean_backward_4_layers_us": round(1e3 * sum(ts) / len(ts), 1)}))
