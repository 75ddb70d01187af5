#!/usr/bin/env python
"""Write-only HBM bandwidth on this GPU (the attn-store kernel is a pure write stream: 404 MB out, 6 MB in): torch fill_ and
cudaMemset of the cfg5 store size, L2 flushed between launches, against MEASURED_PEAKS.json's copy figure."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
dev = torch.device("cuda")
flush = torch.empty(512 * 1024 * 1024 // 4, device=dev)
n = 20 * 256 * 256 * 77
x = torch.empty(n, device=dev)
y = torch.empty(n, device=dev)
for name, fn, nbytes in (("fill_ (write only)", lambda: x.fill_(1.5), 4 * n), ("zero_ (memset)", lambda: x.zero_(), 4 * n),
                         ("copy_ (read + write)", lambda: y.copy_(x), 8 * n)):
    ts = []
    for i in range(7):
        flush.fill_(float(i))
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        if i >= 2:
            ts.append(s.elapsed_time(e))
    ms = sum(ts) / len(ts)
    print(json.dumps({"op": name, "MB": round(nbytes / 1e6, 1), "us": round(ms * 1e3, 1), "GBps": round(nbytes / ms / 1e6, 1),
                      "frac_of_copy_peak": round(nbytes / ms / 1e6 / peak, 3)}), flush=True)
