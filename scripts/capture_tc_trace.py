#!/usr/bin/env python
"""Per-role pipeline trace of CTA 0 of the tcgen05 attn-store kernel (clock64 stamps): where does a row's time go?
    python scripts/capture_tc_trace.py [s] [N] [R] [heads] [mode]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stablekeypoints_b200 import ops  # noqa: E402
from stablekeypoints_b200._lib import lib  # noqa: E402

s, n, r, h = (int(sys.argv[i]) if len(sys.argv) > i else d for i, d in ((1, 16), (2, 77), (3, 128), (4, 8)))
mode = sys.argv[5] if len(sys.argv) > 5 else "store"
lg = torch.randn(h, s * s, n, device="cuda") * 3
buf = torch.zeros(3 * 64 * 8, dtype=torch.int64, device="cuda")
for _ in range(3):
    ops.capture_store(lg, r) if mode == "store" else ops.capture_mean([lg, lg, lg, lg], r)
torch.cuda.synchronize()
lib().skp_capture_tc_trace(buf.data_ptr())
ops.capture_store(lg, r) if mode == "store" else ops.capture_mean([lg, lg, lg, lg], r)
torch.cuda.synchronize()
lib().skp_capture_tc_trace(None)
t = buf.cpu().reshape(3, 64, 8)
c0, g0, c1, g1 = (int(v) for v in t[2, 62, :4])
if g1 > g0:
    print(f"CTA 0: {c1 - c0} SM cycles in {g1 - g0} ns -> {1e3 * (c1 - c0) / (g1 - g0):.0f} MHz; kernel start -> first role stamp "
          f"{int(t[:, :62][t[:, :62] > 0].min()) - c0} cycles")
t[2, 62] = 0
t0 = int(t[t > 0].min())
names = {0: "producer: start, got EMPTY, stores done, arrived", 1: "mma: start, got AEMPTY, got FULL, committed",
         2: "epilogue: start, got AFULL, ld done, exp done, bulk-wait done, bar1 done, staged | end"}
for role in range(3):
    print(names[role])
    for it in range(64):
        row = t[role, it]
        if int(row.max()) == 0:
            continue
        print(f"  it {it:2d}: " + " ".join(f"{int(v) - t0:7d}" if int(v) > 0 else "      -" for v in row))
