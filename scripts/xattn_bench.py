#!/usr/bin/env python
"""Cross-attention forward: tcgen05 kernel (skp_xattn_tc.cu) vs the mma.sync flash kernel (skp_selfattn.cu), GPU time per
call measured by replaying a CUDA graph of `reps` back-to-back calls (no launch gaps), plus the error against fp64."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stablekeypoints_b200 import ops  # noqa: E402


def ref(q, k, v, heads, scale):
    s, c = q.shape
    n = k.shape[0]
    d = c // heads
    qh = q.double().reshape(s, heads, d).permute(1, 0, 2)
    kh = k.double().reshape(n, heads, d).permute(1, 0, 2)
    vh = v.double().reshape(n, heads, d).permute(1, 0, 2)
    lg = qh @ kh.transpose(1, 2) * scale
    return (torch.softmax(lg, -1) @ vh).permute(1, 0, 2).reshape(s, c), lg


def gpu_time(fn, reps=20):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); g.replay(); e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e3 / reps


rows = []
for (S, N, heads, d, logits) in [(4096, 77, 8, 40, False), (1024, 77, 8, 80, False), (1024, 77, 8, 80, True), (256, 77, 8, 160, False),
                                 (256, 77, 8, 160, True), (64, 77, 8, 160, False), (1024, 500, 8, 80, True), (256, 256, 8, 160, False)]:
    g = torch.Generator().manual_seed(S + N)
    q = torch.randn(S, heads * d, generator=g).cuda()
    k = torch.randn(N, heads * d, generator=g).cuda()
    v = torch.randn(N, heads * d, generator=g).cuda()
    scale = d ** -0.5
    want, lgw = ref(q.cpu(), k.cpu(), v.cpu(), heads, scale)
    row = {"S": S, "N": N, "heads": heads, "d": d, "logits": logits}
    for tc in (True, False):
        ops.XATTN_TC = tc
        with torch.no_grad():
            o, lg = ops.cross_attn_core(q, k, v, heads, scale, want_logits=logits)
            key = "tcgen05" if tc else "mma_sync"
            row[key + "_err"] = float((o.cpu().double() - want).abs().max() / want.abs().max())
            if logits:
                row[key + "_logits_err"] = float((lg.cpu().double() - lgw).abs().max() / lgw.abs().max())
            row[key + "_us"] = round(gpu_time(lambda: ops.cross_attn_core(q, k, v, heads, scale, want_logits=logits)), 2)
    ops.XATTN_TC = True
    rows.append(row)
    print(json.dumps(row), flush=True)
if "--json" in sys.argv:
    json.dump(rows, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
