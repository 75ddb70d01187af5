#!/usr/bin/env python
"""Soak test of the tcgen05 self-attention kernels: repeated forward + backward on the step's shapes, every repetition compared
with the first (bit-reproducibility) and the first with the mma.sync kernels.
    python scripts/attn_soak.py [--reps 40]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from stablekeypoints_b200 import ops

reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 40
dev = torch.device("cuda")
bad = 0
for (s, heads, d) in [(4096, 8, 40), (1024, 8, 80), (1024, 20, 64), (4096, 20, 32)]:
    c = heads * d
    g = torch.Generator(device="cuda").manual_seed(s + d)
    qkv = torch.randn(s, 3 * c, device=dev, generator=g) * 1.5
    do = torch.randn(s, c, device=dev, generator=g)

    def run():
        x = qkv.clone().requires_grad_(True)
        o = ops.self_attn_core(x, heads, d ** -0.5)
        o.backward(do)
        return o.detach(), x.grad

    ops.SELF_ATTN_TC, ops.SELF_ATTN_TC_BWD = False, False
    o_ref, g_ref = run()
    ops.SELF_ATTN_TC, ops.SELF_ATTN_TC_BWD = True, True
    o0, g0 = run()
    e_o = float((o0 - o_ref).abs().max() / o_ref.abs().max())
    e_g = float((g0 - g_ref).abs().max() / g_ref.abs().max())
    nrep = 0
    for i in range(reps):
        o, gg = run()
        if not (torch.equal(o, o0) and torch.equal(gg, g0)):
            nrep += 1
    ok = e_o < 1e-4 and e_g < 1e-4 and nrep == 0
    bad += not ok
    print(f"S={s} h={heads} d={d}: tcgen05 vs mma.sync o {e_o:.2e} dqkv {e_g:.2e}; {reps} repetitions, {nrep} not bit-identical -> {'ok' if ok else 'FAIL'}", flush=True)
sys.exit(1 if bad else 0)
