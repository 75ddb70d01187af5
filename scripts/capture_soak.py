#!/usr/bin/env python
"""Soak test of the capture kernels whose shared-memory rings are ordered by mbarriers only (compute-sanitizer's racecheck
cannot follow a generic-proxy read -> mbarrier arrive -> test_wait -> cp.async.bulk write chain and reports it as a hazard,
profiles/r02_sanitizer_final.md): the tcgen05 fused capture+collect forward, the tcgen05 store-mode kernel and the register
attn-store kernel run REPS times on the same logits while a second stream keeps the SMs and the memory system busy with
unrelated work; every repetition must be bit-identical to the first, and the first within 1e-3 of the SIMT row kernels.
    python scripts/capture_soak.py [--reps 300]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from stablekeypoints_b200 import ops
from stablekeypoints_b200._lib import lib

reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 300
dev = torch.device("cuda")
side = torch.cuda.Stream()
noise_a = torch.randn(2048, 2048, device=dev)
noise_b = torch.empty(64 << 20, device=dev)
bad = 0


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


for n in (77, 100, 500):
    g = torch.Generator(device="cuda").manual_seed(n)
    logits = [torch.randn(8, s * s, n, device=dev, generator=g) * 3 for s in (16, 16, 32, 32)]
    res = 128
    # references: SIMT row kernels
    lib().skp_capture_tc(0); lib().skp_capture_select(1, 1); ops.CAPTURE_MEAN_FWD = "store"
    m_ref = ops.capture_mean(logits, res)
    p_ref = ops.capture_store(logits[0], res)
    for name, tc, sel in (("tcgen05 (forced)", 2, 2), ("default plan", 1, 2)):
        lib().skp_capture_tc(tc); lib().skp_capture_select(sel, 1)
        m0 = ops.capture_mean(logits, res)
        p0 = ops.capture_store(logits[0], res)
        torch.cuda.synchronize()
        e_m, e_p = rel(m0, m_ref), rel(p0, p_ref)
        nrep = 0
        for i in range(reps):
            with torch.cuda.stream(side):      # unrelated load: perturbs the timing of every ring
                if i % 3 == 0:
                    noise_a @ noise_a
                elif i % 3 == 1:
                    noise_b.fill_(float(i))
            m = ops.capture_mean(logits, res)
            p = ops.capture_store(logits[0], res)
            if not (torch.equal(m, m0) and torch.equal(p, p0)):
                nrep += 1
        torch.cuda.synchronize()
        ok = e_m < 1e-3 and e_p < 1e-3 and nrep == 0
        bad += not ok
        print(f"N={n} {name}: vs SIMT row kernels maps {e_m:.2e} store {e_p:.2e}; {reps} repetitions under load, "
              f"{nrep} not bit-identical -> {'ok' if ok else 'FAIL'}", flush=True)
lib().skp_capture_select(2, 1); lib().skp_capture_tc(1)
sys.exit(1 if bad else 0)
