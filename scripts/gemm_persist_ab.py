#!/usr/bin/env python
"""A/B of the persistent two-accumulator form of the tcgen05 GEMM / implicit-GEMM convolution (gemm_nt_tc_persist_kernel)
against the one-tile-per-CTA kernel on the un-split problems of the step with more tiles than SMs.  GPU time per launch = a
CUDA graph of `reps` back-to-back launches (the situation inside the step graph); results must be bit-identical."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from stablekeypoints_b200 import ops
from stablekeypoints_b200._lib import lib

dev = torch.device("cuda")


def gpu_time(fn, reps=10):
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
    best = 1e30
    for _ in range(3):
        s.record(); g.replay(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) * 1e3 / reps)
    return best


rows = []
gen = torch.Generator(device="cuda").manual_seed(0)
# (kind, H, W, Cin, Cout) convolutions of the VAE encoder / UNet; (kind, M, N, K) projections
cases = [("conv", 512, 512, 128, 128), ("conv", 256, 256, 128, 256), ("conv", 256, 256, 256, 256), ("conv", 128, 128, 256, 512),
         ("conv", 128, 128, 512, 512), ("conv", 64, 64, 512, 512), ("conv", 64, 64, 320, 320), ("conv", 64, 64, 640, 320),
         ("gemm", 4096, 2560, 320), ("gemm", 4096, 960, 320), ("gemm", 4096, 1536, 512), ("gemm", 65536, 256, 128),
         ("gemm", 1024, 5120, 640), ("gemm", 16384, 512, 4608)]
for c in cases:
    if c[0] == "conv":
        _, h, w, cin, cout = c
        x = torch.randn(h * w, cin, device=dev, generator=gen)
        wt = torch.randn(cout, cin, 3, 3, device=dev, generator=gen) / (9 * cin) ** 0.5
        bias = torch.randn(cout, device=dev, generator=gen)
        fcw = ops.FrozenConv3x3(wt, need_dgrad=False)
        hi, lo = ops.split_bf16(x)
        fn = lambda: ops.conv3x3_implicit(hi, lo, h, w, fcw.fwd_split, cout, bias)
        flops = 2.0 * h * w * cout * 9 * cin
    else:
        _, m, n, k = c
        a = torch.randn(m, k, device=dev, generator=gen)
        b = torch.randn(n, k, device=dev, generator=gen) / k ** 0.5
        bias = torch.randn(n, device=dev, generator=gen)
        res = torch.randn(m, n, device=dev, generator=gen)
        fw = ops.FrozenWeight(b, need_dgrad=False)
        hi, lo = ops.split_bf16(a)
        fn = lambda: ops.gemm_nt_presplit(hi, lo, m, fw.w_split, n, bias, res)
        flops = 2.0 * m * n * k
    row = {"case": list(c)}
    outs = {}
    for mode in (0, 1):
        lib().skp_gemm_tc_persist(mode)
        outs[mode] = fn().clone()
        row["persist%d_us" % mode] = round(gpu_time(fn), 2)
    row["bit_identical"] = bool(torch.equal(outs[0], outs[1]))
    row["issued_frac_persist0"] = round(3 * flops / (row["persist0_us"] * 1e-6) / 1673e12, 3)
    row["issued_frac_persist1"] = round(3 * flops / (row["persist1_us"] * 1e-6) / 1673e12, 3)
    rows.append(row)
    print(json.dumps(row), flush=True)
lib().skp_gemm_tc_persist(0)
if "--out" in sys.argv:
    json.dump(rows, open(sys.argv[sys.argv.index("--out") + 1], "w"), indent=1)
