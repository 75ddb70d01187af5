#!/usr/bin/env python
"""Large-problem throughput of the tcgen05 split-bf16 GEMM per N tile (is a small BN shared-memory-bound?).
    python scripts/gemm_bn_probe.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from stablekeypoints_b200._lib import check, lib, ptr, stream

dev = torch.device("cuda")
L = lib()
M, N, K = 16384, 1280, 4608
a_hi = torch.randn(M, K, device=dev).bfloat16(); a_lo = (torch.randn(M, K, device=dev) * 1e-3).bfloat16()
b_hi = torch.randn(N, K, device=dev).bfloat16(); b_lo = (torch.randn(N, K, device=dev) * 1e-3).bfloat16()
out = torch.empty(M, N, device=dev)
for bn in (64, 96, 128, 160, 256):
    L.skp_gemm_tc_force_bn(bn)
    ts = []
    for i in range(6):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        check(L.skp_gemm_nt_tc(ptr(a_hi), ptr(a_lo), ptr(b_hi), ptr(b_lo), K, ptr(out), N, M, N, 1.0, None, None, 0, 1, None, stream()), "gemm")
        e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ms = sorted(ts[1:])[2]
    print(json.dumps({"BN": bn, "ms": round(ms, 4), "algorithmic_TFLOPs": round(2.0 * M * N * K / ms / 1e9, 1), "issued_TFLOPs": round(6.0 * M * N * K / ms / 1e9, 1)}), flush=True)
L.skp_gemm_tc_force_bn(0)
