#!/usr/bin/env python
"""how much of the parity headroom can the (no-grad) VAE encoder spend?  Emulates reduced-term products on
the existing split-bf16 tcgen05 kernel by zeroing `lo` operand planes inside the VAE only:
    full      : hi.hi + hi.lo + lo.hi  (3 MMAs, the shipped numerics)
    w_hi      : weights rounded to bf16, activations split     (2 MMAs: a_hi.w_hi + a_lo.w_hi)
    a_hi      : activations rounded to bf16, weights split     (2 MMAs: a_hi.w_hi + a_hi.w_lo)
    bf16      : both rounded to bf16                            (1 MMA)
and reports the latent error and the error of the captured maps (through the exact UNet) against the full path."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stablekeypoints_b200 import ops, optimize_token, ptp_utils  # noqa: E402

dev = torch.device("cuda")
ldm, controllers, _ = optimize_token.load_ldm("cuda:0", "synthetic:0", feature_upsample_res=128, attn_gain=4.0, precision="fp32")
g = torch.Generator().manual_seed(5)
from stablekeypoints_b200.optimize import SyntheticKeypointDataset  # noqa: E402
img = SyntheticKeypointDataset(length=2, seed=3)[0]["img"][None].to(dev)
ctx = torch.randn(1, 77, 768, generator=g).to(dev)
noise = torch.randn(1, 4, 64, 64, generator=g).to(dev)

real_gemm, real_conv = ops.gemm_nt_presplit, ops.conv3x3_implicit
vae = ldm.vae
saved = {}


def set_mode(mode):
    # weights
    for name, obj in list(vae._conv.items()) + list(vae._fw.items()):
        planes = obj.fwd_split if hasattr(obj, "fwd_split") else obj.w_split
        if name not in saved:
            saved[name] = planes[1].clone()
        planes[1].copy_(saved[name])
        if mode in ("w_hi", "bf16"):
            planes[1].zero_()
    # activations: the lo plane of every A operand that reaches the GEMM / implicit conv while the VAE runs
    if mode in ("a_hi", "bf16"):
        def gemm(a_hi, a_lo, *a, **k):
            a_lo.zero_()
            return real_gemm(a_hi, a_lo, *a, **k)

        def conv(x_hi, x_lo, *a, **k):
            x_lo.zero_()
            return real_conv(x_hi, x_lo, *a, **k)
        ops.gemm_nt_presplit, ops.conv3x3_implicit = gemm, conv
    else:
        ops.gemm_nt_presplit, ops.conv3x3_implicit = real_gemm, real_conv


def run():
    with torch.no_grad():
        lat = ptp_utils.image2latent(ldm, img, "cuda")
        ops.gemm_nt_presplit, ops.conv3x3_implicit = real_gemm, real_conv      # the UNet stays exact: the latent is fed in
        maps = ptp_utils.run_and_find_attn(ldm, lat, ctx, layers=[0, 1, 2, 3], upsample_res=-1, controllers=controllers, noise=noise)[0]
    return lat, maps


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


set_mode("full")
lat0, maps0 = run()
out = {}
for mode in ("full", "w_hi", "a_hi", "bf16"):
    set_mode(mode)
    lat, maps = run()
    out[mode] = {"latent_rel_err": rel(lat, lat0), "maps_rel_err": rel(maps, maps0)}
    print(mode, out[mode], flush=True)
set_mode("full")
print(json.dumps(out))
