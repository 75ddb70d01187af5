#!/usr/bin/env python
"""Per-kernel timing of libskp_b200 at the SD1.5 / 512^2 shapes of the hot path, with CUDA events, L2 flushed between
launches, against the measured roofline (MEASURED_PEAKS.json).  Also the target command for ncu captures.

    python scripts/kernel_bench.py [--tokens 77] [--only NAME] [--reps 10] [--json out.json]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from stablekeypoints_b200 import ops  # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    d = json.load(open(p)) if os.path.exists(p) else {}
    return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0)


def timeit(fn, reps, flush):
    for _ in range(3):
        fn()
    ts = []
    for i in range(reps):
        if flush is not None:
            flush.fill_(float(i))
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=77)
    ap.add_argument("--res", type=int, default=128)
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--json", default="")
    a = ap.parse_args()
    dev = torch.device("cuda")
    hbm, tf = peaks()
    n, r, h = a.tokens, a.res, 8
    flush = torch.empty(512 * 1024 * 1024 // 4, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    rows = []

    def add(name, shape, fn, bytes_=None, flops=None):
        if a.only and a.only not in name:
            return
        med, best = timeit(fn, a.reps, flush)
        row = {"kernel": name, "shape": shape, "ms_median": round(med, 5), "ms_best": round(best, 5)}
        if bytes_ is not None:
            row.update(bound="hbm", algorithmic_bytes=bytes_, achieved_gbs=round(bytes_ / med / 1e6, 1), frac=round(bytes_ / med / 1e6 / hbm, 4))
        if flops is not None:
            row.update(bound="tensor", algorithmic_flops=flops, achieved_tflops=round(flops / med / 1e9, 2), frac=round(flops / med / 1e9 / tf, 4))
        rows.append(row)
        print(json.dumps(row), flush=True)

    # ---- attn-store (capture) kernels
    for s in (16, 32):
        lg = torch.randn(h, s * s, n, device=dev, generator=g) * 3
        store_bytes = h * r * r * n * 4 + lg.numel() * 4
        add("capture_store_fwd", f"h8 s{s} N{n} R{r}", lambda lg=lg: ops.capture_store(lg, r), bytes_=store_bytes)
        dp = torch.randn(h, r * r, n, device=dev, generator=g)
        lgr = lg.clone().requires_grad_(True)
        pr = ops.capture_store(lgr, r)
        add("capture_store_bwd", f"h8 s{s} N{n} R{r}", lambda pr=pr, dp=dp: torch.autograd.grad(pr, lgr, dp, retain_graph=True),
            bytes_=store_bytes + 2 * lg.numel() * 4)
    # BASELINE cfg5: SDXL-shaped capture (20 heads, 32x32 layer, R=256): a 404 MB store per layer, larger than L2
    lg5 = torch.randn(20, 32 * 32, n, device=dev, generator=g) * 3
    add("capture_store_fwd (cfg5: SDXL-shaped)", f"h20 s32 N{n} R256", lambda: ops.capture_store(lg5, 256),
        bytes_=20 * 256 * 256 * n * 4 + lg5.numel() * 4)
    del lg5
    lgs = [torch.randn(h, s * s, n, device=dev, generator=g) * 3 for s in (16, 16, 16, 32)]
    mean_bytes = sum(l.numel() for l in lgs) * 4 + n * r * r * 4
    ops.CAPTURE_MEAN_FWD = "fused"
    add("capture_mean_fwd(fused tile kernel)", f"4 layers N{n} R{r}", lambda: ops.capture_mean(lgs, r), bytes_=mean_bytes)
    ops.CAPTURE_MEAN_FWD = "store"
    add("capture_mean_fwd(store+collect, default)", f"4 layers N{n} R{r}", lambda: ops.capture_mean(lgs, r), bytes_=mean_bytes)
    lgr = [l.clone().requires_grad_(True) for l in lgs]
    mp = ops.capture_mean(lgr, r)
    dm = torch.randn_like(mp)
    add("capture_mean_bwd", f"4 layers N{n} R{r}", lambda: torch.autograd.grad(mp, lgr, dm, retain_graph=True),
        bytes_=2 * mean_bytes)
    # ---- collect_maps
    st = [torch.rand(h, r * r, n, device=dev, generator=g) for _ in range(4)]
    add("collect_maps_fwd(train)", f"4x[8,{r*r},{n}]", lambda: ops.collect_maps_op(st, -1, None),
        bytes_=4 * h * r * r * n * 4 + n * r * r * 4)
    idx = torch.arange(10, device=dev)
    add("collect_maps_fwd(eval K=10 ->512)", f"4x[8,{r*r},{n}]", lambda: ops.collect_maps_op(st, 512, idx),
        bytes_=4 * h * r * r * 10 * 4 + 10 * 512 * 512 * 4)
    str_ = [s_.clone().requires_grad_(True) for s_ in st]
    cm = ops.collect_maps_op(str_, -1, None)
    dcm = torch.randn_like(cm)
    add("collect_maps_bwd(train)", f"4x[8,{r*r},{n}]", lambda: torch.autograd.grad(cm, str_, dcm, retain_graph=True),
        bytes_=4 * h * r * r * n * 4 + n * r * r * 4)
    # ---- projections (tcgen05 split-bf16)
    for (m, nn, k, what) in [(4096, 320, 320, "to_q/to_out 64^2"), (1024, 640, 640, "to_q/to_out 32^2"), (256, 1280, 1280, "to_q/to_out 16^2"),
                             (n, 24960, 768, "K|V all layers"), (n, 768, 24960, "d(context)"), (r * r, 1280, 1280, "literal q' (R^2 rows)")]:
        x = torch.randn(m, k, device=dev, generator=g)
        fw = ops.FrozenWeight(torch.randn(nn, k, device=dev, generator=g) / k ** 0.5, need_dgrad=False)
        add("gemm_nt_tc(+split)", f"{m}x{nn}x{k} {what}", lambda x=x, fw=fw: ops.frozen_linear(x, fw), flops=2.0 * m * nn * k)
        a_hi, a_lo = ops.split_bf16(x)
        from stablekeypoints_b200._lib import check, lib, ptr, stream
        out = torch.empty(m, nn, device=dev)

        def tc_only(a_hi=a_hi, a_lo=a_lo, fw=fw, out=out, m=m, nn=nn):
            ops.gemm_nt_presplit(a_hi, a_lo, m, fw.w_split, nn, out=out)
        add("gemm_nt_tc(kernel only)", f"{m}x{nn}x{k} {what}", tc_only, flops=2.0 * m * nn * k)
    # ---- attention core
    for (s, c) in [(4096, 320), (1024, 640), (256, 1280), (64, 1280)]:
        q = torch.randn(s, c, device=dev, generator=g).requires_grad_(True)
        kv = torch.randn(n, 2 * c, device=dev, generator=g).requires_grad_(True)
        d = c // h
        add("cross_attn_fwd", f"S{s} C{c} N{n}", lambda q=q, kv=kv, c=c, d=d: ops.cross_attn_core(q, kv[:, :c], kv[:, c:], h, d ** -0.5),
            flops=4.0 * s * n * c)
        o, lgt = ops.cross_attn_core(q, kv[:, :c], kv[:, c:], h, d ** -0.5, True)
        do = torch.randn_like(o)
        add("cross_attn_bwd", f"S{s} C{c} N{n}", lambda o=o, q=q, kv=kv, do=do: torch.autograd.grad(o, (q, kv), do, retain_graph=True),
            flops=8.0 * s * n * c)
    # ---- self-attention core (split-bf16 flash kernels; algorithmic FLOPs 4*S^2*C fwd, 14*S^2*C bwd (7 contractions))
    import torch.nn.functional as F
    for (s, c) in [(4096, 320), (1024, 640), (256, 1280), (64, 1280)]:
        d = c // h
        qkv = torch.randn(s, 3 * c, device=dev, generator=g).requires_grad_(True)
        add("self_attn_fwd", f"S{s} C{c} d{d}", lambda qkv=qkv, d=d: ops.self_attn_core(qkv.detach(), h, d ** -0.5), flops=4.0 * s * s * c)
        o = ops.self_attn_core(qkv, h, d ** -0.5)
        do = torch.randn_like(o)
        add("self_attn_bwd", f"S{s} C{c} d{d}", lambda o=o, qkv=qkv, do=do: torch.autograd.grad(o, qkv, do, retain_graph=True),
            flops=14.0 * s * s * c)
        q4 = qkv.detach().reshape(s, 3, h, d).permute(1, 2, 0, 3)
        qq, kk, vv = (q4[i][None].contiguous().requires_grad_(True) for i in range(3))
        add("torch_sdpa_fp32_fwd (library, for comparison)", f"S{s} C{c} d{d}",
            lambda qq=qq, kk=kk, vv=vv: F.scaled_dot_product_attention(qq.detach(), kk.detach(), vv.detach()), flops=4.0 * s * s * c)
        o2 = F.scaled_dot_product_attention(qq, kk, vv)
        do2 = torch.randn_like(o2)
        add("torch_sdpa_fp32_bwd (library, for comparison)", f"S{s} C{c} d{d}",
            lambda o2=o2, qq=qq, kk=kk, vv=vv, do2=do2: torch.autograd.grad(o2, (qq, kk, vv), do2, retain_graph=True), flops=10.0 * s * s * c)
    # ---- losses / selection
    maps = torch.rand(n, r, r, device=dev, generator=g) ** 6
    maps_t = torch.rand(n, r, r, device=dev, generator=g) ** 6
    from stablekeypoints_b200 import ptp_utils
    add("find_top_k_gaussian", f"[{n},{r},{r}]", lambda: ptp_utils.find_top_k_gaussian(maps, 25, sigma=2.0), bytes_=n * r * r * 4)
    cand = ptp_utils.find_top_k_gaussian(maps, 25, sigma=2.0)
    add("furthest_point_sampling", "25 -> 10", lambda: ptp_utils.furthest_point_sampling(maps_t, 10, cand), bytes_=n * r * r * 4)
    sel = ptp_utils.furthest_point_sampling(maps_t, 10, cand)
    th = torch.tensor([[0.9, 0.1, 0.05], [-0.1, 0.9, -0.02]])
    add("sharpen_loss_fwd", "K10", lambda: ops.sharpen_loss_op(maps, sel, 2.0), bytes_=10 * r * r * 4 * 2)
    add("equivariance_loss_fwd", "K10", lambda: ops.equivariance_loss_op(maps, maps_t, sel, th), bytes_=10 * r * r * 4 * 2)
    img = torch.rand(1, 3, 512, 512, device=dev, generator=g)
    add("affine_warp 512^2", "[1,3,512,512]", lambda: ops.affine_warp(img, th[None]), bytes_=2 * img.numel() * 4)
    hm = torch.rand(10, 512, 512, device=dev, generator=g)
    add("soft_argmax", "[10,512,512]", lambda: ops.soft_argmax_(hm), bytes_=hm.numel() * 4 * 2)
    if a.json:
        json.dump(rows, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
