#!/usr/bin/env python
"""One eager Stage-1 optimizer step (cfg2, N=77, R=128) for profilers.
  python scripts/profile_step.py --table out.json      # torch.profiler (CUPTI) per-kernel GPU time of one step
  ncu --nvtx --nvtx-include "skp_step" --metrics gpu__time_duration.sum --csv ... python scripts/profile_step.py
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from stablekeypoints_b200 import optimize, optimize_token
from stablekeypoints_b200.invertable_transform import RandomAffineWithInverse

ap = argparse.ArgumentParser()
ap.add_argument("--table", default="")
ap.add_argument("--tokens", type=int, default=77)
ap.add_argument("--early-exit", action="store_true")
ap.add_argument("--shapes", default="", help="write a per-(entry point, problem shape) table of one eager step (CUDA events)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
ldm, controllers, _ = optimize_token.load_ldm("cuda:0", "synthetic:0", feature_upsample_res=128, attn_gain=4.0, precision="fp32")
ldm.unet.early_exit = a.early_exit
args = bench.stage1_args(a.tokens)
ctx = torch.randn(1, a.tokens, 768, device=dev).requires_grad_(True)
opt = optimize.EmbeddingOptimizer(ctx, lr=5e-3, capturable=True)
tr = RandomAffineWithInverse(degrees=15, scale=(0.8, 1.0), translate=(0.25, 0.25))
img = optimize.SyntheticKeypointDataset(length=2)[0]["img"][None].to(dev)

def step():
    out = optimize.stage1_iteration(ldm, controllers, img, ctx, tr, args)
    opt.step(); opt.zero_grad()
    return out

for _ in range(3):
    step()
torch.cuda.synchronize()
rid = torch.cuda.nvtx.range_start("skp_step")   # start/end range: process-wide (backward runs on autograd's own thread)
step()
torch.cuda.synchronize()
torch.cuda.nvtx.range_end(rid)
if a.shapes:
    from stablekeypoints_b200 import _lib
    _lib.start_profile()
    step()
    prof = _lib.stop_profile(by_shape=True)
    rows = sorted(({"call": k, "calls": len(v), "ms": round(sum(v), 4), "us_each": round(1e3 * sum(v) / len(v), 2)} for k, v in prof.items()),
                  key=lambda r: -r["ms"])
    json.dump(rows, open(a.shapes, "w"), indent=1)
    for r in rows[:60]:
        print(f"{r['ms']:9.3f} ms  x{r['calls']:4d}  {r['us_each']:9.2f} us  {r['call']}")
if a.table:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0)
        if t > 0:
            rows.append({"kernel": e.key[:120], "calls": e.count, "us": round(t, 1)})
    rows.sort(key=lambda r: -r["us"])
    total = sum(r["us"] for r in rows)
    json.dump({"total_gpu_us": round(total, 1), "kernels": rows[:60]}, open(a.table, "w"), indent=1)
    for r in rows[:32]:
        print(f"{r['us']:10.0f} us {100 * r['us'] / total:5.1f}%  x{r['calls']:5d}  {r['kernel'][:100]}")
    print("total GPU us", round(total))
