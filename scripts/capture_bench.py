#!/usr/bin/env python
"""Accuracy + timing of the attn-store / fused capture+collect forward: tcgen05 kernel (skp_capture_tc.cu) vs the SIMT row
kernels vs the torch formulation.  CUDA events on the launching stream, 512 MiB L2 flush between timed launches.
    python scripts/capture_bench.py [--json out.json]"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stablekeypoints_b200 import ops  # noqa: E402
from stablekeypoints_b200._lib import lib  # noqa: E402

PEAK = 6546.6
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    PEAK = json.load(open(p)).get("hbm_gbs", PEAK)


def ref_probs(logits, res):
    h, s2, n = logits.shape
    s = int(s2 ** 0.5)
    up = F.interpolate(logits.reshape(h, s, s, n).permute(0, 3, 1, 2), size=(res, res), mode="bicubic", align_corners=False)
    return torch.softmax(up.permute(0, 2, 3, 1).reshape(h, res * res, n), dim=-1)


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def timeit(fn, flush, reps=6, warm=3):
    ts = []
    for i in range(reps + warm):
        flush.fill_(float(i))
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(s.elapsed_time(e))
    return sum(ts) / len(ts)


def main():
    dev = torch.device("cuda")
    flush = torch.empty(512 * 1024 * 1024 // 4, device=dev)
    out = []
    cases = [("sd15 C=1280 layer", 8, 16, 77, 128), ("sd15 C=640 layer", 8, 32, 77, 128), ("notebook N=100", 8, 16, 100, 128),
             ("cfg5 SDXL-shaped", 20, 32, 77, 256), ("small", 4, 8, 12, 64), ("R=40 odd", 2, 16, 13, 40)]
    for name, h, s, n, r in cases:
        g = torch.Generator().manual_seed(h + s + n)
        lg = (torch.randn(h, s * s, n, generator=g) * 3).to(dev)
        want = ref_probs(lg.double().cpu(), r) if h * r * r * n < 3e7 else None
        row = {"case": name, "heads": h, "s": s, "N": n, "R": r, "store_bytes": h * r * r * n * 4}
        for tc in (1, 0):
            lib().skp_capture_tc(2 if tc else 0)
            probs = ops.capture_store(lg, r)
            torch.cuda.synchronize()
            key = "tc" if tc else "simt"
            if want is not None:
                row[f"{key}_rel_err"] = rel(probs.cpu(), want)
            else:
                row[f"{key}_rowsum_err"] = float((probs.sum(-1) - 1).abs().max())
            ms = timeit(lambda: ops.capture_store(lg, r), flush)
            row[f"{key}_us"] = round(ms * 1e3, 2)
            row[f"{key}_GBps"] = round((row["store_bytes"] + lg.numel() * 4) / (ms * 1e-3) / 1e9, 1)
            row[f"{key}_frac_hbm"] = round(row[f"{key}_GBps"] / PEAK, 4)
            if tc == 1:
                keep = probs
            else:
                row["tc_vs_simt"] = rel(keep, probs)
            del probs
        out.append(row)
        print(json.dumps(row), flush=True)
    # fused capture+collect (training path): 4 layers -> maps[N, R, R]
    for n in (77, 100):
        g = torch.Generator().manual_seed(n)
        lgs = [(torch.randn(8, s * s, n, generator=g) * 3).to(dev) for s in (16, 16, 16, 32)]
        want = torch.stack([ref_probs(l.double().cpu(), 128) for l in lgs]).mean(dim=(0, 1)).t().reshape(n, 128, 128)
        row = {"case": f"capture_mean 4 layers N={n}"}
        for tc in (1, 0):
            lib().skp_capture_tc(2 if tc else 0)
            key = "tc" if tc else "simt"
            m = ops.capture_mean(lgs, 128)
            row[f"{key}_rel_err"] = rel(m.cpu(), want)
            row[f"{key}_us"] = round(timeit(lambda: ops.capture_mean(lgs, 128), flush) * 1e3, 2)
        out.append(row)
        print(json.dumps(row), flush=True)
    lib().skp_capture_tc(1)
    if "--json" in sys.argv:
        json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
