#!/usr/bin/env python
"""Times the attn-store kernel on the out-of-L2 shapes (BASELINE cfg5: 20 heads, 32x32 -> R=256, N=77; the CLI default N=500
at SD1.5 sizes), L2 flushed between launches; meant to be run plain (timing) and under `ncu -k regex:capture_store` (traffic).
    python scripts/attn_store_probe.py [--impl 0|2] [--reps 5] [--cases cfg5,n500,sd15]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stablekeypoints_b200 import ops  # noqa: E402
from stablekeypoints_b200._lib import lib  # noqa: E402

CASES = {"cfg5": (20, 32, 77, 256), "n500": (8, 16, 500, 128), "n500s32": (8, 32, 500, 128), "sd15": (8, 16, 77, 128),
         "n100": (8, 16, 100, 128)}


def arg(name, default):
    return sys.argv[sys.argv.index(name) + 1] if name in sys.argv else default


def main():
    dev = torch.device("cuda")
    impl = int(arg("--impl", "0"))
    reps = int(arg("--reps", "5"))
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    flush = torch.empty(512 * 1024 * 1024 // 4, device=dev)
    lib().skp_capture_tc(impl)
    for name in arg("--cases", "cfg5,n500,sd15").split(","):
        h, s, n, r = CASES[name]
        lg = torch.randn(h, s * s, n, device=dev) * 3
        algo = h * r * r * n * 4 + lg.numel() * 4
        ts = []
        for i in range(reps + 2):
            flush.fill_(float(i))
            st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            st.record()
            p = ops.capture_store(lg, r)
            en.record()
            torch.cuda.synchronize()
            if i >= 2:
                ts.append(st.elapsed_time(en))
        ms = sum(ts) / len(ts)
        rs = float((p.sum(-1) - 1).abs().max())
        print(json.dumps({"case": name, "impl": impl, "h": h, "s": s, "N": n, "R": r, "us": round(ms * 1e3, 2), "min_us": round(min(ts) * 1e3, 2),
                          "GBps": round(algo / ms / 1e6, 1), "frac_hbm": round(algo / ms / 1e6 / peak, 4), "rowsum_err": rs}), flush=True)
        del p, lg


if __name__ == "__main__":
    main()
