#!/usr/bin/env python
"""Exercise the attn-store / fused capture+collect kernels (forward and backward; tcgen05, SIMT row and SIMT tile kernels) at
N in {77, 100, 500} -- meant to be run under compute-sanitizer:

    PYTORCH_NO_CUDA_MEMORY_CACHING=1 compute-sanitizer --tool memcheck  python scripts/sanitize_capture.py
    compute-sanitizer --tool racecheck python scripts/sanitize_capture.py
    compute-sanitizer --tool initcheck python scripts/sanitize_capture.py

Every result is also checked against the torch formulation (bicubic of the logits, softmax over tokens), so a clean
sanitizer log comes with a numerical verdict.  Prints one line per case and SANITIZE_CASES_OK at the end."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stablekeypoints_b200 import ops  # noqa: E402
from stablekeypoints_b200._lib import lib  # noqa: E402


def ref_probs(logits, res):
    h, s2, n = logits.shape
    s = int(s2 ** 0.5)
    up = F.interpolate(logits.reshape(h, s, s, n).permute(0, 3, 1, 2), size=(res, res), mode="bicubic", align_corners=False)
    return torch.softmax(up.permute(0, 2, 3, 1).reshape(h, res * res, n), dim=-1)


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def main():
    heads = int(os.environ.get("SAN_HEADS", "2"))
    res = int(os.environ.get("SAN_RES", "64"))
    worst = 0.0
    rows = tuple(int(r) for r in os.environ.get("SAN_ROWS", "3,2,1,0").split(","))
    for row in rows:   # 3: tcgen05 kernels forced where eligible; 2: register store kernel + row backward; 1: SIMT row kernels; 0: tile kernels
        lib().skp_capture_tc(2 if row == 3 else 0)
        lib().skp_capture_select(min(row, 2), 1 if row else 0)
        ops.CAPTURE_MEAN_FWD = "store" if row else "fused"
        for n in (77, 100, 500):
            for scale, shift in ((3.0, 0.0), (3.0, -60.0), (40.0, 0.0)):
                g = torch.Generator().manual_seed(n + row)
                logits = [torch.randn(heads, s * s, n, generator=g) * scale + shift for s in (16, 32)]
                dm = torch.randn(n, res, res, generator=g)
                dp = torch.randn(heads, res * res, n, generator=g)
                lr = [l.clone().requires_grad_(True) for l in logits]
                stack = torch.stack([ref_probs(l, res) for l in lr])
                mref = stack.mean(dim=(0, 1)).t().reshape(n, res, res)
                ((mref * dm).sum() + (stack[0] * dp).sum()).backward()
                lc = [l.cuda().requires_grad_(True) for l in logits]
                m = ops.capture_mean(lc, res)
                p = ops.capture_store(lc[0], res)
                ((m * dm.cuda()).sum() + (p * dp.cuda()).sum()).backward()
                torch.cuda.synchronize()
                errs = [rel(m.detach().cpu(), mref.detach()), rel(p.detach().cpu(), stack[0].detach())] + \
                       [rel(a.grad.cpu(), b.grad) for a, b in zip(lc, lr)]
                finite = all(bool(torch.isfinite(t).all()) for t in (m, p, lc[0].grad, lc[1].grad))
                worst = max(worst, max(errs))
                print(f"row={row} N={n} scale={scale} shift={shift}: finite={finite} errs=" + " ".join(f"{e:.1e}" for e in errs), flush=True)
                assert finite and max(errs) < 1e-3, errs
    lib().skp_capture_select(2, 1)
    lib().skp_capture_tc(1)
    print(f"SANITIZE_CASES_OK worst_rel_err={worst:.2e}")


if __name__ == "__main__":
    main()
