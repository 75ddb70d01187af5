#!/usr/bin/env python
"""Times the self-attention backward (tcgen05 kernel of skp_attn_tc_bwd.cu vs the mma.sync kernels) on the step's shapes.
    python scripts/attn_bwd_bench.py [--reps 20]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stablekeypoints_b200 import ops  # noqa: E402


def main():
    reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 20
    dev = torch.device("cuda")
    for (s, heads, d) in [(4096, 8, 40), (1024, 8, 80), (1024, 8, 64), (1024, 4, 32)]:
        c = heads * d
        if d <= 64:      # forward on tcgen05 (skp_attn_tc.cu): CUDA events around the op (operand split + kernel)
            x = torch.randn(s, 3 * c, device=dev) * 1.5
            ts = []
            for i in range(reps + 3):
                st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                st.record()
                ops.self_attn_core(x, heads, d ** -0.5)
                en.record()
                torch.cuda.synchronize()
                if i >= 3:
                    ts.append(st.elapsed_time(en))
            print(json.dumps({"S": s, "heads": heads, "d": d, "forward": "tcgen05", "us": round(sorted(ts)[len(ts) // 2] * 1e3, 1)}), flush=True)
        for mode in ("tcgen05", "mma"):
            ops.SELF_ATTN_TC_BWD = mode == "tcgen05"
            x = (torch.randn(s, 3 * c, device=dev) * 1.5).requires_grad_(True)
            do = torch.randn(s, c, device=dev)
            o = ops.self_attn_core(x, heads, d ** -0.5)
            ts = []
            for i in range(reps + 3):
                x.grad = None
                st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                st.record()
                o.backward(do, retain_graph=True)
                en.record()
                torch.cuda.synchronize()
                if i >= 3:
                    ts.append(st.elapsed_time(en))
            us = sorted(ts)[len(ts) // 2] * 1e3
            flops = 14.0 * s * s * c if mode == "mma" else 16.0 * s * s * c   # 7 (two-kernel mma) / 8 (recompute) contractions
            print(json.dumps({"S": s, "heads": heads, "d": d, "backward": mode, "us": round(us, 1),
                              "algorithmic_TFLOPs": round(10.0 * s * s * c / us / 1e6, 1), "executed_TFLOPs": round(flops / us / 1e6, 1)}), flush=True)


if __name__ == "__main__":
    main()
