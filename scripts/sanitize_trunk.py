#!/usr/bin/env python
"""Exercise the trunk kernels (tcgen05 GEMM incl. split-K, implicit-GEMM 3x3 convolution, GroupNorm one-launch cluster and
two-kernel paths, LayerNorm / GEGLU projections, self-attention on tcgen05 forward + backward and on mma.sync,
cross-attention on mma.sync and tcgen05, selection / loss kernels) at small shapes -- meant to run under compute-sanitizer:

    compute-sanitizer --tool memcheck python scripts/sanitize_trunk.py

The cases ARE the parity tests of tests/test_gpu_kernels.py (same bodies, called with small parameters), so a clean
sanitizer log comes with a numerical verdict against the fp64 / oracle references.  Prints one line per case and
SANITIZE_TRUNK_OK at the end."""
import os
import sys
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stablekeypoints_b200 import ops  # noqa: E402
from tests import test_gpu_kernels as T  # noqa: E402


def main():
    cases = []
    for m, n, k in ((256, 1280, 1280), (300, 200, 100), (77, 1536, 768), (16, 32, 48)):
        cases.append(("gemm_nt tc %dx%dx%d" % (m, n, k), lambda mp, a=(m, n, k): T.test_gemm_nt(ops, "tc", 3e-5, *a)))
    cases.append(("gemm dgrad", lambda mp: T.test_gemm_tc_dgrad_matches_autograd(ops)))
    for geo in ((32, 32, 128, 128, 1, 1), (16, 16, 320, 320, 1, 1), (32, 32, 128, 128, 2, 1), (17, 23, 12, 20, 1, 1), (32, 32, 128, 128, 2, 0)):
        cases.append(("conv3x3 %s" % (geo,), lambda mp, a=geo: T.test_frozen_conv3x3_fwd_bwd_vs_torch(ops, *a)))
    for geo in ((16, 16, 320, 320, 32, True), (32, 32, 128, 64, 32, False), (9, 13, 24, 20, 4, True), (64, 64, 128, 64, 32, True)):
        cases.append(("groupnorm %s" % (geo,), lambda mp, a=geo: T.test_groupnorm_fused_ops_vs_torch(ops, *a)))
    cases.append(("ln_linear", lambda mp: T.test_ln_linear_fwd_bwd(ops, 256, 320, 960)))
    cases.append(("ln_linear ragged", lambda mp: T.test_ln_linear_fwd_bwd(ops, 100, 64, 40)))
    cases.append(("geglu_linear", lambda mp: T.test_geglu_linear_fwd_bwd(ops, 256, 640, 320)))
    cases.append(("geglu_linear ragged", lambda mp: T.test_geglu_linear_fwd_bwd(ops, 100, 64, 24)))
    for s, h, d in ((256, 8, 40), (128, 4, 80), (384, 2, 48), (256, 2, 96), (256, 3, 64)):
        cases.append(("self-attn tcgen05 fwd+bwd S=%d h=%d d=%d" % (s, h, d),
                      lambda mp, a=(s, h, d): T.test_self_attn_fwd_bwd(ops, *a, "tcgen05", mp)))
    for s, h, d in ((64, 8, 160), (100, 2, 16), (256, 2, 160), (77, 2, 48)):
        cases.append(("self-attn mma.sync S=%d h=%d d=%d" % (s, h, d),
                      lambda mp, a=(s, h, d): T.test_self_attn_fwd_bwd(ops, *a, "mma", mp)))
    for impl, tol in (("tc", 1e-4), ("tcgen05", 1e-4)):
        for s, n, h, d in ((256, 77, 8, 160), (64, 500, 8, 40), (100, 33, 2, 24), (256, 100, 4, 80)):
            cases.append(("cross-attn %s S=%d N=%d h=%d d=%d" % (impl, s, n, h, d),
                          lambda mp, a=(s, n, h, d), i=impl, t=tol: T.test_cross_attn_fwd_bwd(ops, mp, *a, i, t)))
    cases.append(("dense attention", lambda mp: T.test_dense_attention_vs_float64(ops, 256, 32)))
    cases.append(("collect_maps golden", lambda mp: T.test_collect_maps_golden(ops)))
    cases.append(("selection golden", lambda mp: T.test_selection_golden(ops)))
    cases.append(("entropy sort", lambda mp: T.test_entropy_sort_vs_oracle(ops)))
    cases.append(("losses golden", lambda mp: T.test_losses_golden(ops)))
    cases.append(("affine warp", lambda mp: T.test_affine_warp_and_rng_order(ops)))
    cases.append(("adam", lambda mp: T.test_adam_step(ops)))
    only = os.environ.get("SAN_ONLY")
    for name, fn in cases:
        if only and only not in name:
            continue
        t0 = time.time()
        with pytest.MonkeyPatch.context() as mp:
            fn(mp)
        print("case ok: %s (%.1f s)" % (name, time.time() - t0), flush=True)
    print("SANITIZE_TRUNK_OK %d cases" % len(cases))


if __name__ == "__main__":
    main()
