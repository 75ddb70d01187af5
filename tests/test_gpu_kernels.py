"""-m gpu: every C-ABI kernel of libskp_b200 against the CPU oracle / torch-CPU autograd on the same seeded inputs,
and against the committed golden fixtures minted by the reference's own code."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import hotpath as hp
from tests._util import load_golden, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from stablekeypoints_b200 import ops as _ops
    return _ops


def cu(t):
    return torch.as_tensor(t).cuda()


# ----------------------------------------------------------------------------- GEMMs
GEMM_SHAPES = [(77, 24960, 768), (256, 1280, 1280), (4096, 320, 320), (64, 1280, 1280), (1024, 640, 640),
               (77, 768, 24960), (16, 32, 48), (300, 200, 100)]


@pytest.mark.parametrize("impl,tol", [("simt", 2e-6), ("tc", 3e-5)])
@pytest.mark.parametrize("m,n,k", GEMM_SHAPES)
def test_gemm_nt(ops, impl, tol, m, n, k):
    g = torch.Generator().manual_seed(m * 7 + n * 3 + k)
    a = torch.randn(m, k, generator=g)
    b = torch.randn(n, k, generator=g) / k ** 0.5
    bias = torch.randn(n, generator=g)
    res = torch.randn(m, n, generator=g)
    want = (a.double() @ b.double().t() + bias.double() + res.double())
    ops.set_gemm_impl(impl)
    try:
        bw = ops.FrozenWeight(cu(b))
        got = ops.frozen_linear(cu(a), bw, cu(bias), residual=cu(res))
        plain = ops.frozen_linear(cu(a), bw)
    finally:
        ops.set_gemm_impl("tc")
    tol = tol * max(1.0, (k / 1024) ** 0.5)   # fp32 accumulation error grows ~sqrt(K)
    assert rel_err(got.cpu(), want) < tol
    assert rel_err(plain.cpu(), a.double() @ b.double().t()) < tol


def test_gemm_tc_dgrad_matches_autograd(ops):
    g = torch.Generator().manual_seed(0)
    a = torch.randn(200, 320, generator=g)
    w = torch.randn(640, 320, generator=g) / 18
    dy = torch.randn(200, 640, generator=g)
    x = cu(a).requires_grad_(True)
    y = ops.frozen_linear(x, ops.FrozenWeight(cu(w)))
    y.backward(cu(dy))
    assert rel_err(x.grad.cpu(), dy.double() @ w.double()) < 3e-5


@pytest.mark.parametrize("h,w,cin,cout,stride,pad", [(16, 16, 1280, 1280, 1, 1), (64, 64, 320, 320, 1, 1), (8, 8, 2560, 1280, 1, 1),
                                                      (64, 64, 4, 320, 1, 1), (32, 32, 640, 640, 2, 1), (16, 16, 320, 4, 1, 1),
                                                      (17, 23, 12, 20, 1, 1), (32, 32, 128, 128, 2, 0)])
def test_frozen_conv3x3_fwd_bwd_vs_torch(ops, h, w, cin, cout, stride, pad):
    """im2col-to-split-bf16 + tcgen05 GEMM == F.conv2d (fp64 reference), forward and input gradient."""
    g = torch.Generator().manual_seed(h + cin + cout)
    x = torch.randn(1, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5
    b = torch.randn(cout, generator=g)
    xr = x.double().requires_grad_(True)
    xin = F.pad(xr, (0, 1, 0, 1)) if pad == 0 else xr            # the VAE down-sampler pads bottom/right only
    yr = F.conv2d(xin, wt.double(), b.double(), stride=stride, padding=pad)
    dy = torch.randn(yr.shape, generator=g, dtype=torch.float64)
    yr.backward(dy)
    ho, wo = yr.shape[2], yr.shape[3]
    xc = cu(x[0].permute(1, 2, 0).reshape(h * w, cin).contiguous()).requires_grad_(True)
    y, ho2, wo2 = ops.frozen_conv3x3(xc, h, w, ops.FrozenConv3x3(cu(wt)), cu(b), stride=stride, pad=pad, out_hw=(ho, wo))
    assert (ho2, wo2) == (ho, wo)
    y.backward(cu(dy[0].permute(1, 2, 0).reshape(ho * wo, cout).float().contiguous()))
    tol = 3e-5 * max(1.0, (9 * cin / 1024) ** 0.5)
    assert rel_err(y.detach().cpu(), yr.detach()[0].permute(1, 2, 0).reshape(ho * wo, cout)) < tol
    assert rel_err(xc.grad.cpu(), xr.grad[0].permute(1, 2, 0).reshape(h * w, cin)) < 10 * tol


@pytest.mark.parametrize("h,w,c,cout,groups,silu", [(16, 16, 1280, 640, 32, True), (64, 64, 320, 320, 32, True), (8, 8, 2560, 1280, 32, True),
                                                     (32, 32, 128, 64, 32, False), (9, 13, 24, 20, 4, True)])
def test_groupnorm_fused_ops_vs_torch(ops, h, w, c, cout, groups, silu):
    """skp_gn_* : plain GN(+SiLU), GN->3x3 conv and GN->1x1 projection, forward and input gradient, vs torch fp64."""
    g = torch.Generator().manual_seed(h * w + c)
    x = torch.randn(1, c, h, w, generator=g) * 2 + 0.5
    gamma, beta = torch.randn(c, generator=g), torch.randn(c, generator=g)
    wt3 = torch.randn(cout, c, 3, 3, generator=g) / (9 * c) ** 0.5
    wt1 = torch.randn(cout, c, generator=g) / c ** 0.5
    bias = torch.randn(cout, generator=g)
    eps = 1e-5

    def ref_act(xx):
        y = F.group_norm(xx, groups, gamma.double(), beta.double(), eps)
        return F.silu(y) if silu else y

    def cl(t4):  # [1,C,H,W] -> [HW, C]
        return t4[0].permute(1, 2, 0).reshape(h * w, -1)

    xr = x.double().requires_grad_(True)
    act = ref_act(xr)
    y3 = F.conv2d(act, wt3.double(), bias.double(), padding=1)
    y1 = F.linear(cl(act), wt1.double(), bias.double())
    d3, d1 = torch.randn(y3.shape, generator=g, dtype=torch.float64), torch.randn(y1.shape, generator=g, dtype=torch.float64)
    g3, = torch.autograd.grad(y3, xr, d3, retain_graph=True)
    g1, = torch.autograd.grad(y1, xr, d1)
    xc = cu(cl(x).contiguous())
    gm, bt = cu(gamma), cu(beta)
    assert rel_err(ops.group_norm_cl(xc, gm, bt, groups, eps, silu).cpu(), cl(act.detach())) < 1e-5
    x3 = xc.clone().requires_grad_(True)
    o3, _, _ = ops.gn_conv3x3(x3, h, w, gm, bt, groups, eps, silu, ops.FrozenConv3x3(cu(wt3)), cu(bias))
    o3.backward(cu(cl(d3).float().contiguous()))
    x1 = xc.clone().requires_grad_(True)
    o1 = ops.gn_linear(x1, gm, bt, groups, eps, silu, ops.FrozenWeight(cu(wt1)), cu(bias))
    o1.backward(cu(d1.float()))
    tol = 5e-5 * max(1.0, (9 * c / 1024) ** 0.5)
    assert rel_err(o3.detach().cpu(), cl(y3.detach())) < tol
    assert rel_err(o1.detach().cpu(), y1.detach()) < tol
    assert rel_err(x3.grad.cpu(), cl(g3)) < 10 * tol
    assert rel_err(x1.grad.cpu(), cl(g1)) < 10 * tol


@pytest.mark.parametrize("mode", [2, 1])
def test_gemm_conv_persistent_kernel(ops, mode):
    """gemm_nt_tc_persist_kernel (one CTA walks several output tiles, two accumulators in tensor memory): bit-identical to the
    one-tile-per-CTA kernel -- same MMAs in the same order per tile, same epilogue arithmetic -- and within the GEMM
    tolerance of fp64.  mode 2 forces it on every un-split problem with ~3 tiles per CTA (ragged M / N tails, non-vector N,
    bias / residual, implicit convolution with image borders); mode 1 is the rule `SKP_GEMM_PERSIST=1` applies (more tiles
    than SMs).  The kernel is opt-in (default 0): see profiles/r02_gemm_persist.md."""
    from stablekeypoints_b200._lib import lib
    g = torch.Generator().manual_seed(11 + mode)
    gemms = [(1300, 200, 128), (513, 70, 200), (4096, 640, 320), (300, 1280, 64)] if mode == 2 else [(20000, 256, 192), (4096, 2560, 320)]
    convs = [(64, 64, 64, 320), (24, 40, 128, 96), (32, 32, 192, 64)] if mode == 2 else [(128, 128, 64, 320)]
    try:
        for m, n, k in gemms:
            a, b = torch.randn(m, k, generator=g), torch.randn(n, k, generator=g) / k ** 0.5
            bias, res = torch.randn(n, generator=g), torch.randn(m, n, generator=g)
            fw = ops.FrozenWeight(cu(b))
            outs = []
            for md in (0, mode):
                lib().skp_gemm_tc_persist(md)
                outs.append((ops.frozen_linear(cu(a), fw, cu(bias), residual=cu(res)), ops.frozen_linear(cu(a), fw)))
            assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]), (m, n, k)
            assert rel_err(outs[1][0].cpu(), a.double() @ b.double().t() + bias.double() + res.double()) < 3e-5
        for h, w, cin, cout in convs:
            x = torch.randn(1, cin, h, w, generator=g)
            wt = torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5
            bias = torch.randn(cout, generator=g)
            want = F.conv2d(x.double(), wt.double(), bias.double(), padding=1)[0].permute(1, 2, 0).reshape(h * w, cout)
            xc = cu(x[0].permute(1, 2, 0).reshape(h * w, cin).contiguous())
            fcw = ops.FrozenConv3x3(cu(wt))
            outs = []
            for md in (0, mode):
                lib().skp_gemm_tc_persist(md)
                outs.append(ops.frozen_conv3x3(xc, h, w, fcw, cu(bias))[0])
            assert torch.equal(outs[0], outs[1]), (h, w, cin, cout)
            assert rel_err(outs[1].cpu(), want) < 3e-5 * max(1.0, (9 * cin / 1024) ** 0.5)
    finally:
        lib().skp_gemm_tc_persist(0)


# ----------------------------------------------------------------------------- LayerNorm / GEGLU fused projections
@pytest.mark.parametrize("rows,c,n_out", [(4096, 320, 960), (256, 1280, 1280), (64, 1280, 10240), (16, 32, 48), (100, 64, 40)])
def test_ln_linear_fwd_bwd(ops, rows, c, n_out):
    g = torch.Generator().manual_seed(rows + c)
    x = torch.randn(rows, c, generator=g) * 2 + 0.3
    gamma, beta = torch.randn(c, generator=g), torch.randn(c, generator=g)
    w = torch.randn(n_out, c, generator=g) / c ** 0.5
    b = torch.randn(n_out, generator=g)
    res = torch.randn(rows, n_out, generator=g)
    dy = torch.randn(rows, n_out, generator=g)
    xr = x.double().requires_grad_(True)
    yr = F.linear(F.layer_norm(xr, (c,), gamma.double(), beta.double()), w.double(), b.double()) + res.double()
    yr.backward(dy.double())
    xc = cu(x).requires_grad_(True)
    y = ops.ln_linear(xc, cu(gamma), cu(beta), ops.FrozenWeight(cu(w)), cu(b), residual=cu(res))
    y.backward(cu(dy))
    tol = 3e-5 * max(1.0, (c / 1024) ** 0.5)
    assert rel_err(y.detach().cpu(), yr.detach()) < tol
    assert rel_err(xc.grad.cpu(), xr.grad) < 4 * tol


@pytest.mark.parametrize("rows,h,n_out", [(4096, 1280, 320), (256, 5120, 1280), (16, 128, 32), (100, 64, 24)])
def test_geglu_linear_fwd_bwd(ops, rows, h, n_out):
    g = torch.Generator().manual_seed(rows + h)
    proj = torch.randn(rows, 2 * h, generator=g) * 1.5
    w = torch.randn(n_out, h, generator=g) / h ** 0.5
    b = torch.randn(n_out, generator=g)
    res = torch.randn(rows, n_out, generator=g)
    dy = torch.randn(rows, n_out, generator=g)
    pr = proj.double().requires_grad_(True)
    a, gate = pr.chunk(2, dim=-1)
    yr = F.linear(a * F.gelu(gate), w.double(), b.double()) + res.double()
    yr.backward(dy.double())
    pc = cu(proj).requires_grad_(True)
    y = ops.geglu_linear(pc, ops.FrozenWeight(cu(w)), cu(b), residual=cu(res))
    y.backward(cu(dy))
    tol = 3e-5 * max(1.0, (h / 1024) ** 0.5)
    assert rel_err(y.detach().cpu(), yr.detach()) < tol
    assert rel_err(pc.grad.cpu(), pr.grad) < 4 * tol


# ----------------------------------------------------------------------------- cross-attention core
def _attn_ref(q, k, v, heads, scale, extra_w=None):
    s, c = q.shape
    n = k.shape[0]
    d = c // heads
    qh = q.reshape(s, heads, d).transpose(0, 1)
    kh = k.reshape(n, heads, d).transpose(0, 1)
    vh = v.reshape(n, heads, d).transpose(0, 1)
    logits = torch.bmm(qh, kh.transpose(1, 2)) * scale
    o = torch.bmm(torch.softmax(logits, -1), vh).transpose(0, 1).reshape(s, c)
    return o, logits


@pytest.mark.parametrize("s,n,heads,d", [(256, 77, 8, 160), (1024, 77, 8, 80), (64, 500, 8, 40), (16, 12, 4, 8),
                                         (4096, 100, 8, 40), (100, 33, 2, 24)])
@pytest.mark.parametrize("impl,tol", [("simt", 2e-5), ("tc", 1e-4), ("tcgen05", 1e-4)])
def test_cross_attn_fwd_bwd(ops, monkeypatch, s, n, heads, d, impl, tol):
    # "tc": mma.sync flash kernels (the default); "tcgen05": forward on skp_xattn_tc.cu (TMEM / TMA), backward on mma.sync
    monkeypatch.setattr(ops, "XATTN_TC", impl == "tcgen05")
    impl = "tc" if impl == "tcgen05" else impl
    g = torch.Generator().manual_seed(s + n)
    c = heads * d
    q = torch.randn(s, c, generator=g, dtype=torch.float64)
    kv = torch.randn(n, 2 * c + 8, generator=g, dtype=torch.float64)   # strided K/V views like the batched projection
    do = torch.randn(s, c, generator=g, dtype=torch.float64)
    dl = torch.randn(heads, s, n, generator=g, dtype=torch.float64) * 0.1
    scale = d ** -0.5
    qr, kvr = q.clone().requires_grad_(True), kv.clone().requires_grad_(True)
    o_ref, l_ref = _attn_ref(qr, kvr[:, :c], kvr[:, c:2 * c], heads, scale)
    ((o_ref * do).sum() + (l_ref * dl).sum()).backward()
    qc, kvc = cu(q.float()).requires_grad_(True), cu(kv.float()).requires_grad_(True)
    o, logits = ops.cross_attn_core(qc, kvc[:, :c], kvc[:, c:2 * c], heads, scale, want_logits=True, impl=impl)
    ((o * cu(do.float())).sum() + (logits * cu(dl.float())).sum()).backward()
    errs = [rel_err(o.detach().cpu(), o_ref.detach()), rel_err(logits.detach().cpu(), l_ref.detach()),
            rel_err(qc.grad.cpu(), qr.grad), rel_err(kvc.grad[:, :c].cpu(), kvr.grad[:, :c]),
            rel_err(kvc.grad[:, c:2 * c].cpu(), kvr.grad[:, c:2 * c])]
    print(f"cross-attn[{impl}] S={s} N={n} h={heads} d={d}: rel err o/logits/dq/dk/dv = {errs}")
    assert max(errs[:2]) < tol / 2 and max(errs[2:]) < tol, errs
    if impl == "tc":   # the un-captured form: no logits tensor, same output and gradients
        q2, kv2 = cu(q.float()).requires_grad_(True), cu(kv.float()).requires_grad_(True)
        o2, none = ops.cross_attn_core(q2, kv2[:, :c], kv2[:, c:2 * c], heads, scale, want_logits=False, impl=impl)
        assert none.numel() == 0 and torch.equal(o2.detach(), o.detach())
        qr2, kvr2 = q.clone().requires_grad_(True), kv.clone().requires_grad_(True)
        o_ref2, _ = _attn_ref(qr2, kvr2[:, :c], kvr2[:, c:2 * c], heads, scale)
        (o_ref2 * do).sum().backward()
        (o2 * cu(do.float())).sum().backward()
        assert rel_err(q2.grad.cpu(), qr2.grad) < tol and rel_err(kv2.grad.cpu(), kvr2.grad) < tol


# ----------------------------------------------------------------------------- self-attention (attn1)
@pytest.mark.parametrize("s,heads,d", [(4096, 8, 40), (1024, 8, 80), (256, 8, 160), (64, 8, 160), (16, 4, 8), (100, 2, 16),
                                       (4, 4, 32), (1100, 3, 24), (77, 2, 48), (128, 2, 16), (256, 3, 64), (1024, 4, 32),
                                       (384, 2, 48), (256, 2, 96), (128, 4, 80), (640, 2, 72)])
@pytest.mark.parametrize("tc", ["tcgen05", "tcgen05_fwd", "mma"])
def test_self_attn_fwd_bwd(ops, s, heads, d, tc, monkeypatch):
    # tcgen05 forward + backward (eligible shapes only) / tcgen05 forward + mma.sync backward / mma.sync both
    monkeypatch.setattr(ops, "SELF_ATTN_TC", tc != "mma")
    monkeypatch.setattr(ops, "SELF_ATTN_TC_BWD", tc == "tcgen05")
    monkeypatch.setattr(ops, "SELF_ATTN_TC_MIN_S", 128)
    """Flash-style split-bf16 kernels vs float64 attention of the same packed [S, 3C] projection (fp32-grade: the
    softmax sits upstream of every captured map).  Ragged S, padded d and every tile shape are covered."""
    g = torch.Generator().manual_seed(s + heads + d)
    c = heads * d
    qkv = torch.randn(s, 3 * c, generator=g, dtype=torch.float64) * 1.5
    do = torch.randn(s, c, generator=g, dtype=torch.float64)
    scale = d ** -0.5
    ref_in = qkv.clone().requires_grad_(True)
    q, k, v = (ref_in[:, i * c:(i + 1) * c].reshape(s, heads, d).permute(1, 0, 2) for i in range(3))
    p = torch.softmax(q @ k.transpose(1, 2) * scale, dim=-1)
    o_ref = (p @ v).permute(1, 0, 2).reshape(s, c)
    (o_ref * do).sum().backward()
    x = cu(qkv.float()).requires_grad_(True)
    o = ops.self_attn_core(x, heads, scale)
    (o * cu(do.float())).sum().backward()
    errs = [rel_err(o.detach().cpu(), o_ref.detach())]
    errs += [rel_err(x.grad[:, i * c:(i + 1) * c].cpu(), ref_in.grad[:, i * c:(i + 1) * c]) for i in range(3)]   # dq, dk, dv
    print(f"self-attn[{tc}] S={s} h={heads} d={d}: rel err o/dq/dk/dv = {errs}")
    assert errs[0] < 5e-5 and max(errs[1:]) < 1e-4, errs


@pytest.mark.parametrize("s,d", [(4096, 512), (256, 32), (100, 24)])
def test_dense_attention_vs_float64(ops, s, d):
    """VAE mid-block attention (one wide head) as two split-bf16 GEMMs around the fused softmax/operand-split kernel."""
    g = torch.Generator().manual_seed(s + d)
    qkv = torch.randn(s, 3 * d, generator=g, dtype=torch.float64)
    q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
    want = torch.softmax(q @ k.t() * d ** -0.5, -1) @ v
    x = cu(qkv.float())
    got = ops.dense_attention(x[:, :d], x[:, d:2 * d], x[:, 2 * d:], d ** -0.5)
    assert rel_err(got.cpu(), want) < 5e-5


# ----------------------------------------------------------------------------- capture
def _capture_ref(logits, res):
    """softmax_tokens(bicubic_pixels(low-res logits)) -> [h, res*res, N]  (linearity form of ptp_utils.py:513-536)."""
    h, s2, n = logits.shape
    s = int(s2 ** 0.5)
    up = F.interpolate(logits.reshape(h, s, s, n).permute(0, 3, 1, 2), size=(res, res), mode="bicubic", align_corners=False)
    return torch.softmax(up.permute(0, 2, 3, 1).reshape(h, res * res, n), dim=-1)


@pytest.mark.parametrize("heads,s,n,res", [(8, 16, 77, 128), (8, 32, 77, 128), (4, 4, 12, 16), (8, 16, 500, 128),
                                            (8, 8, 100, 64), (2, 16, 20, 24), (8, 32, 16, 16), (8, 16, 100, 128),
                                            (3, 32, 77, 256), (2, 16, 13, 40)])
def test_capture_store_fwd_bwd(ops, capture_policy, heads, s, n, res):
    g = torch.Generator().manual_seed(heads + s + n)
    logits = torch.randn(heads, s * s, n, generator=g) * 3
    dp = torch.randn(heads, res * res, n, generator=g)
    lr = logits.clone().requires_grad_(True)
    pref = _capture_ref(lr, res)
    (pref * dp).sum().backward()
    lc = cu(logits).requires_grad_(True)
    p = ops.capture_store(lc, res)
    (p * cu(dp)).sum().backward()
    assert p.shape == (heads, res * res, n)
    # the tcgen05 kernel does the horizontal bicubic pass as a split-bf16 GEMM (~2e-5); the SIMT kernels are fp32 (~1e-6)
    assert rel_err(p.detach().cpu(), pref.detach()) < 5e-5
    assert torch.allclose(p.detach().sum(-1).cpu(), torch.ones(heads, res * res), atol=1e-5)
    assert rel_err(lc.grad.cpu(), lr.grad) < 5e-5


def _grad_err(a, b, floor=1e-3):
    """max|a-b| / max(max|b|, floor): saturated softmaxes have (numerically) zero gradients, where a relative error against
    max|b| ~ 1e-36 would only measure fp32 cancellation noise of size 1e-8."""
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).abs().max() / max(float(b.abs().max()), floor))


def _skewed_logits(kind, heads, s, n, g):
    """Logit distributions that trained weights produce and randn*3 never does: a large common offset,
    a wide range, one hugely negative (masked-like) token, one dominant token.  The cheap softmax bound of the row kernel
    (max |V| over the whole row) is useless on all of them."""
    base = torch.randn(heads, s * s, n, generator=g) * 3
    if kind == "offset":
        return base - 60.0
    if kind == "wide":
        return (torch.rand(heads, s * s, n, generator=g) * 2 - 1) * 80.0
    if kind == "neg_outlier":
        base[:, :, n // 3] = -3.0e4
        return base
    if kind == "dominant":
        base[:, :, 1] += 90.0
        base[:, : (s * s) // 2, 0] -= 200.0
        return base
    raise ValueError(kind)


@pytest.fixture(params=["auto", "tc"])
def capture_policy(request):
    """auto: the library's choice per shape; tc: the tcgen05 kernel wherever the shape is eligible."""
    from stablekeypoints_b200._lib import lib
    lib().skp_capture_tc(2 if request.param == "tc" else 1)
    yield request.param
    lib().skp_capture_tc(1)


@pytest.mark.parametrize("kind", ["offset", "wide", "neg_outlier", "dominant"])
@pytest.mark.parametrize("heads,s,n,res", [(8, 16, 77, 128), (2, 32, 500, 128), (3, 16, 100, 64)])
def test_capture_store_skewed_logits_fwd_bwd(ops, capture_policy, kind, heads, s, n, res):
    g = torch.Generator().manual_seed(heads * 7 + n)
    logits = _skewed_logits(kind, heads, s, n, g)
    dp = torch.randn(heads, res * res, n, generator=g)
    lr = logits.double().requires_grad_(True)
    pref = _capture_ref(lr, res)
    (pref * dp.double()).sum().backward()
    lc = cu(logits).requires_grad_(True)
    p = ops.capture_store(lc, res)
    (p * cu(dp)).sum().backward()
    assert bool(torch.isfinite(p).all()) and bool(torch.isfinite(lc.grad).all())
    # wide-range logits lose absolute precision in the fp32 bicubic itself (|x| ~ 100 -> ulp 8e-6 before the exp)
    assert rel_err(p.detach().cpu(), pref.detach()) < 2e-4, kind
    assert torch.allclose(p.detach().sum(-1).cpu(), torch.ones(heads, res * res), atol=2e-5)
    assert _grad_err(lc.grad.cpu(), lr.grad) < 5e-4, kind


@pytest.mark.parametrize("kind", ["offset", "wide", "neg_outlier", "dominant"])
def test_capture_mean_skewed_logits_fwd_bwd(ops, kind):
    heads, sides, n, res = 8, (16, 16, 16, 32), 77, 128
    g = torch.Generator().manual_seed(11)
    logits = [_skewed_logits(kind, heads, s, n, g) for s in sides]
    dm = torch.randn(n, res, res, generator=g)
    lr = [l.double().requires_grad_(True) for l in logits]
    mref = torch.stack([_capture_ref(l, res) for l in lr]).mean(dim=(0, 1)).t().reshape(n, res, res)
    (mref * dm.double()).sum().backward()
    lc = [cu(l).requires_grad_(True) for l in logits]
    m = ops.capture_mean(lc, res)
    (m * cu(dm)).sum().backward()
    assert bool(torch.isfinite(m).all())
    assert rel_err(m.detach().cpu(), mref.detach()) < 2e-4, kind
    for a, b in zip(lc, lr):
        assert bool(torch.isfinite(a.grad).all())
        assert _grad_err(a.grad.cpu(), b.grad) < 5e-4, kind


def test_capture_matches_literal_reference_formulation(ops):
    """The kernel's linearity form vs the reference's literal form (bicubic of x, second to_q) on one layer."""
    torch.manual_seed(3)
    heads, d, s, n, res, cdim = 8, 20, 16, 40, 64, 96
    from oracle import sd15
    mod = sd15.CrossAttention(heads * d, cdim, heads, d)
    x = torch.randn(1, s * s, heads * d)
    ctx = torch.randn(1, n, cdim)
    store = hp.AttentionStore()
    hp.attention_with_capture(mod, x, ctx, store, res)
    want = store.step_store["attn"][0]
    q = mod.to_q(x)[0]
    k = mod.to_k(ctx)[0]
    logits = torch.einsum("shd,nhd->hsn", q.reshape(s * s, heads, d), k.reshape(n, heads, d)) * mod.scale
    got = ops.capture_store(cu(logits.detach().contiguous()), res)
    assert rel_err(got.cpu(), want.detach()) < 5e-5


@pytest.mark.parametrize("heads,sides,n,res", [(8, (16, 16, 16, 32), 77, 128), (4, (4, 4, 4, 8), 12, 16),
                                                (8, (16, 32), 500, 128), (8, (16, 16, 16, 32), 100, 128)])
def test_capture_mean_fwd_bwd(ops, heads, sides, n, res):
    g = torch.Generator().manual_seed(n)
    logits = [torch.randn(heads, s * s, n, generator=g) * 3 for s in sides]
    dm = torch.randn(n, res, res, generator=g)
    lr = [l.clone().requires_grad_(True) for l in logits]
    stack = torch.stack([_capture_ref(l, res) for l in lr])            # [L, h, R*R, N]
    mref = stack.mean(dim=(0, 1)).t().reshape(n, res, res)
    (mref * dm).sum().backward()
    lc = [cu(l).requires_grad_(True) for l in logits]
    m = ops.capture_mean(lc, res)
    (m * cu(dm)).sum().backward()
    assert rel_err(m.detach().cpu(), mref.detach()) < 5e-5
    for a, b in zip(lc, lr):
        assert rel_err(a.grad.cpu(), b.grad) < 5e-5


@pytest.mark.parametrize("row", ["tc", "2", "1", "0"])
@pytest.mark.parametrize("n", [77, 500])
def test_capture_kernel_variants_agree(ops, monkeypatch, row, n):
    """The tcgen05 forward, the SIMT row attn-store / row backward kernels and the tile kernels they fall back to
    (skp_capture_select(0, 0); also taken when R*N is not a multiple of 4) for both modes, forward and backward."""
    from stablekeypoints_b200._lib import lib
    lib().skp_capture_tc(2 if row == "tc" else 0)          # tcgen05 forward (N <= 128) vs the SIMT kernels
    if row != "tc":
        lib().skp_capture_select(int(row), min(int(row), 1))
    monkeypatch.setattr(ops, "CAPTURE_MEAN_FWD", "fused" if row == "0" else "store")
    try:
        _variants_body(ops, row, n)
    finally:
        lib().skp_capture_select(2, 1)
        lib().skp_capture_tc(1)


def _variants_body(ops, row, n):
    g = torch.Generator().manual_seed({"tc": 2, "2": 3, "1": 1, "0": 0}[row] + 5)
    logits = [torch.randn(8, s * s, n, generator=g) * 3 for s in (16, 32)]
    dm = torch.randn(n, 128, 128, generator=g)
    lr = [l.clone().requires_grad_(True) for l in logits]
    stack = torch.stack([_capture_ref(l, 128) for l in lr])
    mref = stack.mean(dim=(0, 1)).t().reshape(n, 128, 128)
    (mref * dm).sum().backward()
    lc = [cu(l).requires_grad_(True) for l in logits]
    m = ops.capture_mean(lc, 128)
    (m * cu(dm)).sum().backward()
    assert rel_err(m.detach().cpu(), mref.detach()) < 5e-5
    for a, b in zip(lc, lr):
        assert rel_err(a.grad.cpu(), b.grad) < 5e-5
    p = ops.capture_store(cu(logits[1]), 128)
    assert bool(torch.isfinite(p).all())
    assert rel_err(p.cpu(), stack[1].detach()) < 5e-5


# ----------------------------------------------------------------------------- collect_maps
def _store(tensors):
    c = hp.AttentionStore()
    c.step_store = {"attn": list(tensors)}
    return c


def test_collect_maps_golden(ops):
    from stablekeypoints_b200 import optimize, ptp_utils
    post = load_golden("post_unet.npz")
    st = [cu(post[f"store_{i}"]) for i in range(4)]

    def run(**kw):
        c = ptp_utils.AttentionStore()
        c.step_store = {"attn": list(st)}
        out = optimize.collect_maps(c, **kw)
        assert c.step_store["attn"] == []
        return out.cpu()

    assert rel_err(run(upsample_res=-1, layers=[0, 1, 2, 3]), post["collect_train"]) < 2e-6
    assert rel_err(run(upsample_res=-1, layers=[0, 2]), post["collect_layers_02"]) < 2e-6
    assert rel_err(run(upsample_res=48, layers=[0, 1, 2, 3], indices=torch.from_numpy(post["collect_idx"])), post["collect_eval_48"]) < 2e-6
    assert rel_err(run(upsample_res=16, layers=[0, 1, 2, 3]), post["collect_same_res"]) < 2e-6


@pytest.mark.parametrize("layers,bh,r,n,idx,r2", [(4, 8, 128, 77, None, -1), (4, 8, 128, 77, [5, 0, 76, 33, 9, 1, 2, 3, 4, 10], 512),
                                                    (2, 8, 64, 500, None, -1), (4, 4, 16, 12, [3, 3, 1], 40),
                                                    (3, 8, 32, 13, None, 20)])
def test_collect_maps_fwd_bwd_vs_oracle(ops, layers, bh, r, n, idx, r2):
    g = torch.Generator().manual_seed(r + n)
    st = [torch.rand(bh, r * r, n, generator=g) for _ in range(layers)]
    ref_in = [s.clone().requires_grad_(True) for s in st]
    it = torch.tensor(idx) if idx is not None else None
    want = hp.collect_maps(_store(ref_in), r2, tuple(range(layers)), it)
    dout = torch.randn(want.shape, generator=g)
    (want * dout).sum().backward()
    cin = [cu(s).requires_grad_(True) for s in st]
    resize = r2 != -1 and (n if idx is None else len(idx)) ** 0.5 != r2
    got = ops.collect_maps_op(cin, r2 if resize else -1, cu(it) if it is not None else None)
    (got * cu(dout)).sum().backward()
    assert got.shape == want.shape
    assert rel_err(got.detach().cpu(), want.detach()) < 3e-6
    for a, b in zip(cin, ref_in):
        assert rel_err(a.grad.cpu(), b.grad) < 3e-6


# ----------------------------------------------------------------------------- selection (bit-exact)
def test_selection_golden(ops):
    from stablekeypoints_b200 import eval as skp_eval, ptp_utils
    post = load_golden("post_unet.npz")
    maps, maps_t = cu(post["maps"]), cu(post["maps_t"])
    assert np.array_equal(skp_eval.find_max_pixel(maps).cpu().numpy(), post["find_max_pixel"])
    assert np.array_equal(skp_eval.find_k_max_pixels(maps, 3).cpu().numpy(), post["find_k_max_pixels_3"])
    peaks = ops.k_argmax_flat(maps, 1)
    assert rel_err(ops.gaussian_kl_scores(maps, peaks, 2.0).cpu(), post["kl_s2.0"]) < 2e-5
    for s in (1.0, 2.0):
        cand = ptp_utils.find_top_k_gaussian(maps, 9, sigma=s)
        assert np.array_equal(cand.cpu().numpy(), post[f"topk_gaussian_s{s}"])
        fps = ptp_utils.furthest_point_sampling(maps_t, 5, cand)
        assert np.array_equal(fps.cpu().numpy(), post[f"fps_s{s}"])
    allc = ptp_utils.furthest_point_sampling(maps, 6, torch.arange(maps.shape[0]).cuda())
    assert np.array_equal(allc.cpu().numpy(), post["fps_all_candidates"])


def test_entropy_sort_vs_oracle(ops):
    """--top_k_strategy entropy (ptp_utils.py:165-187): entropies to 1e-5, the ascending order bit-exact on separated maps."""
    from stablekeypoints_b200 import ptp_utils
    g = torch.Generator().manual_seed(4)
    maps = torch.rand(77, 128, 128, generator=g) * torch.linspace(0.5, 12.0, 77).reshape(77, 1, 1)
    maps = maps[torch.randperm(77, generator=g)]
    p = torch.softmax(maps.reshape(77, -1).double(), -1)
    want = -(p * p.log()).sum(-1)
    got = ops.entropy_scores(cu(maps))
    assert rel_err(got.cpu(), want) < 1e-5
    assert np.array_equal(ptp_utils.entropy_sort(cu(maps), 25).cpu().numpy(), hp.entropy_sort(maps, 25).numpy())


@pytest.mark.parametrize("t,h,seed", [(77, 128, 0), (500, 128, 1), (10, 512, 2), (25, 33, 3)])
def test_selection_vs_oracle_full_size(ops, t, h, seed):
    from stablekeypoints_b200 import ptp_utils
    g = torch.Generator().manual_seed(seed)
    maps = torch.rand(t, h, h, generator=g) ** 8
    maps_t = torch.rand(t, h, h, generator=g) ** 8
    assert np.array_equal(ops.argmax_flat(cu(maps)).cpu().numpy(), maps.reshape(t, -1).argmax(-1).numpy())
    if h <= 128:
        ncand, k = min(25, t), min(10, t)
        cand_ref = hp.find_top_k_gaussian(maps, ncand, sigma=2.0)
        cand = ptp_utils.find_top_k_gaussian(cu(maps), ncand, sigma=2.0)
        assert np.array_equal(cand.cpu().numpy(), cand_ref.numpy())
        assert np.array_equal(ptp_utils.furthest_point_sampling(cu(maps_t), k, cand).cpu().numpy(),
                              hp.furthest_point_sampling(maps_t, k, cand_ref).numpy())


def test_fps_edge_cases_vs_oracle(ops):
    """ptp_utils.py:115-159 at its edges: fewer candidates than top_k (a shorter list comes back, as the reference's loop
    gives), coincident peaks (zero distances, strict-'>' tie-breaks), the pair only; the oracle is pinned to the reference
    on the same cases by tests/test_oracle_vs_reference.py::test_fps_edge_cases_vs_reference."""
    from stablekeypoints_b200 import ptp_utils
    from tests.test_oracle_vs_reference import _fps_edge_cases
    for maps, top_k, cand in _fps_edge_cases():
        want = hp.furthest_point_sampling(maps, top_k, cand)
        got = ptp_utils.furthest_point_sampling(cu(maps), top_k, cu(cand))
        assert np.array_equal(got.cpu().numpy(), want.numpy()), (top_k, cand.tolist(), got.tolist(), want.tolist())


# ----------------------------------------------------------------------------- losses
def test_losses_golden(ops):
    post = load_golden("post_unet.npz")
    maps = cu(post["maps"]).requires_grad_(True)
    maps_t = cu(post["maps_t"]).requires_grad_(True)
    sel = cu(post["sel"])
    from stablekeypoints_b200.invertable_transform import invert_theta
    sharp = ops.sharpen_loss_op(maps, sel, 2.0)
    equiv = ops.equivariance_loss_op(maps, maps_t, sel, invert_theta(torch.from_numpy(post["theta"]))[0])
    (100.0 * sharp + 1000.0 * equiv).backward()
    assert rel_err(sharp.detach().cpu(), post["sharp"]) < 5e-6
    assert rel_err(equiv.detach().cpu(), post["equiv"]) < 5e-6
    assert rel_err(maps.grad.cpu(), post["dmaps"]) < 1e-5
    assert rel_err(maps_t.grad.cpu(), post["dmaps_t"]) < 1e-5


@pytest.mark.parametrize("theta,gathered", [
    ([[0.85, 0.20, 0.10], [-0.20, 0.85, -0.15]], True),      # rotation + scale + translation (the augmentation's range)
    ([[1.20, -0.31, -0.30], [0.31, 1.20, 0.25]], True),      # an inverse transform: magnification, part of the map leaves the frame
    ([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]], True),              # identity: every footprint is a single pixel
    ([[0.05, 0.0, 0.0], [0.0, 0.05, 0.0]], False),           # near-singular: pre-image box too wide, the scatter path serves it
])
def test_equivariance_backward_gather_vs_autograd(ops, theta, gathered):
    """d(maps_t) of the equivariance loss (optimize.py:157-163 through invertable_transform.py:74-92) is a deterministic
    GATHER over the bilinear footprints: equal to torch's grid_sample autograd (fp64) and bit-identical from run to run."""
    g = torch.Generator().manual_seed(5)
    n, r, k = 24, 128, 10
    maps, maps_t = torch.rand(n, r, r, generator=g), torch.rand(n, r, r, generator=g)
    sel = torch.randperm(n, generator=g)[:k]
    th = torch.tensor(theta, dtype=torch.float32)
    mr, mtr = maps.double().requires_grad_(True), maps_t.double().requires_grad_(True)
    grid = F.affine_grid(th.double()[None], (1, k, r, r), align_corners=False)
    un = F.grid_sample(mtr[sel][None], grid, mode="bilinear", padding_mode="zeros", align_corners=False)[0]
    want = F.mse_loss(mr[sel], un)
    want.backward()
    runs = []
    for _ in range(4):
        mc, mtc = cu(maps).requires_grad_(True), cu(maps_t).requires_grad_(True)
        loss = ops.equivariance_loss_op(mc, mtc, cu(sel), cu(th))
        loss.backward()
        runs.append((loss.detach().cpu(), mc.grad.cpu(), mtc.grad.cpu()))
    # uncorrelated pixels: the fp32 sampling coordinate (ulp 8e-6 at 128) times a unit pixel-to-pixel step bounds the error
    assert rel_err(runs[0][0], want.detach()) < 5e-6
    assert rel_err(runs[0][1], mr.grad) < 1e-4
    assert rel_err(runs[0][2], mtr.grad) < 1e-4
    if gathered:
        assert all(torch.equal(runs[0][2], r_[2]) and torch.equal(runs[0][1], r_[1]) for r_ in runs[1:])


def test_reference_style_loss_api(ops):
    """optimize.sharpening_loss / equivariance_loss called the way optimize.py:397-401 calls them."""
    from stablekeypoints_b200 import optimize
    from stablekeypoints_b200.invertable_transform import RandomAffineWithInverse
    post = load_golden("post_unet.npz")
    maps, maps_t, sel = cu(post["maps"]), cu(post["maps_t"]), cu(post["sel"])
    tr = RandomAffineWithInverse(degrees=15, scale=(0.8, 1.0), translate=(0.25, 0.25))
    tr.last_params = {"theta": torch.from_numpy(post["theta"])}
    sharp = optimize.sharpening_loss(maps[sel], device="cuda", sigma=2.0, num_subjects=1)
    equiv = optimize.equivariance_loss(maps[sel], maps_t[sel][None].repeat(1, 1, 1, 1), tr, 0)
    assert rel_err(sharp.cpu(), post["sharp"]) < 5e-6
    assert rel_err(equiv.cpu(), post["equiv"]) < 5e-6
    assert rel_err(tr.inverse(maps_t[sel][None]).cpu(), post["unwarp"]) < 5e-6


def test_affine_warp_and_rng_order(ops):
    from stablekeypoints_b200.invertable_transform import RandomAffineWithInverse
    post = load_golden("post_unet.npz")
    tr = RandomAffineWithInverse(degrees=15, scale=(0.8, 1.0), translate=(0.25, 0.25))
    out = tr(cu(post["img"]), theta=torch.from_numpy(post["theta2"]))
    assert rel_err(out.cpu(), post["warp"]) < 5e-6
    torch.manual_seed(123)
    tr(torch.zeros(3, 1, 4, 4).cuda())
    assert np.allclose(tr.last_params["theta"].cpu().numpy(), post["theta_seed123"], atol=1e-7)
    # 512^2 image, gradient of the sampler vs torch autograd
    g = torch.Generator().manual_seed(0)
    img = torch.rand(1, 3, 512, 512, generator=g)
    th = hp.affine_theta(7.0, 0.93, 0.2, -0.1)
    ir = img.clone().requires_grad_(True)
    wr = hp.affine_warp(ir, th)
    dy = torch.randn(wr.shape, generator=g)
    (wr * dy).sum().backward()
    ic = cu(img).requires_grad_(True)
    wc = ops.affine_warp(ic, th)
    (wc * cu(dy)).sum().backward()
    # 512-px coordinates carry ~3e-5 px of fp32 rounding; on a white-noise image that is ~1e-4 of the value range
    assert rel_err(wc.detach().cpu(), wr.detach()) < 3e-4
    assert rel_err(ic.grad.cpu(), ir.grad) < 3e-4


def test_soft_argmax_golden_and_full_size(ops):
    from stablekeypoints_b200 import eval as skp_eval
    post = load_golden("post_unet.npz")
    hm = cu(post["soft_in"]).clone()
    out = skp_eval.pixel_from_weighted_avg(hm)
    assert rel_err(out.cpu(), post["soft_out"]) < 2e-6
    assert np.array_equal(hm.cpu().numpy(), post["soft_in_after"])
    g = torch.Generator().manual_seed(5)
    big = torch.rand(10, 512, 512, generator=g) ** 6
    want = hp.pixel_from_weighted_avg(big.clone())
    assert rel_err(skp_eval.pixel_from_weighted_avg(cu(big).clone()).cpu(), want) < 2e-6


def test_adam_step(ops):
    g = torch.Generator().manual_seed(9)
    p = torch.randn(1, 77, 768, generator=g)
    pc, m, v = cu(p).clone(), torch.zeros(1, 77, 768).cuda(), torch.zeros(1, 77, 768).cuda()
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=5e-3)
    for step in range(1, 4):
        grad = torch.randn(1, 77, 768, generator=g) * 0.1
        pr.grad = grad.clone()
        opt.step()
        ops.adam_step_(pc, cu(grad * 2), m, v, step, grad_scale=0.5)
    assert rel_err(pc.cpu(), pr.detach()) < 1e-6
