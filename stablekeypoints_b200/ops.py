"""torch-facing wrappers (autograd.Function) over the C ABI of libskp_b200.so.

Every op launches hand-written sm_100a kernels on the current CUDA stream into torch-allocated buffers.
There is no CPU or PyTorch fallback: CPU tensors raise.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import SkpError, check, int_array, lib, ptr, ptr_array, require_cuda, stream

# "tc": tcgen05 split-bf16 tensor-core GEMM (default); "simt": exact-fp32 FMA GEMM.
GEMM_IMPL = os.environ.get("SKP_GEMM_IMPL", "tc")


def set_gemm_impl(name: str) -> None:
    global GEMM_IMPL
    if name not in ("tc", "simt"):
        raise ValueError(name)
    GEMM_IMPL = name


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


# ----------------------------------------------------------------------------- dense projections
def _pad64(k: int) -> int:
    return (k + 63) // 64 * 64


def split_bf16(x: torch.Tensor):
    """x[rows, cols] fp32 -> (hi, lo) bf16 [rows, pad64(cols)], x ~= hi + lo."""
    require_cuda(x)
    assert x.dim() == 2 and x.stride(1) == 1
    rows, cols = x.shape
    kp = _pad64(cols)
    hi = torch.empty(rows, kp, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty(rows, kp, dtype=torch.bfloat16, device=x.device)
    check(lib().skp_split_bf16(ptr(x), x.stride(0), rows, cols, kp, ptr(hi), ptr(lo), stream()), "skp_split_bf16")
    return hi, lo


class FrozenWeight:
    """A frozen [out, in] projection weight prepared once for both directions:
    forward  y = x W^T   (B operand = W,   [out, in]  K-major)
    dgrad    dx = dy W   (B operand = W^T, [in, out]  K-major)
    each as fp32 (SIMT path) and as split-bf16 pairs (tcgen05 path)."""

    def __init__(self, w: torch.Tensor, need_dgrad: bool = True):
        require_cuda(w)
        self.w = _f32c(w.detach())
        self.out_features, self.in_features = self.w.shape
        self.w_split = split_bf16(self.w)
        self.wt = self.wt_split = None
        if need_dgrad:
            self.wt = self.w.t().contiguous()
            self.wt_split = split_bf16(self.wt)


def gemm_nt(a: torch.Tensor, b: torch.Tensor, b_split, bias: Optional[torch.Tensor] = None,
            residual: Optional[torch.Tensor] = None, alpha: float = 1.0) -> torch.Tensor:
    """out[M, N] = alpha * a[M, K] @ b[N, K]^T (+ bias[N]) (+ residual[M, N])."""
    require_cuda(a, b)
    assert a.dim() == 2 and a.stride(1) == 1 and a.dtype == torch.float32
    m, k = a.shape
    n = b.shape[0]
    assert b.shape[1] == k
    out = torch.empty(m, n, dtype=torch.float32, device=a.device)
    if residual is not None:
        assert tuple(residual.shape) == (m, n), f"residual shape {tuple(residual.shape)} != {(m, n)}"
        if residual.stride(1) != 1 or residual.dtype != torch.float32:
            residual = _f32c(residual)
    ldr = residual.stride(0) if residual is not None else 0
    if GEMM_IMPL == "tc":
        a_hi, a_lo = split_bf16(a)
        return gemm_nt_presplit(a_hi, a_lo, m, b_split, n, bias, residual, alpha, out)
    check(lib().skp_gemm_nt_simt(ptr(a), a.stride(0), ptr(b), b.stride(0), ptr(out), out.stride(0), m, n, k, alpha,
                                 ptr(bias), ptr(residual), ldr, stream()), "skp_gemm_nt_simt")
    return out


# K-splits of the small-M layers: "ws" (default) = partial sums in a workspace + one reduce/epilogue kernel; "cluster" = reduced
# inside the GEMM kernel by the CTAs of one thread-block cluster through distributed shared memory (no workspace, no second
# launch, 319 fewer launches per step).  Measured on B200 inside the 3-stream step graph the cluster form is 1.3 % SLOWER
# (43.8 vs 44.4 images/s: 8 co-scheduled 200 KB CTAs per tile fragment the SMs the other two streams are using), so it is opt-in.
SPLITK = os.environ.get("SKP_SPLITK", "ws")


def _splitk(m: int, n: int, kp: int, device):
    if SPLITK != "ws":
        return 0, None
    splits = lib().skp_gemm_nt_tc_plan(m, n, kp)
    return splits, (torch.empty(splits * m * n, dtype=torch.float32, device=device) if splits > 1 else None)


def gemm_nt_presplit(a_hi, a_lo, m: int, b_split, n: int, bias=None, residual=None, alpha: float = 1.0, out=None):
    """tcgen05 GEMM on operands that are already split-bf16 K-major pairs; split-K when the tile count is small."""
    kp = a_hi.shape[1]
    b_hi, b_lo = b_split
    assert b_hi.shape[1] == kp, f"K padding mismatch: A {kp} vs B {b_hi.shape[1]}"
    if out is None:
        out = torch.empty(m, n, dtype=torch.float32, device=a_hi.device)
    if residual is not None and (residual.stride(1) != 1 or residual.dtype != torch.float32):
        residual = _f32c(residual)
    ldr = residual.stride(0) if residual is not None else 0
    splits, ws = _splitk(m, n, kp, out.device)
    check(lib().skp_gemm_nt_tc(ptr(a_hi), ptr(a_lo), ptr(b_hi), ptr(b_lo), kp, ptr(out), out.stride(0), m, n, alpha,
                               ptr(bias), ptr(residual), ldr, splits, ptr(ws), stream()), "skp_gemm_nt_tc")
    return out


class _StreamAlias(torch.autograd.Function):
    """Identity whose backward node belongs to the stream it was created on (autograd replays a node on its forward stream)."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g


def stream_alias(x: torch.Tensor) -> torch.Tensor:
    return _StreamAlias.apply(x) if x.requires_grad else x


# ----------------------------------------------------------------------------- frozen 3x3 convolution (channels-last)
def im2col3x3_split(x2d: torch.Tensor, h: int, w: int, ho: int, wo: int, stride: int, pad: int):
    """x2d [h*w, C] fp32 channels-last -> split-bf16 im2col operand ([ho*wo, pad64(9C)] hi, lo)."""
    require_cuda(x2d)
    assert x2d.dim() == 2 and x2d.stride(1) == 1 and x2d.shape[0] == h * w and x2d.dtype == torch.float32
    c = x2d.shape[1]
    kp = _pad64(9 * c)
    hi = torch.empty(ho * wo, kp, dtype=torch.bfloat16, device=x2d.device)
    lo = torch.empty(ho * wo, kp, dtype=torch.bfloat16, device=x2d.device)
    check(lib().skp_im2col3x3_split(ptr(x2d), x2d.stride(0), h, w, c, ho, wo, stride, pad, kp, ptr(hi), ptr(lo), stream()),
          "skp_im2col3x3_split")
    return hi, lo


def _implicit_ok(w: int, cin: int) -> bool:
    return cin % 64 == 0 and w % 8 == 0 and os.environ.get("SKP_CONV_IMPLICIT", "1") != "0"


def conv3x3_implicit(x_hi, x_lo, h: int, w: int, b_split, cout: int, bias=None, residual=None):
    """Stride-1 / pad-1 3x3 convolution as an implicit GEMM: the split-bf16 channels-last activation [h*w, Cin] is read
    through 3-D TMA boxes shifted per tap (no im2col buffer)."""
    cin = x_hi.shape[1]
    b_hi, b_lo = b_split
    assert b_hi.shape[1] == 9 * cin
    out = torch.empty(h * w, cout, dtype=torch.float32, device=x_hi.device)
    if residual is not None and (residual.stride(1) != 1 or residual.dtype != torch.float32):
        residual = _f32c(residual)
    ldr = residual.stride(0) if residual is not None else 0
    splits, ws = _splitk(h * w, cout, 9 * cin, out.device)
    check(lib().skp_conv3x3_tc(ptr(x_hi), ptr(x_lo), h, w, cin, ptr(b_hi), ptr(b_lo), ptr(out), out.stride(0), cout, 1.0,
                               ptr(bias), ptr(residual), ldr, splits, ptr(ws), stream()), "skp_conv3x3_tc")
    return out


class FrozenConv3x3:
    """Frozen [Cout, Cin, 3, 3] filter prepared as GEMM operands: forward  W[cout, tap*Cin + cin]; input-gradient
    (stride 1) W'[cin, tap*Cout + cout] with the taps flipped.  Both as split-bf16 K-major pairs."""

    def __init__(self, w: torch.Tensor, need_dgrad: bool = True):
        require_cuda(w)
        w = w.detach().float()
        self.cout, self.cin = w.shape[0], w.shape[1]
        self.fwd_split = split_bf16(w.permute(0, 2, 3, 1).reshape(self.cout, 9 * self.cin).contiguous())
        self.dgrad_split = None
        self.w_nchw = None
        if need_dgrad:
            self.dgrad_split = split_bf16(w.flip(2, 3).permute(1, 2, 3, 0).reshape(self.cin, 9 * self.cout).contiguous())
            self.w_nchw = w.contiguous()  # only used by the stride-2 input gradient (3 down-samplers)


class _FrozenConv3x3(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2d, residual, fcw: FrozenConv3x3, bias, h, w, stride, pad, ho, wo):
        x2d = _f32c(x2d)
        ctx.fcw, ctx.geom, ctx.has_res = fcw, (h, w, stride, pad, ho, wo), residual is not None
        if stride == 1 and pad == 1 and _implicit_ok(w, fcw.cin):
            hi, lo = split_bf16(x2d)
            return conv3x3_implicit(hi, lo, h, w, fcw.fwd_split, fcw.cout, bias, residual)
        hi, lo = im2col3x3_split(x2d, h, w, ho, wo, stride, pad)
        return gemm_nt_presplit(hi, lo, ho * wo, fcw.fwd_split, fcw.cout, bias, residual)

    @staticmethod
    def backward(ctx, dy):
        fcw = ctx.fcw
        h, w, stride, pad, ho, wo = ctx.geom
        dy = _f32c(dy)
        dx = None
        if ctx.needs_input_grad[0]:
            if stride == 1 and pad == 1 and _implicit_ok(w, fcw.cout):
                hi, lo = split_bf16(dy)
                dx = conv3x3_implicit(hi, lo, h, w, fcw.dgrad_split, fcw.cin)
            elif stride == 1 and pad == 1:
                hi, lo = im2col3x3_split(dy, ho, wo, h, w, 1, 1)
                dx = gemm_nt_presplit(hi, lo, h * w, fcw.dgrad_split, fcw.cin)
            elif stride == 2 and pad == 1 and h == 2 * ho and w == 2 * wo and _implicit_ok(w, fcw.cout):
                # the UNet down-samplers: a stride-2 convolution's input gradient is the stride-1 input gradient of the
                # zero-stuffed output gradient (dyz[2o] = dy[o]) -> same implicit-GEMM kernel, no library call
                dyz = torch.zeros(h, w, fcw.cout, dtype=torch.float32, device=dy.device)
                dyz[::2, ::2] = dy.reshape(ho, wo, fcw.cout)
                hi, lo = split_bf16(dyz.reshape(h * w, fcw.cout))
                dx = conv3x3_implicit(hi, lo, h, w, fcw.dgrad_split, fcw.cin)
            else:  # odd geometries only (not reached by SD1.5): cuDNN dgrad on an NCHW view
                g = dy.reshape(1, ho, wo, fcw.cout).permute(0, 3, 1, 2)
                # the geometry is "rows/cols beyond the image read zero": the padded extent that makes (ho, wo) exact
                hp_, wp_ = (ho - 1) * stride + 3 - pad, (wo - 1) * stride + 3 - pad
                dxn = torch.nn.grad.conv2d_input((1, fcw.cin, max(h, hp_) + pad, max(w, wp_) + pad), fcw.w_nchw, g,
                                                 stride=stride, padding=0)
                dx = dxn[:, :, pad:pad + h, pad:pad + w].permute(0, 2, 3, 1).reshape(h * w, fcw.cin).contiguous()
        dres = dy if (ctx.has_res and ctx.needs_input_grad[1]) else None
        return dx, dres, None, None, None, None, None, None, None, None


def frozen_conv3x3(x2d, h: int, w: int, fcw: FrozenConv3x3, bias=None, residual=None, stride: int = 1, pad: int = 1,
                   out_hw=None):
    """3x3 convolution of a channels-last activation [h*w, Cin] with a frozen filter -> [ho*wo, Cout] (+bias +residual):
    im2col straight to split-bf16, then the tcgen05 GEMM.  Returns (y2d, ho, wo)."""
    if out_hw is None:
        ho, wo = (h + 2 * pad - 3) // stride + 1, (w + 2 * pad - 3) // stride + 1
    else:
        ho, wo = out_hw
    return _FrozenConv3x3.apply(x2d, residual, fcw, bias, h, w, stride, pad, ho, wo), ho, wo


class _FrozenLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, fw: FrozenWeight, bias):
        ctx.fw = fw
        ctx.has_res = residual is not None
        return gemm_nt(_f32c(x), fw.w, fw.w_split, bias, residual)

    @staticmethod
    def backward(ctx, dy):
        fw = ctx.fw
        dy = _f32c(dy)
        dx = gemm_nt(dy, fw.wt, fw.wt_split) if ctx.needs_input_grad[0] else None
        dres = dy if (ctx.has_res and ctx.needs_input_grad[1]) else None
        return dx, dres, None, None


def frozen_linear(x: torch.Tensor, fw: FrozenWeight, bias: Optional[torch.Tensor] = None,
                  residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = x W^T + bias + residual with a frozen weight (only input gradients exist)."""
    return _FrozenLinear.apply(x, residual, fw, bias)


# ----------------------------------------------------------------------------- GroupNorm(+SiLU) fused into what follows
GN_REPL = 8   # replicas of the fp64 accumulators (skp_norm.cu GN_REPL; include/skp_b200.h)


def _gn_stats(x2d: torch.Tensor, groups: int) -> torch.Tensor:
    sums = torch.empty(2 * groups * GN_REPL, dtype=torch.float64, device=x2d.device)
    check(lib().skp_gn_stats(ptr(x2d), x2d.stride(0), x2d.shape[0], x2d.shape[1], groups, ptr(sums), stream()), "skp_gn_stats")
    return sums


def _gn_fwd(x2d, groups, eps, gamma, beta, silu, y, hi, lo, kp) -> torch.Tensor:
    """Statistics + act(GroupNorm(x)) into fp32 y and/or the split-bf16 operand (hi, lo) [rows, kp]; returns the fp64 sums
    the backward needs.  Small activations take one launch (one CTA per group), large ones stats + apply."""
    sums = torch.empty(2 * groups * GN_REPL, dtype=torch.float64, device=x2d.device)
    check(lib().skp_gn_fwd(ptr(x2d), x2d.stride(0), x2d.shape[0], x2d.shape[1], groups, eps, ptr(gamma), ptr(beta), int(silu),
                           ptr(y), y.stride(0) if y is not None else 0, ptr(hi), ptr(lo), kp, ptr(sums), stream()), "skp_gn_fwd")
    return sums


def _gn_backward(x2d, g, sums, gamma, beta, groups, eps, silu):
    g = _f32c(g)
    dx = torch.empty_like(x2d)
    bs = torch.empty(2 * groups * GN_REPL, dtype=torch.float64, device=x2d.device)
    check(lib().skp_gn_bwd(ptr(x2d), x2d.stride(0), ptr(g), g.stride(0), x2d.shape[0], x2d.shape[1], groups, ptr(sums), eps,
                           ptr(gamma), ptr(beta), int(silu), ptr(bs), ptr(dx), dx.stride(0), stream()), "skp_gn_bwd")
    return dx


class _GNConv3x3(torch.autograd.Function):
    """y = conv3x3(act(GroupNorm(x))) + bias (+ residual) on channels-last [h*w, C]; the normalised activation only ever
    exists as the split-bf16 im2col operand."""

    @staticmethod
    def forward(ctx, x2d, residual, gamma, beta, fcw: FrozenConv3x3, bias, groups, eps, silu, h, w, stride, pad, ho, wo):
        x2d = _f32c(x2d)
        c = x2d.shape[1]
        ctx.meta = (fcw, groups, eps, silu, h, w, stride, pad, ho, wo, residual is not None)
        if stride == 1 and pad == 1 and _implicit_ok(w, c):
            # normalised activation written ONCE as split-bf16 channels-last; the conv reads it through shifted TMA boxes
            hi = torch.empty(h * w, c, dtype=torch.bfloat16, device=x2d.device)
            lo = torch.empty(h * w, c, dtype=torch.bfloat16, device=x2d.device)
            sums = _gn_fwd(x2d, groups, eps, gamma, beta, silu, None, hi, lo, c)
            ctx.save_for_backward(x2d, sums, gamma, beta)
            return conv3x3_implicit(hi, lo, h, w, fcw.fwd_split, fcw.cout, bias, residual)
        sums = _gn_stats(x2d, groups)
        ctx.save_for_backward(x2d, sums, gamma, beta)
        kp = _pad64(9 * c)
        hi = torch.empty(ho * wo, kp, dtype=torch.bfloat16, device=x2d.device)
        lo = torch.empty(ho * wo, kp, dtype=torch.bfloat16, device=x2d.device)
        check(lib().skp_gn_im2col3x3_split(ptr(x2d), x2d.stride(0), h, w, c, groups, ptr(sums), eps, ptr(gamma), ptr(beta),
                                           int(silu), ho, wo, stride, pad, kp, ptr(hi), ptr(lo), stream()),
              "skp_gn_im2col3x3_split")
        return gemm_nt_presplit(hi, lo, ho * wo, fcw.fwd_split, fcw.cout, bias, residual)

    @staticmethod
    def backward(ctx, dy):
        x2d, sums, gamma, beta = ctx.saved_tensors
        fcw, groups, eps, silu, h, w, stride, pad, ho, wo, has_res = ctx.meta
        dy = _f32c(dy)
        dx = None
        if ctx.needs_input_grad[0]:
            if stride == 1 and pad == 1 and _implicit_ok(w, fcw.cout):
                hi, lo = split_bf16(dy)
                d_act = conv3x3_implicit(hi, lo, h, w, fcw.dgrad_split, fcw.cin)
            elif stride == 1 and pad == 1:
                hi, lo = im2col3x3_split(dy, ho, wo, h, w, 1, 1)
                d_act = gemm_nt_presplit(hi, lo, h * w, fcw.dgrad_split, fcw.cin)
            else:
                g = dy.reshape(1, ho, wo, fcw.cout).permute(0, 3, 1, 2)
                hp_, wp_ = (ho - 1) * stride + 3 - pad, (wo - 1) * stride + 3 - pad
                dxn = torch.nn.grad.conv2d_input((1, fcw.cin, max(h, hp_) + pad, max(w, wp_) + pad), fcw.w_nchw, g,
                                                 stride=stride, padding=0)
                d_act = dxn[:, :, pad:pad + h, pad:pad + w].permute(0, 2, 3, 1).reshape(h * w, fcw.cin).contiguous()
            dx = _gn_backward(x2d, d_act, sums, gamma, beta, groups, eps, silu)
        dres = dy if (has_res and ctx.needs_input_grad[1]) else None
        return (dx, dres) + (None,) * 13


def gn_conv3x3(x2d, h, w, gamma, beta, groups, eps, silu, fcw: FrozenConv3x3, bias=None, residual=None, stride=1, pad=1,
               out_hw=None):
    if out_hw is None:
        ho, wo = (h + 2 * pad - 3) // stride + 1, (w + 2 * pad - 3) // stride + 1
    else:
        ho, wo = out_hw
    return _GNConv3x3.apply(x2d, residual, gamma, beta, fcw, bias, groups, eps, silu, h, w, stride, pad, ho, wo), ho, wo


class _GNLinear(torch.autograd.Function):
    """y = act(GroupNorm(x)) W^T + bias (+ residual): the normalised activation is written straight as the split-bf16
    A operand of the projection."""

    @staticmethod
    def forward(ctx, x2d, residual, gamma, beta, fw: FrozenWeight, bias, groups, eps, silu):
        x2d = _f32c(x2d)
        rows, c = x2d.shape
        kp = _pad64(c)
        hi = torch.empty(rows, kp, dtype=torch.bfloat16, device=x2d.device)
        lo = torch.empty(rows, kp, dtype=torch.bfloat16, device=x2d.device)
        sums = _gn_fwd(x2d, groups, eps, gamma, beta, silu, None, hi, lo, kp)
        ctx.save_for_backward(x2d, sums, gamma, beta)
        ctx.meta = (fw, groups, eps, silu, residual is not None)
        return gemm_nt_presplit(hi, lo, rows, fw.w_split, fw.out_features, bias, residual)

    @staticmethod
    def backward(ctx, dy):
        x2d, sums, gamma, beta = ctx.saved_tensors
        fw, groups, eps, silu, has_res = ctx.meta
        dy = _f32c(dy)
        dx = None
        if ctx.needs_input_grad[0]:
            d_act = gemm_nt(dy, fw.wt, fw.wt_split)
            dx = _gn_backward(x2d, d_act, sums, gamma, beta, groups, eps, silu)
        dres = dy if (has_res and ctx.needs_input_grad[1]) else None
        return (dx, dres) + (None,) * 7


def gn_linear(x2d, gamma, beta, groups, eps, silu, fw: FrozenWeight, bias=None, residual=None):
    return _GNLinear.apply(x2d, residual, gamma, beta, fw, bias, groups, eps, silu)


# ----------------------------------------------------------------------------- LayerNorm / GEGLU fused into the projection
class _LNLinear(torch.autograd.Function):
    """y = LayerNorm(x) W^T + bias (+ residual) (BasicTransformerBlock norm1/2/3 -> qkv / to_q / ff.net.0.proj): the
    normalised rows are written straight as the split-bf16 A operand (skp_rowops.cu)."""

    @staticmethod
    def forward(ctx, x2d, residual, gamma, beta, fw: FrozenWeight, bias, eps):
        x2d = _f32c(x2d)
        rows, c = x2d.shape
        kp = _pad64(c)
        hi = torch.empty(rows, kp, dtype=torch.bfloat16, device=x2d.device)
        lo = torch.empty(rows, kp, dtype=torch.bfloat16, device=x2d.device)
        stats = torch.empty(rows, 2, dtype=torch.float32, device=x2d.device)
        check(lib().skp_ln_split_fwd(ptr(x2d), x2d.stride(0), rows, c, ptr(gamma), ptr(beta), eps, ptr(hi), ptr(lo), kp, ptr(stats),
                                     stream()), "skp_ln_split_fwd")
        ctx.save_for_backward(x2d, stats, gamma)
        ctx.meta = (fw, residual is not None)
        return gemm_nt_presplit(hi, lo, rows, fw.w_split, fw.out_features, bias, residual)

    @staticmethod
    def backward(ctx, dy):
        x2d, stats, gamma = ctx.saved_tensors
        fw, has_res = ctx.meta
        dy = _f32c(dy)
        dx = None
        if ctx.needs_input_grad[0]:
            d_act = gemm_nt(dy, fw.wt, fw.wt_split)
            dx = torch.empty_like(x2d)
            check(lib().skp_ln_bwd(ptr(x2d), x2d.stride(0), ptr(d_act), d_act.stride(0), x2d.shape[0], x2d.shape[1], ptr(gamma),
                                   ptr(stats), ptr(dx), dx.stride(0), stream()), "skp_ln_bwd")
        dres = dy if (has_res and ctx.needs_input_grad[1]) else None
        return (dx, dres) + (None,) * 5


def ln_linear(x2d, gamma, beta, fw: FrozenWeight, bias=None, residual=None, eps: float = 1e-5):
    return _LNLinear.apply(x2d, residual, gamma, beta, fw, bias, eps)


class _GegluLinear(torch.autograd.Function):
    """y = (a * gelu(gate)) W^T + bias (+ residual) with proj = (a | gate) (diffusers GEGLU + ff.net.2): the gated
    activation only exists as the split-bf16 A operand."""

    @staticmethod
    def forward(ctx, proj, residual, fw: FrozenWeight, bias):
        proj = _f32c(proj)
        rows, h2 = proj.shape
        h = h2 // 2
        kp = _pad64(h)
        hi = torch.empty(rows, kp, dtype=torch.bfloat16, device=proj.device)
        lo = torch.empty(rows, kp, dtype=torch.bfloat16, device=proj.device)
        check(lib().skp_geglu_split_fwd(ptr(proj), proj.stride(0), rows, h, ptr(hi), ptr(lo), kp, stream()), "skp_geglu_split_fwd")
        ctx.save_for_backward(proj)
        ctx.meta = (fw, residual is not None)
        return gemm_nt_presplit(hi, lo, rows, fw.w_split, fw.out_features, bias, residual)

    @staticmethod
    def backward(ctx, dy):
        (proj,) = ctx.saved_tensors
        fw, has_res = ctx.meta
        dy = _f32c(dy)
        dproj = None
        if ctx.needs_input_grad[0]:
            d_act = gemm_nt(dy, fw.wt, fw.wt_split)
            dproj = torch.empty_like(proj)
            check(lib().skp_geglu_bwd(ptr(proj), proj.stride(0), ptr(d_act), d_act.stride(0), proj.shape[0], proj.shape[1] // 2,
                                      ptr(dproj), dproj.stride(0), stream()), "skp_geglu_bwd")
        dres = dy if (has_res and ctx.needs_input_grad[1]) else None
        return dproj, dres, None, None


def geglu_linear(proj, fw: FrozenWeight, bias=None, residual=None):
    return _GegluLinear.apply(proj, residual, fw, bias)


def dense_attention(q, k, v, scale: float) -> torch.Tensor:
    """Single-head softmax(q k^T scale) v for wide heads (VAE mid block: d = 512), forward only: two tcgen05 split-bf16
    GEMMs around a fused softmax+operand-split kernel.  q, k, v: [S, d] fp32 (row stride free)."""
    require_cuda(q, k, v)
    s_q, d = q.shape
    s_k = k.shape[0]
    if s_k % 4:
        raise SkpError("dense_attention needs a key count that is a multiple of 4")
    scores = gemm_nt(q, k, split_bf16(k), alpha=scale)                    # [s_q, s_k]
    kp = _pad64(s_k)
    p_hi = torch.empty(s_q, kp, dtype=torch.bfloat16, device=q.device)
    p_lo = torch.empty(s_q, kp, dtype=torch.bfloat16, device=q.device)
    check(lib().skp_softmax_split_fwd(ptr(scores), scores.stride(0), s_q, s_k, ptr(p_hi), ptr(p_lo), kp, stream()),
          "skp_softmax_split_fwd")
    vt = v.t().contiguous()                                               # [d, s_k]: K-major B operand of P v
    return gemm_nt_presplit(p_hi, p_lo, s_q, split_bf16(vt), d)


def group_norm_cl(x2d, gamma, beta, groups, eps, silu=False):
    """Plain act(GroupNorm(x)) -> fp32 [rows, C] (forward only; used by tests and no-grad paths)."""
    x2d = _f32c(x2d)
    y = torch.empty_like(x2d)
    _gn_fwd(x2d, groups, eps, gamma, beta, silu, y, None, None, 0)
    return y


# ----------------------------------------------------------------------------- cross-attention core
class _CrossAttnCore(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, heads: int, scale: float, want_logits: bool):
        require_cuda(q, k, v)
        q = _f32c(q)
        assert k.stride(1) == 1 and v.stride(1) == 1 and k.dtype == torch.float32 and v.dtype == torch.float32
        s, c = q.shape
        n = k.shape[0]
        d = c // heads
        o = torch.empty_like(q)
        logits = torch.empty(heads, s, n, dtype=torch.float32, device=q.device)
        check(lib().skp_cross_attn_fwd(ptr(q), ptr(k), k.stride(0), ptr(v), v.stride(0), ptr(o), ptr(logits), s, n, heads, d,
                                       scale, stream()), "skp_cross_attn_fwd")
        ctx.save_for_backward(q, k, v, logits)
        ctx.heads, ctx.scale = heads, scale
        if not want_logits:
            ctx.mark_non_differentiable(logits)
        return o, logits

    @staticmethod
    def backward(ctx, d_o, d_logits):
        q, k, v, logits = ctx.saved_tensors
        heads, scale = ctx.heads, ctx.scale
        s, c = q.shape
        n = k.shape[0]
        d = c // heads
        if d_o is None:
            d_o = torch.zeros_like(q)
        d_o = _f32c(d_o)
        extra = _f32c(d_logits) if d_logits is not None else None
        ws = torch.empty(heads * s * (n + 2), dtype=torch.float32, device=q.device)
        dq = torch.empty_like(q)
        dk = torch.zeros(n, c, dtype=torch.float32, device=q.device)
        dv = torch.zeros(n, c, dtype=torch.float32, device=q.device)
        check(lib().skp_cross_attn_bwd(ptr(d_o), ptr(q), ptr(k), k.stride(0), ptr(v), v.stride(0), ptr(logits), ptr(extra),
                                       ptr(ws), ptr(dq), ptr(dk), ptr(dv), s, n, heads, d, scale, stream()),
              "skp_cross_attn_bwd")
        return dq, dk, dv, None, None, None


class _CrossAttnTC(torch.autograd.Function):
    """Same contract as _CrossAttnCore on the flash-style split-bf16 tensor-core kernels (skp_selfattn.cu): no [h,S,N]
    tensor unless the layer is captured (want_logits), log-sum-exp saved instead of the probabilities."""

    @staticmethod
    def forward(ctx, q, k, v, heads: int, scale: float, want_logits: bool):
        require_cuda(q, k, v)
        q = _f32c(q)
        assert k.stride(1) == 1 and v.stride(1) == 1 and k.dtype == torch.float32 and v.dtype == torch.float32
        s, c = q.shape
        n = k.shape[0]
        d = c // heads
        dp = int(lib().skp_self_attn_dp(d))
        if dp == 0 or d % 2:
            raise SkpError(f"cross-attention head dim {d} unsupported on the tensor-core path (even, <= {SELF_ATTN_MAX_D})")
        dev = q.device
        o = torch.empty_like(q)
        lse = torch.empty(heads, s, dtype=torch.float32, device=dev)
        logits = torch.empty(heads, s, n, dtype=torch.float32, device=dev) if want_logits else None
        ws_bytes = int(lib().skp_xattn_tc_workspace(s, n, heads, d)) if XATTN_TC else 0
        if ws_bytes > 0:
            # tcgen05 / TMEM / TMA forward (skp_xattn_tc.cu); the captured layers get their scaled logits from the same kernel
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            check(lib().skp_xattn_tc_fwd(ptr(q), c, ptr(k), k.stride(0), ptr(v), v.stride(0), ptr(o), c, ptr(lse), ptr(logits),
                                         ptr(ws), s, n, heads, d, scale, stream()), "skp_xattn_tc_fwd")
            ctx.save_for_backward(o, lse, q, k, v)
            ctx.tc = True
        else:
            qp = torch.empty(2 * heads * s * dp, dtype=torch.bfloat16, device=dev)
            kvp = torch.empty(4 * heads * n * dp, dtype=torch.bfloat16, device=dev)
            check(lib().skp_cross_attn_tc_fwd(ptr(q), c, ptr(k), k.stride(0), ptr(v), v.stride(0), ptr(o), c, ptr(lse), ptr(logits),
                                              ptr(qp), ptr(kvp), s, n, heads, d, scale, stream()), "skp_cross_attn_tc_fwd")
            ctx.save_for_backward(o, lse, qp, kvp)
            ctx.tc = False
        ctx.meta = (s, n, c, heads, d, dp, scale)
        if logits is None:
            logits = torch.empty(0, dtype=torch.float32, device=dev)
            ctx.mark_non_differentiable(logits)
        return o, logits

    @staticmethod
    def backward(ctx, d_o, d_logits):
        s, n, c, heads, d, dp, scale = ctx.meta
        if ctx.tc:       # the tcgen05 forward keeps its own operand layout: make the mma.sync planes from q / k / v now
            o, lse, q, k, v = ctx.saved_tensors
            qp = torch.empty(2 * heads * s * dp, dtype=torch.bfloat16, device=o.device)
            kvp = torch.empty(4 * heads * n * dp, dtype=torch.bfloat16, device=o.device)
            check(lib().skp_cross_attn_split(ptr(q), c, ptr(k), k.stride(0), ptr(v), v.stride(0), ptr(qp), ptr(kvp), s, n, heads, d,
                                             scale, stream()), "skp_cross_attn_split")
        else:
            o, lse, qp, kvp = ctx.saved_tensors
        dev = o.device
        if d_o is None:
            d_o = torch.zeros_like(o)
        d_o = _f32c(d_o)
        extra = _f32c(d_logits) if d_logits is not None else None
        do_planes = torch.empty(2 * heads * s * dp, dtype=torch.bfloat16, device=dev)
        dvec = torch.empty(heads, s, dtype=torch.float32, device=dev)
        dq = torch.empty(s, c, dtype=torch.float32, device=dev)
        dkv = torch.zeros(2, n, c, dtype=torch.float32, device=dev)
        check(lib().skp_cross_attn_tc_bwd(ptr(d_o), c, ptr(o), c, ptr(lse), ptr(qp), ptr(kvp), ptr(do_planes), ptr(dvec),
                                          ptr(extra), ptr(dq), c, ptr(dkv[0]), c, ptr(dkv[1]), c, s, n, heads, d, scale,
                                          stream()), "skp_cross_attn_tc_bwd")
        return dq, dkv[0], dkv[1], None, None, None


# "tc": split-bf16 tensor-core kernels (default); "simt": the exact-fp32 FMA kernels of skp_attn.cu
CROSS_ATTN_IMPL = os.environ.get("SKP_CROSS_ATTN", "tc")
# cross-attention forward on the tcgen05 kernel of skp_xattn_tc.cu.  Measured on B200 (scripts/xattn_bench.py,
# profiles/r02_xattn_tc.md): with N = 77 .. 500 keys the op is a latency chain of 2 - 8 key tiles on 16 - 256 CTAs, and the
# register-resident mma.sync flash kernel of skp_selfattn.cu finishes it sooner (12 - 19 us vs 17 - 26 us per call) -- the
# tensor-core issue rate is not what bounds it.  So the tcgen05 kernel is opt-in (SKP_XATTN_TC=1; parity-tested either way).
XATTN_TC = os.environ.get("SKP_XATTN_TC", "0") == "1"


def cross_attn_core(q, k, v, heads: int, scale: float, want_logits: bool = False, impl: Optional[str] = None):
    """(out[S,C], scaled logits[h,S,N]) = softmax(q k^T scale) v per head.  The logits tensor is only meaningful when
    want_logits (captured layers); the tensor-core path returns an empty tensor otherwise."""
    impl = impl or CROSS_ATTN_IMPL
    d = q.shape[1] // heads
    if impl == "tc" and d % 2 == 0 and d <= SELF_ATTN_MAX_D:
        return _CrossAttnTC.apply(q, k, v, heads, scale, want_logits)
    return _CrossAttnCore.apply(q, k, v, heads, scale, want_logits)


# ----------------------------------------------------------------------------- self-attention core
SELF_ATTN_MAX_D = 160


# long sequences (S % 128 == 0, d <= 64): forward on the tcgen05/TMEM kernel of skp_attn_tc.cu ("0" = mma.sync kernels only)
SELF_ATTN_TC = os.environ.get("SKP_SELF_ATTN_TC", "1") != "0"
SELF_ATTN_TC_MIN_S = 1024
# ... and their backward on the tcgen05 kernel of skp_attn_tc_bwd.cu ("0" = mma.sync backward on re-split planes)
SELF_ATTN_TC_BWD = os.environ.get("SKP_SELF_ATTN_TC_BWD", "1") != "0"


class _SelfAttnCore(torch.autograd.Function):
    """softmax(q k^T scale) v per head from the packed projection qkv[S, 3C] (columns q | k | v, head-major inside
    each): flash-style split-bf16 tensor-core kernels.  Forward: tcgen05 kernel (skp_attn_tc.cu) for long sequences with
    d <= 64, mma.sync kernel (skp_selfattn.cu) otherwise; backward: tcgen05 kernel (skp_attn_tc_bwd.cu) for long sequences
    with d <= 96, mma.sync kernels on split operand planes otherwise."""

    @staticmethod
    def forward(ctx, qkv, heads: int, scale: float):
        require_cuda(qkv)
        qkv = _f32c(qkv)
        s, c3 = qkv.shape
        c = c3 // 3
        d = c // heads
        dp = int(lib().skp_self_attn_dp(d))
        if dp == 0 or d % 2:
            raise SkpError(f"self-attention head dim {d} unsupported (even, <= {SELF_ATTN_MAX_D})")
        o = torch.empty(s, c, dtype=torch.float32, device=qkv.device)
        lse = torch.empty(heads, s, dtype=torch.float32, device=qkv.device)
        e = qkv.element_size()
        qp, kp, vp = qkv.data_ptr(), qkv.data_ptr() + c * e, qkv.data_ptr() + 2 * c * e
        ws_bytes = int(lib().skp_self_attn_tc_workspace(s, heads, d)) if (SELF_ATTN_TC and s >= SELF_ATTN_TC_MIN_S) else 0
        # backward on tcgen05 (skp_attn_tc_bwd.cu: S % 128 == 0, d <= 96) whichever forward ran: both write o and a base-2 lse
        ctx.tc_bwd = bool(SELF_ATTN_TC and SELF_ATTN_TC_BWD and s >= SELF_ATTN_TC_MIN_S
                          and int(lib().skp_self_attn_tc_bwd_workspace(s, heads, d)) > 0)
        if ws_bytes > 0:
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=qkv.device)
            check(lib().skp_self_attn_tc_fwd(qp, c3, kp, c3, vp, c3, ptr(o), c, ptr(lse), ptr(ws), s, heads, d, scale, stream()),
                  "skp_self_attn_tc_fwd")
            ctx.save_for_backward(o, lse, qkv)
            ctx.tc = True
        else:
            planes = torch.empty(6 * heads * s * dp, dtype=torch.bfloat16, device=qkv.device)
            check(lib().skp_self_attn_fwd(qp, c3, kp, c3, vp, c3, ptr(o), c, ptr(lse), ptr(planes), s, heads, d, scale, stream()),
                  "skp_self_attn_fwd")
            ctx.save_for_backward(o, lse, qkv if ctx.tc_bwd else planes)
            ctx.tc = False
        ctx.meta = (s, c, heads, d, dp, scale)
        return o

    @staticmethod
    def backward(ctx, d_o):
        o, lse, third = ctx.saved_tensors
        s, c, heads, d, dp, scale = ctx.meta
        d_o = _f32c(d_o)
        if ctx.tc_bwd:   # tcgen05 backward: re-splits q / k / v / dO into its own operand planes
            qkv, e, c3 = third, third.element_size(), 3 * c
            ws = torch.empty(int(lib().skp_self_attn_tc_bwd_workspace(s, heads, d)), dtype=torch.uint8, device=o.device)
            dqkv = torch.empty(s, 3 * c, dtype=torch.float32, device=o.device)
            base = dqkv.data_ptr()
            check(lib().skp_self_attn_tc_bwd(ptr(d_o), c, ptr(o), c, ptr(lse), qkv.data_ptr(), c3, qkv.data_ptr() + c * e, c3,
                                             qkv.data_ptr() + 2 * c * e, c3, ptr(ws), base, 3 * c, base + 4 * c, 3 * c,
                                             base + 8 * c, 3 * c, s, heads, d, scale, stream()), "skp_self_attn_tc_bwd")
            return dqkv, None, None
        if ctx.tc:        # the tcgen05 forward keeps its own operand layout: make the mma.sync planes from qkv now
            qkv, e, c3 = third, third.element_size(), 3 * c
            planes = torch.empty(6 * heads * s * dp, dtype=torch.bfloat16, device=o.device)
            check(lib().skp_self_attn_split(qkv.data_ptr(), c3, qkv.data_ptr() + c * e, c3, qkv.data_ptr() + 2 * c * e, c3,
                                            ptr(planes), s, heads, d, scale, stream()), "skp_self_attn_split")
        else:
            planes = third
        do_planes = torch.empty(2 * heads * s * dp, dtype=torch.bfloat16, device=o.device)
        dvec = torch.empty(heads, s, dtype=torch.float32, device=o.device)
        dqkv = torch.empty(s, 3 * c, dtype=torch.float32, device=o.device)
        base = dqkv.data_ptr()
        check(lib().skp_self_attn_bwd(ptr(d_o), c, ptr(o), c, ptr(lse), ptr(planes), ptr(do_planes), ptr(dvec),
                                      base, 3 * c, base + 4 * c, 3 * c, base + 8 * c, 3 * c, s, heads, d, scale, stream()),
              "skp_self_attn_bwd")
        return dqkv, None, None


def self_attn_core(qkv: torch.Tensor, heads: int, scale: float) -> torch.Tensor:
    """out[S, C] = softmax(q k^T scale) v per head, qkv = [S, 3C] packed projection (attn1, ptp_utils.py:480-506)."""
    return _SelfAttnCore.apply(qkv, heads, scale)


# ----------------------------------------------------------------------------- capture (attention store)
def _side(logits: torch.Tensor) -> int:
    s = int(round(logits.shape[1] ** 0.5))
    assert s * s == logits.shape[1], "captured layers must have a square token grid"
    return s


class _CaptureStore(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, res: int):
        require_cuda(logits)
        logits = _f32c(logits)
        h, _, n = logits.shape
        s = _side(logits)
        probs = torch.empty(h, res * res, n, dtype=torch.float32, device=logits.device)
        if lib().skp_capture_tc_ok(int_array([s]), 1, n, res, 1):      # tcgen05 attn-store kernel (N <= 128)
            ws = torch.empty(int(lib().skp_capture_tc_workspace(int_array([s]), 1, h)), dtype=torch.uint8, device=logits.device)
            check(lib().skp_capture_store_tc_fwd(ptr(logits), ptr(probs), h, s, n, res, ptr(ws), stream()), "skp_capture_store_tc_fwd")
        else:
            check(lib().skp_capture_store_fwd(ptr(logits), ptr(probs), h, s, n, res, stream()), "skp_capture_store_fwd")
        ctx.save_for_backward(logits)
        ctx.res = res
        return probs

    @staticmethod
    def backward(ctx, d_probs):
        (logits,) = ctx.saved_tensors
        h, _, n = logits.shape
        d_logits = torch.zeros_like(logits)
        check(lib().skp_capture_store_bwd(ptr(logits), ptr(_f32c(d_probs)), ptr(d_logits), h, _side(logits), n, ctx.res,
                                          stream()), "skp_capture_store_bwd")
        return d_logits, None


def capture_store(logits: torch.Tensor, res: int) -> torch.Tensor:
    """probs[h, res*res, N]: what AttentionStore.step_store["attn"] holds (ptp_utils.py:508-538)."""
    return _CaptureStore.apply(logits, res)


# forward of the fused capture+collect: "fused" = one tile kernel over all (layer, head) slices; "store" = row attn-store
# kernel per layer + collect mean (A/B measured in scripts/kernel_bench.py; the backward is the fused kernel either way)
CAPTURE_MEAN_FWD = os.environ.get("SKP_CAPTURE_MEAN_FWD", "store")


class _CaptureMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, res: int, *logits):
        require_cuda(*logits)
        logits = [_f32c(l) for l in logits]
        h, _, n = logits[0].shape
        sides = [_side(l) for l in logits]
        maps = torch.empty(n, res, res, dtype=torch.float32, device=logits[0].device)
        if lib().skp_capture_tc_ok(int_array(sides), len(logits), n, res, 0):
            # tcgen05 kernel: one launch, CTA = output row, the (layer, head) mean accumulated in registers
            ws = torch.empty(int(lib().skp_capture_tc_workspace(int_array(sides), len(logits), h)), dtype=torch.uint8, device=maps.device)
            check(lib().skp_capture_mean_tc_fwd(ptr_array(logits), int_array(sides), len(logits), ptr(maps), h, n, res, ptr(ws), stream()),
                  "skp_capture_mean_tc_fwd")
        elif CAPTURE_MEAN_FWD == "store" and h * res * res * n * 4 * len(logits) <= (1 << 30):
            # forward through the row attn-store kernel + the collect mean: the per-layer probabilities make a round trip
            # through L2 (161 MB at N=77) but both kernels run near their memory roofline, which beats the fused tile kernel
            stored = [torch.empty(h, res * res, n, dtype=torch.float32, device=maps.device) for _ in logits]
            for lg, sd, pr in zip(logits, sides, stored):
                check(lib().skp_capture_store_fwd(ptr(lg), ptr(pr), h, sd, n, res, stream()), "skp_capture_store_fwd")
            check(lib().skp_collect_maps_fwd(ptr_array(stored), len(stored), h, res, n, None, 0, res, None, ptr(maps), stream()),
                  "skp_collect_maps_fwd")
        else:
            check(lib().skp_capture_mean_fwd(ptr_array(logits), int_array(sides), len(logits), ptr(maps), h, n, res, stream()),
                  "skp_capture_mean_fwd")
        ctx.save_for_backward(*logits)
        ctx.res, ctx.sides = res, sides
        return maps

    @staticmethod
    def backward(ctx, d_maps):
        logits = list(ctx.saved_tensors)
        h, _, n = logits[0].shape
        d_logits = [torch.zeros_like(l) for l in logits]
        ws = torch.empty(int(lib().skp_capture_mean_bwd_workspace(int_array(ctx.sides), len(logits), h, n, ctx.res)), dtype=torch.uint8,
                         device=d_maps.device)
        check(lib().skp_capture_mean_bwd(ptr_array(logits), int_array(ctx.sides), len(logits), ptr(_f32c(d_maps)),
                                         ptr_array(d_logits), h, n, ctx.res, ptr(ws), stream()), "skp_capture_mean_bwd")
        return (None, *d_logits)


def capture_mean(logits: Sequence[torch.Tensor], res: int) -> torch.Tensor:
    """maps[N, res, res] = mean over (layer, head) of the captured probabilities (capture + collect_maps fused)."""
    return _CaptureMean.apply(res, *logits)


# ----------------------------------------------------------------------------- collect_maps
class _CollectMaps(torch.autograd.Function):
    @staticmethod
    def forward(ctx, idx, res2: int, *stored):
        require_cuda(*stored)
        stored = [_f32c(s) for s in stored]
        bh, rr, n = stored[0].shape
        r = int(round(rr ** 0.5))
        assert r * r == rr
        for s in stored:
            assert s.shape == stored[0].shape, "stored maps must share a shape"
        t = n if idx is None else idx.numel()
        if idx is not None:
            idx = idx.to(device=stored[0].device, dtype=torch.int64).contiguous()
        r2 = r if res2 in (-1, r) else res2
        out = torch.empty(t, r2, r2, dtype=torch.float32, device=stored[0].device)
        tmp = torch.empty(t, r, r, dtype=torch.float32, device=out.device) if r2 != r else None
        check(lib().skp_collect_maps_fwd(ptr_array(stored), len(stored), bh, r, n, ptr(idx), t if idx is not None else 0, r2,
                                         ptr(tmp), ptr(out), stream()), "skp_collect_maps_fwd")
        ctx.meta = (len(stored), bh, r, n, r2, t)
        ctx.idx = idx
        return out

    @staticmethod
    def backward(ctx, d_out):
        nl, bh, r, n, r2, t = ctx.meta
        d_out = _f32c(d_out)
        d_stored = [torch.empty(bh, r * r, n, dtype=torch.float32, device=d_out.device) for _ in range(nl)]
        tmp = torch.empty(t, r, r, dtype=torch.float32, device=d_out.device) if r2 != r else None
        idx = ctx.idx
        check(lib().skp_collect_maps_bwd(ptr(d_out), nl, bh, r, n, ptr(idx), t if idx is not None else 0, r2, ptr(tmp),
                                         ptr_array(d_stored), stream()), "skp_collect_maps_bwd")
        return (None, None, *d_stored)


def collect_maps_op(stored: Sequence[torch.Tensor], upsample_res: int = -1, indices=None) -> torch.Tensor:
    return _CollectMaps.apply(indices, upsample_res, *stored)


# ----------------------------------------------------------------------------- arg-max / selection
def argmax_flat(maps: torch.Tensor) -> torch.Tensor:
    """First-occurrence flat arg-max of each [H, W] map -> int64 [T]."""
    require_cuda(maps)
    maps = _f32c(maps.detach())
    t = maps.shape[0]
    out = torch.empty(t, dtype=torch.int64, device=maps.device)
    check(lib().skp_argmax_rows(ptr(maps), t, maps[0].numel(), ptr(out), stream()), "skp_argmax_rows")
    return out


def k_argmax_flat(maps: torch.Tensor, num: int) -> torch.Tensor:
    """eval.find_k_max_pixels as flat indices [num, T]."""
    require_cuda(maps)
    maps = _f32c(maps.detach())
    t, h, w = maps.shape
    out = torch.empty(num, t, dtype=torch.int64, device=maps.device)
    check(lib().skp_k_argmax(ptr(maps), t, h, w, num, None, ptr(out), stream()), "skp_k_argmax")
    return out


def gaussian_kl_scores(maps: torch.Tensor, peaks: torch.Tensor, sigma: float, eps: float = 1e-5) -> torch.Tensor:
    require_cuda(maps, peaks)
    maps = _f32c(maps.detach())
    t, h, w = maps.shape
    kl = torch.empty(t, dtype=torch.float32, device=maps.device)
    check(lib().skp_gaussian_kl_scores(ptr(maps), t, h, w, ptr(peaks), peaks.shape[0], float(sigma), float(eps), ptr(kl),
                                       stream()), "skp_gaussian_kl_scores")
    return kl


def entropy_scores(maps: torch.Tensor) -> torch.Tensor:
    """Entropy of softmax-over-pixels of every token map (ptp_utils.py:178-181) -> fp32 [T]."""
    require_cuda(maps)
    maps = _f32c(maps.detach())
    t = maps.shape[0]
    ent = torch.empty(t, dtype=torch.float32, device=maps.device)
    check(lib().skp_entropy_scores(ptr(maps), t, maps[0].numel(), ptr(ent), stream()), "skp_entropy_scores")
    return ent


def argsort_topk(scores: torch.Tensor, top_k: int) -> torch.Tensor:
    require_cuda(scores)
    scores = _f32c(scores)
    out = torch.empty(top_k, dtype=torch.int64, device=scores.device)
    check(lib().skp_argsort_topk(ptr(scores), scores.numel(), top_k, ptr(out), stream()), "skp_argsort_topk")
    return out


def furthest_point_sampling_flat(peaks_flat: torch.Tensor, h: int, w: int, candidates: torch.Tensor, top_k: int):
    """Returns (indices int64 [top_k], count int32 [1]) on device; no host sync."""
    require_cuda(peaks_flat, candidates)
    candidates = candidates.to(device=peaks_flat.device, dtype=torch.int64).contiguous()
    out = torch.zeros(top_k, dtype=torch.int64, device=peaks_flat.device)
    n_out = torch.zeros(1, dtype=torch.int32, device=peaks_flat.device)
    check(lib().skp_furthest_point_sampling(ptr(peaks_flat), h, w, ptr(candidates), candidates.numel(), top_k, ptr(out),
                                            ptr(n_out), stream()), "skp_furthest_point_sampling")
    return out, n_out


# ----------------------------------------------------------------------------- losses
class _SharpenLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, maps, sel, peaks, sigma: float):
        require_cuda(maps, sel, peaks)
        maps = _f32c(maps)
        _, h, w = maps.shape
        loss = torch.empty((), dtype=torch.float32, device=maps.device)
        check(lib().skp_sharpen_loss_fwd(ptr(maps), h, w, ptr(sel), sel.numel(), ptr(peaks), peaks.shape[0], float(sigma),
                                         ptr(loss), stream()), "skp_sharpen_loss_fwd")
        ctx.save_for_backward(maps, sel, peaks)
        ctx.sigma = sigma
        return loss

    @staticmethod
    def backward(ctx, d_loss):
        maps, sel, peaks = ctx.saved_tensors
        _, h, w = maps.shape
        d_maps = torch.zeros_like(maps)
        check(lib().skp_sharpen_loss_bwd(ptr(maps), h, w, ptr(sel), sel.numel(), ptr(peaks), peaks.shape[0], float(ctx.sigma),
                                         ptr(_f32c(d_loss)), 1.0, ptr(d_maps), stream()), "skp_sharpen_loss_bwd")
        return d_maps, None, None, None


def sharpen_loss_op(maps: torch.Tensor, sel: torch.Tensor, sigma: float, num_subjects: int = 1) -> torch.Tensor:
    """optimize.sharpening_loss(maps[sel]) without materialising the gather or the target."""
    sel = sel.to(device=maps.device, dtype=torch.int64).contiguous()
    picked = maps.detach()[sel]
    peaks = k_argmax_flat(picked, num_subjects)
    return _SharpenLoss.apply(maps, sel, peaks, sigma)


class _EquivLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, maps, maps_t, sel, theta_inv):
        require_cuda(maps, maps_t, sel, theta_inv)
        maps, maps_t = _f32c(maps), _f32c(maps_t)
        _, h, w = maps.shape
        loss = torch.empty((), dtype=torch.float32, device=maps.device)
        check(lib().skp_equivariance_loss_fwd(ptr(maps), ptr(maps_t), h, w, ptr(sel), sel.numel(), ptr(theta_inv), ptr(loss),
                                              stream()), "skp_equivariance_loss_fwd")
        ctx.save_for_backward(maps, maps_t, sel, theta_inv)
        return loss

    @staticmethod
    def backward(ctx, d_loss):
        maps, maps_t, sel, theta_inv = ctx.saved_tensors
        _, h, w = maps.shape
        d_maps = torch.zeros_like(maps) if ctx.needs_input_grad[0] else None
        d_maps_t = torch.zeros_like(maps_t) if ctx.needs_input_grad[1] else None
        check(lib().skp_equivariance_loss_bwd(ptr(maps), ptr(maps_t), h, w, ptr(sel), sel.numel(), ptr(theta_inv),
                                              ptr(_f32c(d_loss)), 1.0, ptr(d_maps), ptr(d_maps_t), stream()),
              "skp_equivariance_loss_bwd")
        return d_maps, d_maps_t, None, None


def equivariance_loss_op(maps: torch.Tensor, maps_t: torch.Tensor, sel: torch.Tensor, theta_inv: torch.Tensor):
    """MSE(maps[sel], unwarp(maps_t[sel])) with theta_inv a [2,3] (or flat 6) fp32 device tensor."""
    sel = sel.to(device=maps.device, dtype=torch.int64).contiguous()
    theta_inv = theta_inv.to(device=maps.device, dtype=torch.float32).reshape(6).contiguous()
    return _EquivLoss.apply(maps, maps_t, sel, theta_inv)


class _AffineWarp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, theta):
        require_cuda(img, theta)
        img = _f32c(img)
        b, c, h, w = img.shape
        out = torch.empty_like(img)
        check(lib().skp_affine_warp(ptr(img), b, c, h, w, ptr(theta), ptr(out), stream()), "skp_affine_warp")
        ctx.save_for_backward(theta)
        ctx.shape = (b, c, h, w)
        return out

    @staticmethod
    def backward(ctx, d_out):
        (theta,) = ctx.saved_tensors
        b, c, h, w = ctx.shape
        d_out = _f32c(d_out)
        d_img = torch.zeros_like(d_out)
        check(lib().skp_affine_warp_bwd(ptr(d_out), b, c, h, w, ptr(theta), ptr(d_img), stream()), "skp_affine_warp_bwd")
        return d_img, None


def affine_warp(img: torch.Tensor, theta: torch.Tensor) -> torch.Tensor:
    """grid_sample(img, affine_grid(theta)), bilinear, zeros, align_corners=False; img [B,C,H,W], theta [B,2,3]."""
    b = img.shape[0]
    theta = theta.detach().to(device=img.device, dtype=torch.float32).reshape(b, 6).contiguous()
    return _AffineWarp.apply(img, theta)


def soft_argmax_(heatmaps: torch.Tensor, distance: float = 5.0) -> torch.Tensor:
    """eval.pixel_from_weighted_avg; zeroes `heatmaps` in place like the reference. Returns [T, 2]."""
    require_cuda(heatmaps)
    assert heatmaps.dtype == torch.float32 and heatmaps.is_contiguous()
    t, h, w = heatmaps.shape
    peaks = argmax_flat(heatmaps)
    out = torch.empty(t, 2, dtype=torch.float32, device=heatmaps.device)
    check(lib().skp_soft_argmax(ptr(heatmaps), t, h, w, ptr(peaks), float(distance), ptr(out), stream()), "skp_soft_argmax")
    return out


def adam_step_(param, grad, exp_avg, exp_avg_sq, step, lr=5e-3, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    """`step` is a Python int (1-based) or an int32 device tensor holding the number of steps taken so far (the kernel
    increments it): the second form is CUDA-graph capturable."""
    require_cuda(param, grad, exp_avg, exp_avg_sq)
    assert param.is_contiguous() and grad.is_contiguous() and param.dtype == torch.float32
    if isinstance(step, torch.Tensor):
        assert step.dtype == torch.int32 and step.is_cuda
        check(lib().skp_adam_step_dev(ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), param.numel(), ptr(step), lr,
                                      beta1, beta2, eps, grad_scale, stream()), "skp_adam_step_dev")
        return param
    check(lib().skp_adam_step(ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), param.numel(), int(step), lr, beta1,
                              beta2, eps, grad_scale, stream()), "skp_adam_step")
    return param
