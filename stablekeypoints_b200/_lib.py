"""ctypes binding of libskp_b200.so (C ABI declared in include/skp_b200.h).

The product path has NO fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import Dict, List, Optional

import torch

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libskp_b200.so")
HEADER = os.path.join(PKG, "..", "include", "skp_b200.h")


class SkpError(RuntimeError):
    pass


_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> argtypes (restype is int unless listed in _RESTYPE)
_SIGNATURES: Dict[str, list] = {
    "skp_version": [],
    "skp_last_error": [],
    "skp_launch_count": [],
    "skp_gemm_nt_simt": [_P, _L, _P, _L, _P, _L, _I, _I, _I, _F, _P, _P, _L, _P],
    "skp_split_bf16": [_P, _L, _I, _I, _I, _P, _P, _P],
    "skp_gemm_nt_tc_plan": [_I, _I, _I],
    "skp_gemm_tc_force_bn": [_I],
    "skp_gemm_tc_persist": [_I],
    "skp_gemm_nt_tc": [_P, _P, _P, _P, _I, _P, _L, _I, _I, _F, _P, _P, _L, _I, _P, _P],
    "skp_im2col3x3_split": [_P, _L, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P],
    "skp_conv3x3_tc": [_P, _P, _I, _I, _I, _P, _P, _P, _L, _I, _F, _P, _P, _L, _I, _P, _P],
    "skp_gn_stats": [_P, _L, _I, _I, _I, _P, _P],
    "skp_gn_fwd": [_P, _L, _I, _I, _I, _F, _P, _P, _I, _P, _L, _P, _P, _I, _P, _P],
    "skp_gn_apply": [_P, _L, _I, _I, _I, _P, _F, _P, _P, _I, _P, _L, _P, _P, _I, _P],
    "skp_gn_im2col3x3_split": [_P, _L, _I, _I, _I, _I, _P, _F, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P],
    "skp_gn_bwd": [_P, _L, _P, _L, _I, _I, _I, _P, _F, _P, _P, _I, _P, _P, _L, _P],
    "skp_ln_split_fwd": [_P, _L, _I, _I, _P, _P, _F, _P, _P, _I, _P, _P],
    "skp_ln_bwd": [_P, _L, _P, _L, _I, _I, _P, _P, _P, _L, _P],
    "skp_geglu_split_fwd": [_P, _L, _I, _I, _P, _P, _I, _P],
    "skp_geglu_bwd": [_P, _L, _P, _L, _I, _I, _P, _L, _P],
    "skp_softmax_split_fwd": [_P, _L, _I, _I, _P, _P, _I, _P],
    "skp_cross_attn_fwd": [_P, _P, _L, _P, _L, _P, _P, _I, _I, _I, _I, _F, _P],
    "skp_cross_attn_bwd": [_P, _P, _P, _L, _P, _L, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P],
    "skp_self_attn_dp": [_I],
    "skp_self_attn_tc_workspace": [_I, _I, _I],
    "skp_self_attn_tc_fwd": [_P, _L, _P, _L, _P, _L, _P, _L, _P, _P, _I, _I, _I, _F, _P],
    "skp_self_attn_split": [_P, _L, _P, _L, _P, _L, _P, _I, _I, _I, _F, _P],
    "skp_self_attn_tc_bwd_workspace": [_I, _I, _I],
    "skp_self_attn_tc_bwd": [_P, _L, _P, _L, _P, _P, _L, _P, _L, _P, _L, _P, _P, _L, _P, _L, _P, _L, _I, _I, _I, _F, _P],
    "skp_self_attn_fwd": [_P, _L, _P, _L, _P, _L, _P, _L, _P, _P, _I, _I, _I, _F, _P],
    "skp_self_attn_bwd": [_P, _L, _P, _L, _P, _P, _P, _P, _P, _L, _P, _L, _P, _L, _I, _I, _I, _F, _P],
    "skp_xattn_tc_workspace": [_I, _I, _I, _I],
    "skp_xattn_tc_fwd": [_P, _L, _P, _L, _P, _L, _P, _L, _P, _P, _P, _I, _I, _I, _I, _F, _P],
    "skp_cross_attn_split": [_P, _L, _P, _L, _P, _L, _P, _P, _I, _I, _I, _I, _F, _P],
    "skp_cross_attn_tc_fwd": [_P, _L, _P, _L, _P, _L, _P, _L, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P],
    "skp_cross_attn_tc_bwd": [_P, _L, _P, _L, _P, _P, _P, _P, _P, _P, _P, _L, _P, _L, _P, _L, _I, _I, _I, _I, _F, _P],
    "skp_capture_select": [_I, _I],
    "skp_capture_tc": [_I],
    "skp_capture_tc_trace": [_P],
    "skp_capture_tc_ok": [_P, _I, _I, _I, _I],
    "skp_capture_tc_workspace": [_P, _I, _I],
    "skp_capture_store_tc_fwd": [_P, _P, _I, _I, _I, _I, _P, _P],
    "skp_capture_mean_tc_fwd": [_P, _P, _I, _P, _I, _I, _I, _P, _P],
    "skp_capture_store_fwd": [_P, _P, _I, _I, _I, _I, _P],
    "skp_capture_store_bwd": [_P, _P, _P, _I, _I, _I, _I, _P],
    "skp_capture_mean_fwd": [_P, _P, _I, _P, _I, _I, _I, _P],
    "skp_capture_mean_bwd_workspace": [_P, _I, _I, _I, _I],
    "skp_capture_mean_bwd": [_P, _P, _I, _P, _P, _I, _I, _I, _P, _P],
    "skp_collect_maps_fwd": [_P, _I, _I, _I, _I, _P, _I, _I, _P, _P, _P],
    "skp_collect_maps_bwd": [_P, _I, _I, _I, _I, _P, _I, _I, _P, _P, _P],
    "skp_argmax_rows": [_P, _I, _I, _P, _P],
    "skp_k_argmax": [_P, _I, _I, _I, _I, _P, _P, _P],
    "skp_gaussian_kl_scores": [_P, _I, _I, _I, _P, _I, _F, _F, _P, _P],
    "skp_entropy_scores": [_P, _I, _I, _P, _P],
    "skp_argsort_topk": [_P, _I, _I, _P, _P],
    "skp_furthest_point_sampling": [_P, _I, _I, _P, _I, _I, _P, _P, _P],
    "skp_sharpen_loss_fwd": [_P, _I, _I, _P, _I, _P, _I, _F, _P, _P],
    "skp_sharpen_loss_bwd": [_P, _I, _I, _P, _I, _P, _I, _F, _P, _F, _P, _P],
    "skp_equivariance_loss_fwd": [_P, _P, _I, _I, _P, _I, _P, _P, _P],
    "skp_equivariance_loss_bwd": [_P, _P, _I, _I, _P, _I, _P, _P, _F, _P, _P, _P],
    "skp_affine_warp": [_P, _I, _I, _I, _I, _P, _P, _P],
    "skp_affine_warp_bwd": [_P, _I, _I, _I, _I, _P, _P, _P],
    "skp_soft_argmax": [_P, _I, _I, _I, _P, _F, _P, _P],
    "skp_unwarp_accumulate": [_P, _I, _I, _I, _P, _P, _P, _P],
    "skp_ensemble_finalize": [_P, _P, _P, _L, _P],
    "skp_adam_step": [_P, _P, _P, _P, _L, _I, _F, _F, _F, _F, _F, _P],
    "skp_adam_step_dev": [_P, _P, _P, _P, _L, _P, _F, _F, _F, _F, _F, _P],
}
_RESTYPE = {"skp_last_error": C.c_char_p, "skp_launch_count": C.c_int64, "skp_gemm_tc_force_bn": None, "skp_gemm_tc_persist": None, "skp_capture_select": None, "skp_capture_tc": None, "skp_capture_tc_trace": None,
            "skp_self_attn_tc_workspace": C.c_int64, "skp_self_attn_tc_bwd_workspace": C.c_int64,
            "skp_capture_tc_workspace": C.c_int64, "skp_xattn_tc_workspace": C.c_int64,
            "skp_capture_mean_bwd_workspace": C.c_int64}


def declared_symbols() -> List[str]:
    """Every function name declared in include/skp_b200.h."""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(skp_[a-z0-9_]+)\s*\(", text)))


_lib: Optional[C.CDLL] = None
_profile: Optional[list] = None  # when a list: (name, start_event, end_event) per C-ABI call (bench.py kernel shares)


class _Profiled:
    """Proxy that brackets every kernel-launching C-ABI call with CUDA events on the current stream."""

    def __init__(self, handle):
        self._h = handle

    def __getattr__(self, name):
        fn = getattr(self._h, name)
        if _profile is None or name in ("skp_last_error", "skp_version", "skp_launch_count"):
            return fn

        def wrapped(*args):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            rc = fn(*args)
            e.record()
            _profile.append((name, s, e, _shape_key(name, args)))
            return rc

        return wrapped


def _shape_key(name, args):
    """Problem shape of the GEMM-like entry points (scripts/profile_step.py --shapes)."""
    if name == "skp_gemm_nt_tc":
        return "M%d N%d K%d split%d" % (args[7], args[8], args[4], args[13])
    if name == "skp_conv3x3_tc":
        return "H%d W%d Cin%d Cout%d split%d" % (args[2], args[3], args[4], args[9], args[14])
    if name in ("skp_self_attn_fwd", "skp_self_attn_bwd"):
        return "S%d h%d d%d" % (args[-5], args[-4], args[-3])
    if name in ("skp_cross_attn_tc_fwd", "skp_cross_attn_tc_bwd"):
        return "S%d N%d h%d d%d" % (args[-6], args[-5], args[-4], args[-3])
    if name in ("skp_gn_stats", "skp_gn_apply"):
        return "rows%d C%d" % (args[2], args[3])
    if name == "skp_gn_bwd":
        return "rows%d C%d" % (args[4], args[5])
    if name == "skp_split_bf16":
        return "rows%d cols%d" % (args[2], args[3])
    if name == "skp_ln_split_fwd":
        return "rows%d C%d" % (args[2], args[3])
    return ""


def start_profile() -> None:
    global _profile
    _profile = []


def stop_profile(by_shape: bool = False) -> Dict[str, list]:
    """Returns {kernel entry point: [ms per call]} (keys get the problem shape appended when by_shape) and disables
    profiling."""
    global _profile
    rec, _profile = _profile or [], None
    torch.cuda.synchronize()
    out: Dict[str, list] = {}
    for name, s, e, key in rec:
        out.setdefault(f"{name} {key}".strip() if by_shape else name, []).append(s.elapsed_time(e))
    return out


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SkpError(f"{LIB_PATH} not built: run `python -m stablekeypoints_b200.build` (no CPU fallback exists)")
        handle = C.CDLL(LIB_PATH)
        for name, args in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = args
            fn.restype = _RESTYPE.get(name, C.c_int)
        _lib = handle
    return _Profiled(_lib) if _profile is not None else _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise SkpError(f"{what} failed with status {rc}: {lib().skp_last_error().decode()}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise SkpError("stablekeypoints_b200 kernels need CUDA tensors (the product has no CPU path)")


def launch_count() -> int:
    return int(lib().skp_launch_count())


def ptr_array(tensors) -> "C.Array":
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


def int_array(values) -> "C.Array":
    arr = (C.c_int * len(values))()
    for i, v in enumerate(values):
        arr[i] = int(v)
    return arr
