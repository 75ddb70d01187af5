"""Drop-in mirror of the live part of the reference's ``unsupervised_keypoints/ptp_utils.py`` on the B200 kernels.

Same names, argument meaning and side effects as the reference (cited per function); the tensor math is the
hand-written CUDA of libskp_b200 (see ops.py) and the model underneath is sd15_engine.UNetEngine.
"""
from __future__ import annotations

import abc
import os
from typing import Dict, List, Optional

import numpy as np
import torch

from . import ops
from .eval import find_k_max_pixels, find_max_pixel  # noqa: F401  (re-exported like the reference does)


class AttentionControl(abc.ABC):
    """ptp_utils.py:32-60."""

    def step_callback(self, x_t):
        return x_t

    def between_steps(self):
        return

    @property
    def num_uncond_att_layers(self):
        return 0

    @abc.abstractmethod
    def forward(self, dict, is_cross: bool, place_in_unet: str):
        raise NotImplementedError

    def __call__(self, dict, is_cross: bool, place_in_unet: str):
        dict = self.forward(dict, is_cross, place_in_unet)
        return dict["attn"]

    def reset(self):
        self.cur_step = 0
        self.cur_att_layer = 0

    def __init__(self):
        self.cur_step = 0
        self.num_att_layers = -1
        self.cur_att_layer = 0


class AttentionStore(AttentionControl):
    """ptp_utils.py:63-83: ``step_store["attn"]`` is a list of [B*h, R*R, N] tensors; ``reset()`` empties it."""

    @staticmethod
    def get_empty_store():
        return {"attn": []}

    def forward(self, dict, is_cross: bool, place_in_unet: str):
        self.step_store["attn"].append(dict["attn"])
        return dict

    def reset(self):
        super().reset()
        self.step_store = self.get_empty_store()
        self._skp_logits = None

    def __init__(self):
        super().__init__()
        self.step_store = self.get_empty_store()
        self._skp_logits = None  # fused mode: low-res logits of the captured layers instead of [h,R*R,N] tensors


# ----------------------------------------------------------------------------- token selection
def find_top_k_gaussian(attention_maps, top_k, sigma=3, epsilon=1e-5, num_subjects=1):
    """ptp_utils.py:86-112: tokens whose map is closest (KL) to a Gaussian at its own arg-max; [top_k] int64."""
    peaks = ops.k_argmax_flat(attention_maps, num_subjects)
    kl = ops.gaussian_kl_scores(attention_maps, peaks, sigma, epsilon)
    return ops.argsort_topk(kl, top_k)


def furthest_point_sampling(attention_maps, top_k, top_initial_candidates):
    """ptp_utils.py:115-159, one kernel instead of a Python double loop with a .item() per pair."""
    _, h, w = attention_maps.shape
    peaks = ops.argmax_flat(attention_maps)
    out, n_out = ops.furthest_point_sampling_flat(peaks, h, w, top_initial_candidates, top_k)
    if top_initial_candidates.numel() < top_k:
        # the reference returns fewer indices when the candidates run out (ptp_utils.py:139-159).  That needs the count on
        # the host, which a CUDA-graph capture cannot give: refuse instead of returning zero-padded indices.
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("furthest_point_sampling: fewer candidates (%d) than top_k (%d) cannot be captured in a CUDA graph"
                               % (top_initial_candidates.numel(), top_k))
        return out[: int(n_out.item())]
    return out


def entropy_sort(attention_maps, top_k, min_dist=0.05):
    """ptp_utils.py:165-187 (non-default --top_k_strategy entropy): ascending entropy of softmax-over-pixels."""
    return ops.argsort_topk(ops.entropy_scores(attention_maps), top_k)


# ----------------------------------------------------------------------------- capture registration
def register_attention_control(model, controller, feature_upsample_res=256):
    """ptp_utils.py:472-573.  The reference monkey-patches every CrossAttention under ``up_blocks``; the engine has
    that hook built in (sd15_engine.UNetEngine._cross_attention), so registration only binds the controller."""
    unet = getattr(model, "module", model)
    count = sum(2 for l in unet.cross_layers if l.in_up)  # attn1 + attn2 per up-block transformer (18 for SD1.x)
    assert count != 0, "No cross attention layers found in the model. Please check to make sure you're using diffusers==0.8.0."
    unet.controller = controller
    unet.feature_upsample_res = feature_upsample_res
    if controller is not None:
        controller.num_att_layers = count


# ----------------------------------------------------------------------------- per-image driver
def image2latent(model, image, device):
    """ptp_utils.py:289-304: (img*2-1) NCHW -> vae.encode(...).latent_dist.mean * 0.18215 (no grad).
    numpy NHWC or torch NCHW images in [0,1]; 4-D latents pass through like the reference."""
    with torch.no_grad():
        dev = model.unet.device
        if isinstance(image, torch.Tensor) and image.dim() == 4 and image.shape[1] == 4:
            return image.to(dev)
        if isinstance(image, np.ndarray):
            image = torch.from_numpy(np.ascontiguousarray(image)).permute(0, 3, 1, 2)
        image = image.to(dev, torch.float32, non_blocking=True) * 2 - 1
        return model.vae.encode(image)["latent_dist"].mean * 0.18215


def find_pred_noise(ldm, image, context, noise_level=-1, device="cuda", noise=None):
    """ptp_utils.py:205-231.  ``noise`` (optional, not in the reference) injects the Gaussian noise for seed-free
    parity runs; otherwise it is drawn with randn_like on the device exactly where the reference draws it."""
    latent = image2latent(ldm, image, device)
    if noise is None:
        noise = torch.randn_like(latent)
    else:
        noise = noise.to(latent.device, latent.dtype)
    t = ldm.scheduler.timesteps[noise_level]
    noisy = ldm.scheduler.add_noise(latent, noise, t)
    b = noisy.shape[0]
    # the reference passes context.repeat(B,1,1) (ptp_utils.py:229); at B == 1 (one image per rank) that is the tensor itself,
    # and handing the leaf over keeps the engine's K|V projection cache valid for both forwards of an iteration
    ctx = context if (b == 1 and os.environ.get("SKP_CTX_REPEAT", "0") != "1") else context.repeat(b, 1, 1)
    pred = ldm.unet(noisy, t.repeat(b), ctx)["sample"]
    return noise, pred


# run_and_find_attn drops pred_noise (ptp_utils.py:246), so nothing observable depends on the ~40 % of the UNet forward
# that follows the 4th captured layer: the engine stops there.  find_pred_noise itself always runs the full forward.
# SKP_EARLY_EXIT=0 keeps the full forward inside run_and_find_attn too (A/B measurements).
EARLY_EXIT = os.environ.get("SKP_EARLY_EXIT", "1") != "0"


def _fused_ok(ldm, controllers, upsample_res, indices) -> bool:
    if os.environ.get("SKP_CAPTURE_MODE", "fused") == "store":
        return False
    # upsample_res == R is the bilinear identity, so Stage 2 (keypoint_regressor.py:70-80) fuses too
    same_res = upsample_res in (-1, ldm.unet.feature_upsample_res)
    return (same_res and indices is None and len(controllers) == 1
            and all(type(c) is AttentionStore for c in controllers.values()))


def run_and_find_attn(ldm, image, context, noise_level=-1, device="cuda",
                      from_where=["down_cross", "mid_cross", "up_cross"], layers=[0, 1, 2, 3, 4, 5], upsample_res=32,
                      indices=None, controllers=None, noise=None):
    """ptp_utils.py:234-272: one captured forward, then collect_maps + reset per controller.

    Training shape (upsample_res=-1, no indices): the capture and the (layer, head) mean are fused -- the
    [h, R*R, N] probability tensors are never written (skp_capture_mean_*).  Otherwise the engine materialises
    them through the controller exactly like the reference and skp_collect_maps_* aggregates."""
    from .optimize import collect_maps
    unet = ldm.unet
    fused = _fused_ok(ldm, controllers, upsample_res, indices)
    prev_mode, prev_exit = unet.capture_mode, unet.early_exit
    unet.capture_mode = "fused" if fused else "store"
    unet.early_exit = prev_exit or (EARLY_EXIT and len(controllers) == 1)
    try:
        find_pred_noise(ldm, image, context, noise_level=noise_level, device=device, noise=noise)
    finally:
        unet.capture_mode, unet.early_exit = prev_mode, prev_exit
    attention_maps = []
    for key in controllers:
        ctl = controllers[key]
        if fused:
            picked = [l for i, l in enumerate(unet.last_logits) if i in layers]
            attention_maps.append(ops.capture_mean(picked, unet.feature_upsample_res))
            ctl.reset()
        else:
            attention_maps.append(collect_maps(ctl, from_where=from_where, upsample_res=upsample_res, layers=layers,
                                               indices=indices))
        ctl.reset()
    return attention_maps


def init_random_noise(device, num_words=77):
    """ptp_utils.py:649-650."""
    return torch.randn(1, num_words, 768).to(device)
