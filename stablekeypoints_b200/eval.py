"""Mirror of the arg-max / soft-arg-max helpers of the reference's ``unsupervised_keypoints/eval.py`` (:39-155)."""
from __future__ import annotations

import torch

from . import ops


def find_max_pixel(map):
    """eval.py:39-60: [T,h,w] -> [T,2] (row, col) of the first-occurrence arg-max, + 0.5."""
    _, h, w = map.shape
    flat = ops.argmax_flat(map)
    return torch.stack([flat // w, flat % w], dim=-1) + 0.5


def find_k_max_pixels(map, num=3):
    """eval.py:62-81: `num` successive arg-maxes, zeroing a 0.05*h radius around each; [num, T, 2]."""
    _, h, w = map.shape
    flat = ops.k_argmax_flat(map, num)
    return torch.stack([flat // w, flat % w], dim=-1) + 0.5


def mask_radius(map, max_coords, radius):
    """eval.py:83-111 (kept for API parity; the kernels apply this mask on the fly inside skp_k_argmax)."""
    t, h, w = map.shape
    ys = torch.arange(h, device=map.device).reshape(1, h, 1)
    xs = torch.arange(w, device=map.device).reshape(1, 1, w)
    d2 = (xs - max_coords[:, 1].reshape(t, 1, 1)) ** 2 + (ys - max_coords[:, 0].reshape(t, 1, 1)) ** 2
    return map * (d2 > radius ** 2).float()


def pixel_from_weighted_avg(heatmaps, distance=5):
    """eval.py:113-155 soft-arg-max; zeroes ``heatmaps`` in place beyond `distance` px of the peak like the reference."""
    if not heatmaps.is_contiguous() or heatmaps.dtype != torch.float32:
        raise ValueError("pixel_from_weighted_avg mutates its input in place: pass a contiguous fp32 tensor")
    return ops.soft_argmax_(heatmaps, float(distance))
