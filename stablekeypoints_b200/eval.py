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


@torch.no_grad()
def run_image_with_context_augmented(ldm, image, context, indices, device="cuda",
                                     from_where=["down_cross", "mid_cross", "up_cross"], layers=[0, 1, 2, 3, 4, 5],
                                     augmentation_iterations=20, noise_level=-1, augment_degrees=30,
                                     augment_scale=(0.9, 1.1), augment_translate=(0.1, 0.1), visualize=False,
                                     controllers=None, num_gpus=1, save_folder="outputs", upscale_size=512, *,
                                     thetas=None, noises=None):
    """eval.py:197-355 (SURVEY "next" row f2): test-time augmentation ensemble -> [K, upscale, upscale].

    Per iteration: random affine of the image, one captured forward with the K selected tokens at `upscale_size`,
    inverse-warp of the maps and of a ones-mask accumulated (one fused kernel), finally sum/num with 0/0 -> 0.
    `thetas` ([iters,2,3]) / `noises` (list) optionally inject the randomness for reproducible parity runs;
    `visualize` (matplotlib plots) is out of scope and ignored."""
    from . import _lib, ptp_utils
    from .invertable_transform import RandomAffineWithInverse, invert_theta
    from .ops import _f32c, check, lib, ptr, stream
    dev = ldm.unet.device
    if isinstance(image, torch.Tensor):
        img = image.to(dev, torch.float32)
        img = img[None] if img.dim() == 3 else img                      # [3,H,W] like the reference's callers
    else:
        img = torch.from_numpy(image).to(dev, torch.float32).permute(2, 0, 1)[None]
    indices = torch.as_tensor(indices)
    k = indices.numel()
    num_samples = torch.zeros(k, upscale_size, upscale_size, device=dev)
    sum_samples = torch.zeros(k, upscale_size, upscale_size, device=dev)
    transform = RandomAffineWithInverse(degrees=augment_degrees, scale=augment_scale, translate=augment_translate)
    for i in range(augmentation_iterations // num_gpus):
        theta = thetas[i:i + 1] if thetas is not None else None
        augmented = transform(img, theta=theta)
        maps = ptp_utils.run_and_find_attn(ldm, augmented, context, layers=layers, noise_level=noise_level,
                                           from_where=from_where, upsample_res=upscale_size, device=device,
                                           controllers=controllers, indices=indices.cpu(),
                                           noise=None if noises is None else noises[i])[0]
        maps = _f32c(maps)
        th_inv = invert_theta(transform.last_params["theta"])[0].to(dev).reshape(6).contiguous()
        check(lib().skp_unwarp_accumulate(ptr(maps), k, upscale_size, upscale_size, ptr(th_inv), ptr(sum_samples),
                                          ptr(num_samples), stream()), "skp_unwarp_accumulate")
    out = torch.empty_like(sum_samples)
    check(lib().skp_ensemble_finalize(ptr(sum_samples), ptr(num_samples), ptr(out), out.numel(), stream()),
          "skp_ensemble_finalize")
    return out
