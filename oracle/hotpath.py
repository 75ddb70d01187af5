"""CPU fp32 restatement of the reference's per-image hot path -- TEST INFRASTRUCTURE ONLY.

Every function cites the reference lines (paths relative to /root/reference/unsupervised_keypoints)
whose arithmetic it restates.  Written against plain torch-CPU fp32 so it travels to the GPU box
(where /root/reference does not exist) and serves as (1) the parity checker for the CUDA path and
(2) the "port" CPU baseline of bench.py.  It is pinned against the reference's real code by
tests/test_oracle_vs_golden.py (committed fixtures minted with the reference imported in the
build container) and tests/test_oracle_vs_reference.py (live, when the tree is present).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- store (a2)
class AttentionStore:
    """ptp_utils.py:32-83 -- controller(dict, is_cross, place) appends dict["attn"]; reset() empties."""

    def __init__(self):
        self.cur_step = 0
        self.num_att_layers = -1
        self.cur_att_layer = 0
        self.step_store = {"attn": []}

    def __call__(self, payload, is_cross: bool, place_in_unet: str):
        self.step_store["attn"].append(payload["attn"])
        return payload["attn"]

    def reset(self):
        self.cur_step = 0
        self.cur_att_layer = 0
        self.step_store = {"attn": []}


# ----------------------------------------------------------------------------- hook (a3)
def _split_heads(t: torch.Tensor, h: int) -> torch.Tensor:
    b, s, c = t.shape
    return t.reshape(b, s, h, c // h).transpose(1, 2).reshape(b * h, s, c // h)


def _merge_heads(t: torch.Tensor, h: int) -> torch.Tensor:
    bh, s, d = t.shape
    return t.reshape(bh // h, h, s, d).transpose(1, 2).reshape(bh // h, s, h * d)


def attention_with_capture(mod, x, context, controller, upsample_res: int, max_stored: int = 4):
    """ptp_utils.py:480-541.  ``mod`` exposes to_q/to_k/to_v/to_out/heads/scale.

    Regular attention (:483-506); for cross layers with S <= 32**2 while fewer than 4 maps are stored
    (:508-512) the LayerNorm'd input is bicubic-upsampled to upsample_res**2 (:513-529), re-projected
    by to_q, and softmax(q' k^T scale) over the TOKEN axis is handed to the controller (:531-538).
    """
    b, s, c = x.shape
    h = mod.heads
    is_cross = context is not None
    src = context if is_cross else x
    q = _split_heads(mod.to_q(x), h)
    k = _split_heads(mod.to_k(src), h)
    v = _split_heads(mod.to_v(src), h)
    probs = torch.softmax(torch.bmm(q, k.transpose(1, 2)) * mod.scale, dim=-1)
    out = torch.bmm(probs, v)
    if is_cross and s <= 32 ** 2 and len(controller.step_store["attn"]) < max_stored:
        side = int(s ** 0.5)
        grid = x.reshape(b, side, side, c).permute(0, 3, 1, 2)
        up = F.interpolate(grid, size=(upsample_res, upsample_res), mode="bicubic", align_corners=False)
        up = up.permute(0, 2, 3, 1).reshape(b, upsample_res * upsample_res, c)
        q_up = _split_heads(mod.to_q(up), h)
        stored = torch.softmax(torch.bmm(q_up, k.transpose(1, 2)) * mod.scale, dim=-1)
        controller({"attn": stored}, is_cross, "up")
    to_out = mod.to_out[0] if isinstance(mod.to_out, torch.nn.ModuleList) else mod.to_out
    return to_out(_merge_heads(out, h))


def register_capture(unet, controller, upsample_res: int) -> int:
    """ptp_utils.py:472-573 -- patch every module whose class is named CrossAttention under the
    top-level children whose name contains "up"; returns the count (asserted non-zero, :573)."""
    count = 0
    for name, child in unet.named_children():
        if "up" not in name:
            continue
        for mod in child.modules():
            if mod.__class__.__name__ == "CrossAttention":
                def fwd(x, context=None, mask=None, _m=mod):
                    return attention_with_capture(_m, x, context, controller, upsample_res)
                mod.forward = fwd
                count += 1
    controller.num_att_layers = count
    assert count != 0, "no CrossAttention modules under *up* children"
    return count


# ----------------------------------------------------------------------------- aggregation (a7)
def collect_maps(controller, upsample_res: int = 512, layers: Sequence[int] = (0, 1, 2, 3), indices=None):
    """optimize.py:27-79.  [B*h, R*R, N] per layer -> optional token gather -> [B*h, N', R, R] ->
    bilinear (align_corners=False) when upsample_res != -1 (guard at :63 compares sqrt(#tokens) with the
    target, i.e. fires whenever they differ) -> mean over (layer, B*h) -> [N', R', R']; resets."""
    picked = []
    for li, data in enumerate(controller.step_store["attn"]):
        if li not in layers:
            continue
        bh, rr, n = data.shape
        side = int(rr ** 0.5)
        data = data.reshape(bh, side, side, n)
        if indices is not None:
            data = data[..., indices]
        data = data.permute(0, 3, 1, 2)
        if upsample_res != -1 and data.shape[1] ** 0.5 != upsample_res:
            data = F.interpolate(data, size=(upsample_res, upsample_res), mode="bilinear", align_corners=False)
        picked.append(data)
    out = torch.stack(picked, 0).mean(dim=(0, 1))
    controller.reset()
    return out


# ----------------------------------------------------------------------------- arg-max helpers (a8/a13)
def find_max_pixel(maps: torch.Tensor) -> torch.Tensor:
    """eval.py:39-60 -- first-occurrence arg-max per map, (row, col) + 0.5, float."""
    t, h, w = maps.shape
    flat = torch.argmax(maps.reshape(t, h * w), dim=-1)
    return torch.stack([flat // w, flat % w], dim=-1) + 0.5


def mask_radius(maps: torch.Tensor, centres: torch.Tensor, radius: float) -> torch.Tensor:
    """eval.py:83-111 -- multiply by 0 inside the radius (integer pixel coords vs the +0.5 centre)."""
    t, h, w = maps.shape
    ys = torch.arange(h).reshape(1, h, 1)
    xs = torch.arange(w).reshape(1, 1, w)
    d2 = (xs - centres[:, 1].reshape(t, 1, 1)) ** 2 + (ys - centres[:, 0].reshape(t, 1, 1)) ** 2
    return maps * (d2 > radius ** 2).float()


def find_k_max_pixels(maps: torch.Tensor, num: int = 3) -> torch.Tensor:
    """eval.py:62-81 -- `num` successive arg-maxes, masking 0.05*h around each; [num, T, 2]."""
    pts = []
    for _ in range(num):
        p = find_max_pixel(maps)
        pts.append(p)
        maps = mask_radius(maps, p, 0.05 * maps.shape[1])
    return torch.stack(pts)


def gaussian_circles(pos: torch.Tensor, size: int, sigma: float) -> torch.Tensor:
    """optimize_token.py:203-241 -- pos [num, T, 2] in [0,1] (row, col); un-normalised
    exp(-d^2 / (2 sigma^2)) on the +0.5 pixel-centre grid, averaged over `num`."""
    centre = pos * size  # [num, T, 2]
    rows = (torch.arange(size).float() + 0.5).reshape(1, 1, size, 1)
    cols = (torch.arange(size).float() + 0.5).reshape(1, 1, 1, size)
    d2 = (cols - centre[..., 1][..., None, None]) ** 2 + (rows - centre[..., 0][..., None, None]) ** 2
    return torch.exp(-1 * d2 / (2.0 * sigma ** 2.0)).mean(dim=0)


def pixel_from_weighted_avg(heatmaps: torch.Tensor, distance: float = 5) -> torch.Tensor:
    """eval.py:113-155 -- zero (IN PLACE) everything further than `distance` px from the arg-max,
    normalise by (sum + 1e-6), expectation of (row, col), + 0.5."""
    t, m, n = heatmaps.shape
    if distance != -1:
        peak = find_max_pixel(heatmaps).long()
        rows = torch.arange(m).float().reshape(1, m, 1)
        cols = torch.arange(n).float().reshape(1, 1, n)
        far = torch.sqrt((rows - peak[:, 0].reshape(t, 1, 1)) ** 2 + (cols - peak[:, 1].reshape(t, 1, 1)) ** 2)
        heatmaps[far > distance] = 0.0
    norm = heatmaps / (heatmaps.sum(dim=(1, 2), keepdim=True) + 1e-6)
    rows = torch.arange(m).float().reshape(1, m, 1)
    cols = torch.arange(n).float().reshape(1, 1, n)
    return torch.stack([(rows * norm).sum(dim=(1, 2)), (cols * norm).sum(dim=(1, 2))], dim=-1) + 0.5


# ----------------------------------------------------------------------------- token selection (a8/a9)
def gaussian_kl_scores(maps: torch.Tensor, sigma: float, epsilon: float = 1e-5, num_subjects: int = 1):
    """ptp_utils.py:97-108 -- KL(target || softmax_pixels(map + eps)) per token."""
    t, h, w = maps.shape
    locs = find_k_max_pixels(maps, num=num_subjects) / h
    p = torch.softmax(maps.reshape(t, h * w) + epsilon, dim=-1)
    tgt = gaussian_circles(locs, size=h, sigma=sigma).reshape(t, h * w) + epsilon
    tgt = tgt / tgt.sum(dim=-1, keepdim=True)
    return torch.sum(tgt * (torch.log(tgt) - torch.log(p)), dim=-1)


def find_top_k_gaussian(maps: torch.Tensor, top_k: int, sigma: float = 3, epsilon: float = 1e-5,
                        num_subjects: int = 1) -> torch.Tensor:
    """ptp_utils.py:86-112 -- ascending arg-sort of the KL scores, first top_k."""
    return torch.argsort(gaussian_kl_scores(maps, sigma, epsilon, num_subjects), dim=-1, descending=False)[:top_k]


def entropy_sort(maps: torch.Tensor, top_k: int) -> torch.Tensor:
    """ptp_utils.py:165-187 -- ascending entropy of softmax-over-pixels (Categorical(probs).entropy()), first top_k."""
    t = maps.shape[0]
    p = torch.softmax(maps.reshape(t, -1), dim=-1)
    ent = torch.distributions.Categorical(probs=p).entropy()
    return torch.argsort(ent, dim=-1, descending=False)[:top_k]


def furthest_point_sampling(maps: torch.Tensor, top_k: int, candidates: torch.Tensor) -> torch.Tensor:
    """ptp_utils.py:115-159 -- furthest pair among candidates (strict '>' so the first maximum in
    (i<j) lexicographic order wins, :135), then greedy max-min-distance (strict '>', :152)."""
    loc = (find_max_pixel(maps) / maps.shape[1]).detach().numpy().astype(np.float32)
    cand = [int(c) for c in candidates]

    def dist(a, b):
        d = loc[a] - loc[b]
        return np.sqrt(np.float32(d[0] * d[0]) + np.float32(d[1] * d[1]), dtype=np.float32)

    best, pair = np.float32(-1), None
    for i in range(len(cand)):
        for j in range(i + 1, len(cand)):
            d = dist(cand[i], cand[j])
            if d > best:
                best, pair = d, (cand[i], cand[j])
    chosen = [pair[0], pair[1]]
    for _ in range(top_k - 2):
        best, pick = np.float32(-1), None
        for c in cand:
            if c in chosen:
                continue
            m = min(dist(c, s) for s in chosen)
            if m > best:
                best, pick = m, c
        if pick is not None:
            chosen.append(pick)
    return torch.tensor(chosen)


# ----------------------------------------------------------------------------- augmentation (a11)
def affine_theta(angle_deg: float, scale: float, tx: float, ty: float) -> torch.Tensor:
    """invertable_transform.py:21-36 -- scale * [[cos, sin], [-sin, cos]] | (tx, ty); [1,2,3]."""
    a = math.radians(angle_deg)
    th = torch.tensor([[math.cos(a), math.sin(a), tx], [-math.sin(a), math.cos(a), ty]], dtype=torch.float)
    th[:, :2] = th[:, :2] * scale
    return th[None]


def sample_affine_params(batch: int, degrees=15.0, scale=(0.8, 1.0), translate=(0.25, 0.25)):
    """invertable_transform.py:42-57 -- four torch.rand(1) draws per image, in this order."""
    thetas = []
    for _ in range(batch):
        ang = torch.rand(1).item() * (2 * degrees) - degrees
        sc = torch.rand(1).item() * (scale[1] - scale[0]) + scale[0]
        tx = torch.rand(1).item() * (2 * translate[0]) - translate[0]
        ty = torch.rand(1).item() * (2 * translate[1]) - translate[1]
        thetas.append(affine_theta(ang, sc, tx, ty))
    return torch.cat(thetas, 0)


def affine_warp(img: torch.Tensor, theta: torch.Tensor) -> torch.Tensor:
    """invertable_transform.py:65-68 -- affine_grid + bilinear grid_sample, zeros, align_corners=False."""
    grid = F.affine_grid(theta, list(img.shape), align_corners=False)
    return F.grid_sample(img, grid, align_corners=False)


def affine_unwarp(img: torch.Tensor, theta: torch.Tensor) -> torch.Tensor:
    """invertable_transform.py:72-92 -- same sampler with inverse([theta; 0 0 1])[:2]."""
    full = torch.cat([theta, torch.tensor([[[0.0, 0.0, 1.0]]]).expand(theta.shape[0], -1, -1)], dim=1)
    return affine_warp(img, torch.inverse(full)[:, :2, :])


# ----------------------------------------------------------------------------- losses (a10/a11)
def sharpening_loss(maps: torch.Tensor, sigma: float = 1.0, num_subjects: int = 1) -> torch.Tensor:
    """optimize.py:166-206 -- MSE(map, Gaussian of peak 1 centred on the map's own arg-max)."""
    pos = find_k_max_pixels(maps, num=num_subjects) / maps.shape[-1]
    target = gaussian_circles(pos, size=maps.shape[1], sigma=sigma)
    return F.mse_loss(maps, target)


def equivariance_loss(maps: torch.Tensor, maps_t: torch.Tensor, theta: torch.Tensor, index: int) -> torch.Tensor:
    """optimize.py:157-163 -- MSE(map, un-warped transformed map), zero-padded pixels included.
    ``maps_t`` is [G, K, R, R] (the caller repeats it over the device count, optimize.py:399)."""
    return F.mse_loss(maps, affine_unwarp(maps_t, theta)[index])


# ----------------------------------------------------------------------------- per-image driver (a4/a6)
def encode_image(ldm, image) -> torch.Tensor:
    """ptp_utils.py:289-304 -- (img*2-1) NCHW -> vae.encode(...).latent_dist.mean * 0.18215, no grad."""
    with torch.no_grad():
        if isinstance(image, torch.Tensor):
            image = image.permute(0, 2, 3, 1).detach().cpu().numpy()  # ptp_utils.py:213-214 round trip
        x = (torch.from_numpy(np.ascontiguousarray(image)).float() * 2 - 1).permute(0, 3, 1, 2)
        return ldm.vae.encode(x)["latent_dist"].mean * 0.18215


def find_pred_noise(ldm, image, context, noise_level: int = -1, noise: Optional[torch.Tensor] = None):
    """ptp_utils.py:205-231.  ``noise`` may be injected for seed-free parity (reference: randn_like)."""
    latent = encode_image(ldm, image)
    if noise is None:
        noise = torch.randn_like(latent)
    t = ldm.scheduler.timesteps[noise_level]
    noisy = ldm.scheduler.add_noise(latent, noise, t)
    pred = ldm.unet(noisy, t.repeat(noisy.shape[0]), context.repeat(noisy.shape[0], 1, 1))["sample"]
    return noise, pred


def run_and_find_attn(ldm, image, context, controllers: Dict, noise_level: int = -1, layers=(0, 1, 2, 3),
                      upsample_res: int = -1, indices=None, noise=None) -> List[torch.Tensor]:
    """ptp_utils.py:234-272 -- one captured forward, then collect_maps + reset per controller."""
    find_pred_noise(ldm, image, context, noise_level, noise)
    out = []
    for key in controllers:
        out.append(collect_maps(controllers[key], upsample_res=upsample_res, layers=layers, indices=indices))
        controllers[key].reset()
    return out


def load_oracle_ldm(pipeline, feature_upsample_res: int):
    """optimize_token.py:41-69 (CPU branch) -- one store for the cpu device; a forward-pre-hook that
    re-registers the capture on every UNet call."""
    controllers = {torch.device("cpu"): AttentionStore()}

    def pre_hook(module, inputs):
        register_capture(module, controllers[inputs[0].device], feature_upsample_res)

    pipeline.unet.register_forward_pre_hook(pre_hook)
    return pipeline, controllers, 1


# ----------------------------------------------------------------------------- Stage-1 loop body (a12)
def stage1_losses(maps: torch.Tensor, maps_t: torch.Tensor, theta: torch.Tensor, *, top_k: int = 10,
                  num_candidates: int = 25, sigma: float = 2.0, num_subjects: int = 1, forced_indices=None):
    """optimize.py:380-401 for one device: candidate tokens by Gaussian-KL on the ORIGINAL maps,
    furthest-point sampling on the TRANSFORMED maps' arg-maxes, then the two losses."""
    if forced_indices is None:
        cand = find_top_k_gaussian(maps, num_candidates, sigma=sigma, num_subjects=num_subjects)
        idx = furthest_point_sampling(maps_t, top_k, cand)
    else:
        idx = torch.as_tensor(forced_indices)
    sharp = sharpening_loss(maps[idx], sigma=sigma, num_subjects=num_subjects)
    equiv = equivariance_loss(maps[idx], maps_t[idx][None], theta, 0)
    return idx, sharp, equiv


def stage1_iteration(ldm, controllers, image, context, theta, noise_a=None, noise_b=None, *, top_k=10,
                     num_candidates=25, sigma=2.0, sharpening_loss_weight=100.0,
                     equivariance_attn_loss_weight=1000.0, accum: int = 1, layers=(0, 1, 2, 3),
                     forced_indices=None):
    """optimize.py:341-422 for G=1: two captured forwards (fresh noise each, ptp_utils.py:219),
    selection, loss = 1000*equiv + 100*sharp, / (B//G), backward into ``context``."""
    maps = run_and_find_attn(ldm, image, context, controllers, layers=layers, noise=noise_a)[0]
    image_t = affine_warp(image, theta)
    maps_t = run_and_find_attn(ldm, image_t, context, controllers, layers=layers, noise=noise_b)[0]
    idx, sharp, equiv = stage1_losses(maps, maps_t, theta, top_k=top_k, num_candidates=num_candidates,
                                      sigma=sigma, forced_indices=forced_indices)
    loss = (equiv * equivariance_attn_loss_weight + sharp * sharpening_loss_weight) / accum
    loss.backward()
    return {"maps": maps.detach(), "maps_t": maps_t.detach(), "indices": idx, "sharp": sharp.detach(),
            "equiv": equiv.detach(), "loss": loss.detach()}


def adam_step(param, grad, m, v, step: int, lr=5e-3, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam defaults as used at optimize.py:320,424 (no weight decay, no amsgrad)."""
    m.mul_(b1).add_(grad, alpha=1 - b1)
    v.mul_(b2).addcmul_(grad, grad, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    param.addcdiv_(m, denom, value=-lr / bc1)
    return param


# ----------------------------------------------------------------------------- synthetic data (8d)
def synthetic_image(seed: int = 1, size: int = 512, blobs: int = 12, batch: int = 1) -> torch.Tensor:
    """[B,3,size,size] in [0,1]: sums of Gaussian blobs + low-frequency noise."""
    g = torch.Generator().manual_seed(seed)
    ys = (torch.arange(size).float() + 0.5).reshape(1, size, 1) / size
    xs = (torch.arange(size).float() + 0.5).reshape(1, 1, size) / size
    out = []
    for _ in range(batch):
        img = torch.zeros(3, size, size)
        for _ in range(blobs):
            cy, cx = torch.rand(2, generator=g).tolist()
            s = 0.03 + 0.08 * torch.rand(1, generator=g).item()
            col = torch.rand(3, 1, 1, generator=g)
            img += col * torch.exp(-((ys - cy) ** 2 + (xs - cx) ** 2) / (2 * s * s))
        low = F.interpolate(torch.rand(1, 3, 8, 8, generator=g), size=(size, size), mode="bilinear",
                            align_corners=False)[0]
        out.append((0.7 * img + 0.3 * low).clamp(0, 1))
    return torch.stack(out)


# ----------------------------------------------------------------------------- "next" rows f2 / f3
@torch.no_grad()
def run_image_with_context_augmented(ldm, image, context, indices, controllers, thetas, noises=None, layers=(0, 1, 2, 3),
                                     upscale_size: int = 512):
    """eval.py:197-355 -- test-time augmentation ensemble: per iteration warp the image by theta, one captured forward
    with the selected tokens at `upscale_size`, un-warp the maps and a ones mask, accumulate; sum/num with NaN -> 0.
    ``image`` is [3,H,W]; ``thetas`` [iters,2,3] replace the reference's torch.rand draws (:238-243)."""
    k = len(indices)
    num = torch.zeros(k, upscale_size, upscale_size)
    tot = torch.zeros(k, upscale_size, upscale_size)
    for i in range(thetas.shape[0]):
        th = thetas[i:i + 1]
        aug = affine_warp(image[None], th)
        maps = run_and_find_attn(ldm, aug, context, controllers, layers=layers, upsample_res=upscale_size,
                                 indices=torch.as_tensor(indices), noise=None if noises is None else noises[i])
        maps = torch.stack(maps)
        num += affine_unwarp(torch.ones_like(maps), th).sum(dim=0)
        tot += affine_unwarp(maps, th).sum(dim=0)
    out = tot / num
    out[out != out] = 0
    return out


def vote_top_k(indices_list: torch.Tensor, top_k: int) -> torch.Tensor:
    """keypoint_regressor.py:101-106 -- most frequently selected token ids."""
    idx, counts = torch.unique(indices_list, return_counts=True)
    return idx[counts.argsort(descending=True)][:top_k]
