"""Mint the committed golden fixtures by running the REFERENCE'S OWN CODE (imported read-only
from /root/reference through oracle/ref_shim.py) on seeded synthetic inputs.

Run in the build container only:   python tests/golden/make_golden.py

The reference has no tests/golden vectors of its own, so these fixtures are the
pin: ``tests/test_oracle_vs_golden.py`` checks oracle/hotpath.py against them everywhere, and the
``-m gpu`` tests check the CUDA path against both.

Fixtures (all small, fp32/int64 .npz):
  tiny_stage1.npz   one Stage-1 iteration (optimize.py:349-422 call sequence, G=1) through the
                    reference hook (ptp_utils.py:472-573) on the tiny same-topology UNet of
                    oracle/sd15.py: the 4 stored maps, collected maps, token indices, both losses,
                    d(loss)/d(context), plus the eval-shaped collect_maps and arg-max / soft-arg-max.
  tiny_eval.npz     the eval-time augmentation ensemble (eval.py:197-355) run by the reference on the tiny pipeline,
                    plus the Stage-2 vote (keypoint_regressor.py:101-106).
  post_unet.npz     model-free: reference collect_maps / find_top_k_gaussian / furthest_point_sampling /
                    sharpening_loss / equivariance_loss / find_max_pixel / pixel_from_weighted_avg /
                    RandomAffineWithInverse on seeded random stores (includes ties and border cases).
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim, sd15  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

TINY = dict(seed=0, attn_gain=6.0, image_size=128, n_tokens=12, res=16, top_k=4, num_candidates=8, sigma=1.5)


def state_checksum(module) -> str:
    h = hashlib.sha256()
    for k, v in module.state_dict().items():
        h.update(k.encode())
        h.update(v.detach().cpu().numpy().tobytes())
    return h.hexdigest()


def tiny_pipeline():
    return sd15.make_pipeline(sd15.UNetConfig.tiny(), sd15.VAEConfig.tiny(), seed=TINY["seed"],
                              attn_gain=TINY["attn_gain"])


def tiny_inputs():
    from oracle import hotpath
    g = torch.Generator().manual_seed(7)
    image = hotpath.synthetic_image(seed=3, size=TINY["image_size"], blobs=6)
    context = torch.randn(1, TINY["n_tokens"], sd15.UNetConfig.tiny().cross_attention_dim, generator=g)
    lat = TINY["image_size"] // 8
    noise_a = torch.randn(1, 4, lat, lat, generator=g)
    noise_b = torch.randn(1, 4, lat, lat, generator=g)
    theta = hotpath.affine_theta(9.0, 0.9, 0.1, -0.07)
    return image, context, noise_a, noise_b, theta


class _FixedNoise:
    """Make the reference's torch.randn_like (ptp_utils.py:219) return an injected tensor."""

    def __init__(self, ref, noises):
        self.ref, self.noises, self.i = ref, noises, 0

    def __enter__(self):
        self.orig = torch.randn_like

        def fake(t, *a, **k):
            n = self.noises[self.i]
            self.i += 1
            assert n.shape == t.shape
            return n.clone()

        torch.randn_like = fake
        return self

    def __exit__(self, *exc):
        torch.randn_like = self.orig


def make_tiny_stage1(ref):
    pipe = tiny_pipeline()
    image, context, noise_a, noise_b, theta = tiny_inputs()
    R = TINY["res"]
    # optimize_token.py:51-69 (CPU branch) with the reference's own store + registration
    controllers = {torch.device("cpu"): ref.ptp_utils.AttentionStore()}
    pipe.unet.register_forward_pre_hook(
        lambda m, i: ref.ptp_utils.register_attention_control(m, controllers[i[0].device], feature_upsample_res=R))
    out = {"weights_sha256": np.frombuffer(state_checksum(pipe.unet).encode(), dtype=np.uint8),
           "vae_sha256": np.frombuffer(state_checksum(pipe.vae).encode(), dtype=np.uint8),
           "image": image.numpy(), "context": context.numpy(), "noise_a": noise_a.numpy(),
           "noise_b": noise_b.numpy(), "theta": theta.numpy()}

    # ---- store dump (BASELINE cfg1): find_pred_noise then read step_store before collect_maps
    ctx = context.clone().requires_grad_(True)
    with _FixedNoise(ref, [noise_a]):
        _, pred = ref.ptp_utils.find_pred_noise(pipe, image, ctx, noise_level=-1, device="cpu")
    store = controllers[torch.device("cpu")].step_store["attn"]
    assert len(store) == 4
    for i, s in enumerate(store):
        out[f"stored_{i}"] = s.detach().numpy()
    out["pred_noise"] = pred.detach().numpy()
    out["latent"] = ref.ptp_utils.image2latent(pipe, image.permute(0, 2, 3, 1).numpy(), "cpu").numpy()
    # eval-shaped aggregation on the same store (optimize.py:58-70): gather 3 tokens, bilinear to 64
    eval_idx = torch.tensor([5, 0, 9])
    import copy
    ctl2 = ref.ptp_utils.AttentionStore()
    ctl2.step_store = {"attn": [s.detach() for s in store]}
    ev = ref.optimize.collect_maps(ctl2, upsample_res=64, layers=[0, 1, 2, 3], indices=eval_idx)
    out["eval_indices"] = eval_idx.numpy()
    out["eval_maps"] = ev.numpy()
    out["eval_argmax"] = ref.eval.find_max_pixel(ev).numpy()
    out["eval_softargmax"] = ref.eval.pixel_from_weighted_avg(ev.clone()).numpy()
    ctl3 = ref.ptp_utils.AttentionStore()
    ctl3.step_store = {"attn": [s.detach() for s in store]}
    out["maps_layers_1_3"] = ref.optimize.collect_maps(ctl3, upsample_res=-1, layers=[1, 3]).numpy()
    controllers[torch.device("cpu")].reset()

    # ---- one Stage-1 iteration, optimize.py:349-422 call sequence for G=1
    ctx = context.clone().requires_grad_(True)
    transform = ref.invertable_transform.RandomAffineWithInverse(degrees=15, scale=(0.8, 1.0), translate=(0.25, 0.25))
    with _FixedNoise(ref, [noise_a, noise_b]):
        maps = ref.ptp_utils.run_and_find_attn(pipe, image, ctx, layers=[0, 1, 2, 3], noise_level=-1,
                                               upsample_res=-1, device="cpu", controllers=controllers)
        image_t = transform(image, theta=theta)
        maps_t = ref.ptp_utils.run_and_find_attn(pipe, image_t, ctx, layers=[0, 1, 2, 3], noise_level=-1,
                                                 upsample_res=-1, device="cpu", controllers=controllers)
    a, at = maps[0], maps_t[0]
    cand = ref.ptp_utils.find_top_k_gaussian(a, TINY["num_candidates"], sigma=TINY["sigma"], num_subjects=1)
    top = ref.ptp_utils.furthest_point_sampling(at, TINY["top_k"], cand)
    sharp = ref.optimize.sharpening_loss(a[top], device="cpu", sigma=TINY["sigma"], num_subjects=1)
    equiv = ref.optimize.equivariance_loss(a[top], at[top][None].repeat(1, 1, 1, 1), transform, 0)
    loss = equiv * 1000.0 + sharp * 100.0
    loss.backward()
    out.update(image_t=image_t.numpy(), maps=a.detach().numpy(), maps_t=at.detach().numpy(),
               candidates=cand.numpy(), indices=top.numpy(), sharp=sharp.detach().numpy(),
               equiv=equiv.detach().numpy(), loss=loss.detach().numpy(), dcontext=ctx.grad.numpy())
    # one Adam step exactly as optimize.py:320,424
    p = context.clone().requires_grad_(True)
    opt = torch.optim.Adam([p], lr=5e-3)
    p.grad = ctx.grad.clone()
    opt.step()
    out["context_after_adam"] = p.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "tiny_stage1.npz"), **out)
    print("tiny_stage1.npz:", {k: v.shape for k, v in out.items() if k.startswith(("stored", "maps", "dcontext"))},
          "loss", float(loss), "indices", top.tolist())


def make_post_unet(ref):
    g = torch.Generator().manual_seed(11)
    out = {}
    # ---- collect_maps on a random store, train + eval shapes, with layer subsets
    L, BH, R, N = 4, 8, 16, 20
    store = [torch.softmax(3.0 * torch.randn(BH, R * R, N, generator=g), dim=-1) for _ in range(L)]
    for i, s in enumerate(store):
        out[f"store_{i}"] = s.numpy()

    def collect(**kw):
        c = ref.ptp_utils.AttentionStore()
        c.step_store = {"attn": [s.clone() for s in store]}
        return ref.optimize.collect_maps(c, **kw)

    out["collect_train"] = collect(upsample_res=-1, layers=[0, 1, 2, 3]).numpy()
    out["collect_layers_02"] = collect(upsample_res=-1, layers=[0, 2]).numpy()
    idx = torch.tensor([7, 3, 19, 0, 11])
    out["collect_idx"] = idx.numpy()
    out["collect_eval_48"] = collect(upsample_res=48, layers=[0, 1, 2, 3], indices=idx).numpy()
    out["collect_same_res"] = collect(upsample_res=16, layers=[0, 1, 2, 3]).numpy()

    # ---- selection + losses on peaky maps with deliberate ties / border peaks
    T, H = 20, 32
    ys = (torch.arange(H).float() + 0.5).reshape(1, H, 1)
    xs = (torch.arange(H).float() + 0.5).reshape(1, 1, H)
    cy = torch.rand(T, 1, 1, generator=g) * H
    cx = torch.rand(T, 1, 1, generator=g) * H
    sg = 1.0 + 4.0 * torch.rand(T, 1, 1, generator=g)
    maps = torch.exp(-((ys - cy) ** 2 + (xs - cx) ** 2) / (2 * sg * sg)) * (0.2 + torch.rand(T, 1, 1, generator=g))
    maps = maps + 0.02 * torch.rand(T, H, H, generator=g)
    maps[3] = 0.0; maps[3, 5, 9] = 1.0; maps[3, 20, 2] = 1.0          # exact tie: first occurrence must win
    maps[4] = maps[2]                                                  # duplicate location: FPS distance 0 ties
    maps[6, 0, 0] = 5.0                                                # corner peak
    maps[7, H - 1, H - 1] = 5.0
    maps_t = torch.roll(maps, shifts=(3, -2), dims=(1, 2)) + 0.01 * torch.rand(T, H, H, generator=g)
    out["maps"], out["maps_t"] = maps.numpy(), maps_t.numpy()
    out["find_max_pixel"] = ref.eval.find_max_pixel(maps).numpy()
    out["find_k_max_pixels_3"] = ref.eval.find_k_max_pixels(maps, num=3).numpy()
    for sigma in (1.0, 2.0):
        cand = ref.ptp_utils.find_top_k_gaussian(maps, 9, sigma=sigma, num_subjects=1)
        out[f"topk_gaussian_s{sigma}"] = cand.numpy()
        out[f"fps_s{sigma}"] = ref.ptp_utils.furthest_point_sampling(maps_t, 5, cand).numpy()
    # the KL scores themselves (ptp_utils.py:97-108 evaluated piecewise)
    loc = ref.eval.find_k_max_pixels(maps, num=1) / H
    p = torch.softmax(maps.view(T, H * H) + 1e-5, dim=-1)
    tg = ref.optimize_token.gaussian_circles(loc, size=H, sigma=2.0, device="cpu").reshape(T, H * H) + 1e-5
    tg = tg / tg.sum(dim=-1, keepdim=True)
    out["kl_s2.0"] = torch.sum(tg * (torch.log(tg) - torch.log(p)), dim=-1).numpy()
    out["fps_all_candidates"] = ref.ptp_utils.furthest_point_sampling(maps, 6, torch.arange(T)).numpy()
    sel = torch.tensor([1, 6, 3, 12, 7])
    out["sel"] = sel.numpy()
    m = maps.clone().requires_grad_(True)
    mt = maps_t.clone().requires_grad_(True)
    sharp = ref.optimize.sharpening_loss(m[sel], device="cpu", sigma=2.0, num_subjects=1)
    tr = ref.invertable_transform.RandomAffineWithInverse(degrees=15, scale=(0.8, 1.0), translate=(0.25, 0.25))
    from oracle import hotpath
    theta = hotpath.affine_theta(-12.0, 0.85, -0.2, 0.15)
    tr.last_params = {"theta": theta}
    equiv = ref.optimize.equivariance_loss(m[sel], mt[sel][None], tr, 0)
    (100.0 * sharp + 1000.0 * equiv).backward()
    out.update(theta=theta.numpy(), sharp=sharp.detach().numpy(), equiv=equiv.detach().numpy(),
               dmaps=m.grad.numpy(), dmaps_t=mt.grad.numpy())
    out["unwarp"] = tr.inverse(maps_t[sel][None]).numpy()
    img = torch.rand(2, 3, 40, 40, generator=g)
    th2 = torch.cat([theta, hotpath.affine_theta(15.0, 1.0, 0.25, -0.25)], 0)
    out["img"], out["theta2"] = img.numpy(), th2.numpy()
    out["warp"] = tr(img, theta=th2).numpy()
    # RNG draw order of the augmentation (invertable_transform.py:42-57)
    torch.manual_seed(123)
    tr(torch.zeros(3, 1, 4, 4))
    out["theta_seed123"] = tr.last_params["theta"].numpy()
    # soft-arg-max (eval.py:113-155) incl. its in-place zeroing
    hm = torch.rand(4, 48, 48, generator=g) ** 4
    hm[1, 0, 47] = 3.0
    out["soft_in"] = hm.numpy()
    hm2 = hm.clone()
    out["soft_out"] = ref.eval.pixel_from_weighted_avg(hm2).numpy()
    out["soft_in_after"] = hm2.numpy()
    np.savez_compressed(os.path.join(HERE, "post_unet.npz"), **out)
    print("post_unet.npz: sharp", float(sharp), "equiv", float(equiv), "fps", out["fps_s2.0"].tolist())


def make_tiny_eval(ref):
    """eval.py:197-355 run by the reference itself on the tiny pipeline (3 augmentation iterations, 3 tokens, 64^2)."""
    from oracle import hotpath
    pipe = tiny_pipeline()
    image, context, noise_a, noise_b, _ = tiny_inputs()
    g = torch.Generator().manual_seed(21)
    noises = [noise_a, noise_b, torch.randn(noise_a.shape, generator=g)]
    controllers = {torch.device("cpu"): ref.ptp_utils.AttentionStore()}
    pipe.unet.register_forward_pre_hook(
        lambda m, i: ref.ptp_utils.register_attention_control(m, controllers[i[0].device], feature_upsample_res=TINY["res"]))
    indices = torch.tensor([5, 0, 9])
    torch.manual_seed(77)
    thetas = torch.cat([hotpath.sample_affine_params(1, degrees=30, scale=(0.9, 1.1), translate=(0.1, 0.1)) for _ in range(3)], 0)
    torch.manual_seed(77)
    with _FixedNoise(ref, noises):
        out = ref.eval.run_image_with_context_augmented(pipe, image[0], context, indices, device="cpu", layers=[0, 1, 2, 3],
                                                        augmentation_iterations=3, controllers=controllers, num_gpus=1,
                                                        upscale_size=64)
    res = {"indices": indices.numpy(), "thetas": thetas.numpy(), "noise_c": noises[2].numpy(), "ensemble": out.numpy(),
           "keypoints": (ref.eval.find_max_pixel(out) / 64.0).numpy()}
    votes = torch.tensor([3, 7, 7, 1, 3, 7, 9, 1, 1, 0, 3, 3])
    idx, counts = torch.unique(votes, return_counts=True)
    res["votes"], res["voted_top3"] = votes.numpy(), idx[counts.argsort(descending=True)][:3].numpy()
    np.savez_compressed(os.path.join(HERE, "tiny_eval.npz"), **res)
    print("tiny_eval.npz: ensemble", out.shape, "keypoints", res["keypoints"].tolist())


if __name__ == "__main__":
    assert ref_shim.available(), "reference tree not present: fixtures can only be minted in the build container"
    torch.set_num_threads(8)
    ref = ref_shim.load()
    make_post_unet(ref)
    make_tiny_stage1(ref)
    make_tiny_eval(ref)
    for f in ("tiny_stage1.npz", "post_unet.npz", "tiny_eval.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
