"""Pins oracle/hotpath.py (+ oracle/sd15.py driven by it) against fixtures minted by the
reference's own code (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import hotpath as hp
from tests._util import TINY, check_tiny_weights, load_golden, rel_err, tiny_pipeline

TOL = 2e-6  # same fp32 arithmetic, different op grouping


@pytest.fixture(scope="module")
def post():
    return load_golden("post_unet.npz")


@pytest.fixture(scope="module")
def tiny():
    return load_golden("tiny_stage1.npz")


def _store(post):
    c = hp.AttentionStore()
    c.step_store = {"attn": [torch.from_numpy(post[f"store_{i}"]) for i in range(4)]}
    return c


def test_collect_maps_train_eval_layers(post):
    assert rel_err(hp.collect_maps(_store(post), -1, (0, 1, 2, 3)), post["collect_train"]) < TOL
    assert rel_err(hp.collect_maps(_store(post), -1, (0, 2)), post["collect_layers_02"]) < TOL
    idx = torch.from_numpy(post["collect_idx"])
    assert rel_err(hp.collect_maps(_store(post), 48, (0, 1, 2, 3), idx), post["collect_eval_48"]) < TOL
    assert rel_err(hp.collect_maps(_store(post), 16, (0, 1, 2, 3)), post["collect_same_res"]) < TOL
    c = _store(post)
    hp.collect_maps(c, -1)
    assert c.step_store["attn"] == []  # optimize.py:77 side effect


def test_argmax_and_kmax(post):
    maps = torch.from_numpy(post["maps"])
    assert np.array_equal(hp.find_max_pixel(maps).numpy(), post["find_max_pixel"])
    assert np.array_equal(hp.find_k_max_pixels(maps, 3).numpy(), post["find_k_max_pixels_3"])
    # exact tie resolves to the first occurrence (row-major)
    assert hp.find_max_pixel(maps)[3].tolist() == [5.5, 9.5]


def test_selection(post):
    maps, maps_t = torch.from_numpy(post["maps"]), torch.from_numpy(post["maps_t"])
    assert rel_err(hp.gaussian_kl_scores(maps, 2.0), post["kl_s2.0"]) < 1e-5
    for s in (1.0, 2.0):
        cand = hp.find_top_k_gaussian(maps, 9, sigma=s)
        assert np.array_equal(cand.numpy(), post[f"topk_gaussian_s{s}"])
        assert np.array_equal(hp.furthest_point_sampling(maps_t, 5, cand).numpy(), post[f"fps_s{s}"])
    assert np.array_equal(hp.furthest_point_sampling(maps, 6, torch.arange(maps.shape[0])).numpy(),
                          post["fps_all_candidates"])


def test_losses_and_grads(post):
    maps = torch.from_numpy(post["maps"]).requires_grad_(True)
    maps_t = torch.from_numpy(post["maps_t"]).requires_grad_(True)
    sel = torch.from_numpy(post["sel"])
    theta = torch.from_numpy(post["theta"])
    sharp = hp.sharpening_loss(maps[sel], sigma=2.0)
    equiv = hp.equivariance_loss(maps[sel], maps_t[sel][None], theta, 0)
    (100.0 * sharp + 1000.0 * equiv).backward()
    assert rel_err(sharp.detach(), post["sharp"]) < TOL
    assert rel_err(equiv.detach(), post["equiv"]) < TOL
    assert rel_err(maps.grad, post["dmaps"]) < TOL
    assert rel_err(maps_t.grad, post["dmaps_t"]) < TOL
    assert rel_err(hp.affine_unwarp(maps_t.detach()[sel][None], theta), post["unwarp"]) < TOL


def test_affine_warp_and_rng_order(post):
    img, th2 = torch.from_numpy(post["img"]), torch.from_numpy(post["theta2"])
    assert rel_err(hp.affine_warp(img, th2), post["warp"]) < TOL
    torch.manual_seed(123)
    assert np.allclose(hp.sample_affine_params(3).numpy(), post["theta_seed123"], atol=1e-7)


def test_soft_argmax_inplace(post):
    hm = torch.from_numpy(post["soft_in"]).clone()
    out = hp.pixel_from_weighted_avg(hm)
    assert rel_err(out, post["soft_out"]) < TOL
    assert np.array_equal(hm.numpy(), post["soft_in_after"])  # eval.py:138 zeroes its input in place


def test_tiny_store_dump(tiny):
    """BASELINE cfg1: one captured forward + AttentionStore dump through the restated hook."""
    pipe = tiny_pipeline()
    check_tiny_weights(pipe, tiny)
    ldm, controllers, _ = hp.load_oracle_ldm(pipe, TINY["res"])
    ctx = torch.from_numpy(tiny["context"])
    image = torch.from_numpy(tiny["image"])
    assert rel_err(hp.encode_image(ldm, image), tiny["latent"]) < 1e-5
    _, pred = hp.find_pred_noise(ldm, image, ctx, noise=torch.from_numpy(tiny["noise_a"]))
    store = controllers[torch.device("cpu")].step_store["attn"]
    assert len(store) == 4
    for i, s in enumerate(store):
        assert s.shape == tiny[f"stored_{i}"].shape
        assert rel_err(s, tiny[f"stored_{i}"]) < 2e-5
        assert torch.allclose(s.sum(-1), torch.ones(s.shape[:2]), atol=1e-5)  # softmax over tokens
    assert rel_err(pred, tiny["pred_noise"]) < 2e-5
    c2 = hp.AttentionStore(); c2.step_store = {"attn": [s.detach() for s in store]}
    ev = hp.collect_maps(c2, 64, (0, 1, 2, 3), torch.from_numpy(tiny["eval_indices"]))
    assert rel_err(ev, tiny["eval_maps"]) < 2e-5
    assert np.array_equal(hp.find_max_pixel(ev).numpy(), tiny["eval_argmax"])
    assert rel_err(hp.pixel_from_weighted_avg(ev.clone()), tiny["eval_softargmax"]) < 1e-5
    c3 = hp.AttentionStore(); c3.step_store = {"attn": [s.detach() for s in store]}
    assert rel_err(hp.collect_maps(c3, -1, (1, 3)), tiny["maps_layers_1_3"]) < 2e-5


def test_tiny_stage1_iteration(tiny):
    pipe = tiny_pipeline()
    check_tiny_weights(pipe, tiny)
    ldm, controllers, _ = hp.load_oracle_ldm(pipe, TINY["res"])
    ctx = torch.from_numpy(tiny["context"]).clone().requires_grad_(True)
    r = hp.stage1_iteration(ldm, controllers, torch.from_numpy(tiny["image"]), ctx, torch.from_numpy(tiny["theta"]),
                            torch.from_numpy(tiny["noise_a"]), torch.from_numpy(tiny["noise_b"]),
                            top_k=TINY["top_k"], num_candidates=TINY["num_candidates"], sigma=TINY["sigma"])
    assert rel_err(r["maps"], tiny["maps"]) < 2e-5
    assert rel_err(r["maps_t"], tiny["maps_t"]) < 2e-5
    assert np.array_equal(r["indices"].numpy(), tiny["indices"])
    assert rel_err(r["sharp"], tiny["sharp"]) < 2e-5
    assert rel_err(r["equiv"], tiny["equiv"]) < 2e-5
    assert rel_err(r["loss"], tiny["loss"]) < 2e-5
    assert rel_err(ctx.grad, tiny["dcontext"]) < 1e-4
    p = torch.from_numpy(tiny["context"]).clone()
    hp.adam_step(p, torch.from_numpy(tiny["dcontext"]), torch.zeros_like(p), torch.zeros_like(p), 1)
    assert rel_err(p, tiny["context_after_adam"]) < 1e-6


def test_tiny_eval_ensemble_and_vote():
    """'next' rows f2/f3: oracle restatement vs the reference's own run_image_with_context_augmented / vote."""
    g = load_golden("tiny_eval.npz")
    t = load_golden("tiny_stage1.npz")
    pipe = tiny_pipeline()
    check_tiny_weights(pipe, t)
    ldm, controllers, _ = hp.load_oracle_ldm(pipe, TINY["res"])
    noises = [torch.from_numpy(t["noise_a"]), torch.from_numpy(t["noise_b"]), torch.from_numpy(g["noise_c"])]
    out = hp.run_image_with_context_augmented(ldm, torch.from_numpy(t["image"])[0], torch.from_numpy(t["context"]),
                                              torch.from_numpy(g["indices"]), controllers, torch.from_numpy(g["thetas"]),
                                              noises, upscale_size=64)
    assert rel_err(out, g["ensemble"]) < 2e-5
    assert np.array_equal((hp.find_max_pixel(out) / 64.0).numpy(), g["keypoints"])
    assert np.array_equal(hp.vote_top_k(torch.from_numpy(g["votes"]), 3).numpy(), g["voted_top3"])
