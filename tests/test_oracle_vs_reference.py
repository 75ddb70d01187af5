"""Live pin of the oracle against the reference's own code on fresh seeds.
Only runs where /root/reference exists (the build container); skipped on the GPU box."""
import numpy as np
import pytest
import torch

from oracle import hotpath as hp
from oracle import ref_shim
from tests._util import rel_err

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    return ref_shim.load()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_post_unet_functions(ref, seed):
    g = torch.Generator().manual_seed(100 + seed)
    T, H = 16, 24
    maps = torch.rand(T, H, H, generator=g) ** 6
    maps_t = torch.rand(T, H, H, generator=g) ** 6
    assert np.array_equal(hp.find_max_pixel(maps).numpy(), ref.eval.find_max_pixel(maps).numpy())
    cand_ref = ref.ptp_utils.find_top_k_gaussian(maps, 7, sigma=2.0, num_subjects=1)
    cand = hp.find_top_k_gaussian(maps, 7, sigma=2.0)
    assert np.array_equal(cand.numpy(), cand_ref.numpy())
    assert np.array_equal(hp.furthest_point_sampling(maps_t, 4, cand).numpy(),
                          ref.ptp_utils.furthest_point_sampling(maps_t, 4, cand_ref).numpy())
    sel = cand[:4]
    assert rel_err(hp.sharpening_loss(maps[sel], 2.0), ref.optimize.sharpening_loss(maps[sel], sigma=2.0, device="cpu")) < 2e-6
    tr = ref.invertable_transform.RandomAffineWithInverse(degrees=15, scale=(0.8, 1.0), translate=(0.25, 0.25))
    torch.manual_seed(seed)
    img = torch.rand(1, 3, 32, 32, generator=g)
    warped_ref = tr(img)
    theta = tr.last_params["theta"]
    torch.manual_seed(seed)
    assert np.allclose(hp.sample_affine_params(1).numpy(), theta.numpy(), atol=1e-7)
    assert rel_err(hp.affine_warp(img, theta), warped_ref) < 2e-6
    assert rel_err(hp.equivariance_loss(maps[sel], maps_t[sel][None], theta, 0),
                   ref.optimize.equivariance_loss(maps[sel], maps_t[sel][None], tr, 0)) < 2e-6


def test_hook_on_standin_tree(ref):
    """Reference hook vs restated hook on one CrossAttention module (ptp_utils.py:480-541)."""
    from oracle import sd15
    torch.manual_seed(5)

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.up_blocks = torch.nn.ModuleList([sd15.CrossAttention(32, 24, 4, 8) for _ in range(5)])
            self.down_blocks = torch.nn.ModuleList([sd15.CrossAttention(32, 24, 4, 8)])

    x = torch.randn(1, 16, 32)
    ctx = torch.randn(1, 6, 24)
    outs = []
    for reg, store in ((ref.ptp_utils.register_attention_control, ref.ptp_utils.AttentionStore()),
                       (hp.register_capture, hp.AttentionStore())):
        torch.manual_seed(5)
        net = Net()
        reg(net, store, 8)
        ys = [m(x, context=ctx) for m in net.up_blocks] + [net.down_blocks[0](x, context=ctx)]
        outs.append((ys, store.step_store["attn"], store.num_att_layers))
    (ya, sa, na), (yb, sb, nb) = outs
    assert na == nb == 5 and len(sa) == len(sb) == 4
    for a, b in zip(ya, yb):
        assert rel_err(b, a) < 2e-6
    for a, b in zip(sa, sb):
        assert a.shape == b.shape == (4, 64, 6) and rel_err(b, a) < 2e-6


def _fps_edge_cases():
    """(maps, top_k, candidates): fewer candidates than top_k (the reference returns a SHORTER list, ptp_utils.py:139-159),
    candidates whose peaks coincide (zero distances, ties resolved by the strict '>' scans), top_k == 2 (the pair only)."""
    g = torch.Generator().manual_seed(7)
    T, H = 12, 16
    maps = torch.rand(T, H, H, generator=g) * 0.5
    peaks = [(3, 4), (3, 4), (10, 12), (0, 0), (15, 15), (3, 4), (8, 8), (8, 9), (0, 15), (15, 0), (7, 7), (10, 12)]
    for t, (y, x) in enumerate(peaks):
        maps[t, y, x] = 1.0 + 0.01 * t
    yield maps, 10, torch.tensor([5, 0, 2, 7])                   # 4 candidates for top_k = 10
    yield maps, 6, torch.tensor([0, 1, 5, 2, 11, 6, 7])          # three coincident peaks + one coincident pair
    yield maps, 2, torch.tensor([3, 4, 8, 9])                    # only the furthest pair (two diagonals tie)
    yield maps, 5, torch.tensor([0, 1, 5])                       # every distance is zero: the pair is the first (i < j)
    yield maps, 12, torch.arange(12)                             # candidates == top_k == all tokens


def test_fps_edge_cases_vs_reference(ref):
    for maps, top_k, cand in _fps_edge_cases():
        want = ref.ptp_utils.furthest_point_sampling(maps, top_k, cand)
        got = hp.furthest_point_sampling(maps, top_k, cand)
        assert np.array_equal(got.numpy(), want.numpy()), (top_k, cand.tolist(), got.tolist(), want.tolist())
