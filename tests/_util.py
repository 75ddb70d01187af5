"""Shared helpers for the test-suite (golden loading, tiny pipeline rebuild)."""
import hashlib
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TINY = dict(seed=0, attn_gain=6.0, image_size=128, n_tokens=12, res=16, top_k=4, num_candidates=8, sigma=1.5)


def load_golden(name):
    return {k: v for k, v in np.load(os.path.join(GOLDEN, name)).items()}


def state_checksum(module) -> str:
    h = hashlib.sha256()
    for k, v in module.state_dict().items():
        h.update(k.encode())
        h.update(v.detach().cpu().numpy().tobytes())
    return h.hexdigest()


def tiny_pipeline():
    from oracle import sd15
    pipe = sd15.make_pipeline(sd15.UNetConfig.tiny(), sd15.VAEConfig.tiny(), seed=TINY["seed"],
                              attn_gain=TINY["attn_gain"])
    return pipe


def check_tiny_weights(pipe, g):
    want = bytes(g["weights_sha256"]).decode()
    got = state_checksum(pipe.unet)
    assert got == want, ("seeded tiny-UNet weights differ from the ones the golden fixture was minted with "
                         "(torch RNG/init drift): regenerate with tests/golden/make_golden.py")


def rel_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
