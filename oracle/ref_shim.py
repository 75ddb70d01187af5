"""Import the reference's own Python modules read-only from /root/reference -- TEST INFRASTRUCTURE ONLY.

Only usable in the build container (the GPU box has no /root/reference).  Used by
``tests/golden/make_golden.py`` to mint fixtures and by ``tests/test_oracle_vs_reference.py``
(skipped when the tree is absent) to pin ``oracle/hotpath.py`` against the real code.

Two import blockers are stubbed:
  * ``from diffusers import StableDiffusionPipeline, DDIMScheduler`` (optimize_token.py:16)
  * ``from datasets.celeba import CelebA`` etc. (optimize.py:10-17, eval.py:6-12), which collide
    with the HuggingFace ``datasets`` package installed in this image.
Nothing is copied: the modules are executed from where they lie.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SKP_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "unsupervised_keypoints", "ptp_utils.py"))


def _stub(name: str, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


_cached = None


def load():
    """Returns a namespace with the reference modules (ptp_utils, optimize, optimize_token, eval, invertable_transform)."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    saved = {k: v for k, v in sys.modules.items() if k == "datasets" or k.startswith("datasets.") or k == "diffusers"}
    for k in saved:
        del sys.modules[k]
    ds = _stub("datasets")
    ds.__path__ = []
    subs = {"celeba": ["CelebA"], "custom_images": ["CustomDataset"], "cub": [], "cub_parts": [], "taichi": [],
            "human36m": [], "unaligned_human36m": [], "deepfashion": []}
    for sub, names in subs.items():
        setattr(ds, sub, _stub("datasets." + sub, **{n: object for n in names}))
    _stub("diffusers", StableDiffusionPipeline=object, DDIMScheduler=object)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        from unsupervised_keypoints import ptp_utils, optimize, optimize_token, invertable_transform  # noqa
        from unsupervised_keypoints import eval as ref_eval  # noqa
    finally:
        sys.path.remove(REFERENCE_ROOT)
        # un-shadow HuggingFace datasets / diffusers for the rest of the process
        for k in list(sys.modules):
            if k == "datasets" or k.startswith("datasets.") or k == "diffusers":
                del sys.modules[k]
        sys.modules.update(saved)
    _cached = types.SimpleNamespace(ptp_utils=ptp_utils, optimize=optimize, optimize_token=optimize_token,
                                    eval=ref_eval, invertable_transform=invertable_transform)
    return _cached
