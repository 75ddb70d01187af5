"""Drop-in boundary: after ``stablekeypoints_b200.compat.install()`` the reference's own CLI,
``unsupervised_keypoints/main.py``, must import and drive Stage 1 -> Stage 2 through the B200 surface unchanged.

CPU tests (no GPU here): the REAL main.py (read-only from /root/reference, skipped where that tree is absent) is executed in
a subprocess under compat -- its import list (main.py:7-19), its argparse, and its Stage-1/Stage-2 call protocol
(main.py:197-228) with the three heavy entry points replaced by recorders that BIND the call to the B200 signatures.
The -m gpu test replays the same main.py lines on the 1/10-width synthetic model through the ``unsupervised_keypoints.*``
names (the GPU box has no reference checkout), with the argparse Namespace minted from the real parser
(tests/golden/main_args_defaults.json <- tests/golden/make_main_args.py).
"""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_ROOT = os.environ.get("SKP_REFERENCE_ROOT", "/root/reference")
HAVE_REF = os.path.isfile(os.path.join(REFERENCE_ROOT, "unsupervised_keypoints", "main.py"))
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="reference checkout not present (GPU box)")

_PRELUDE = r"""
import os, sys
sys.path.insert(0, %(root)r)
from tests.golden.make_main_args import stub_optional_imports
stub_optional_imports()
import stablekeypoints_b200.compat as compat
used = compat.install(reference_root=%(ref)r)
""" % {"root": ROOT, "ref": REFERENCE_ROOT}


def _run(code, timeout=300, cwd=None):
    out = subprocess.run([sys.executable, "-c", _PRELUDE + code], capture_output=True, text=True, timeout=timeout, cwd=cwd)
    return out


@needs_ref
def test_reference_main_import_list_resolves_under_compat():
    """main.py:7-19 verbatim; every hot-path name is the B200 one, every other name the reference's own, and the
    reference functions that call into the hot path (evaluate, precompute_all_keypoints) see the B200 functions."""
    code = r"""
import inspect
from unsupervised_keypoints.optimize_token import load_ldm
from unsupervised_keypoints.optimize import optimize_embedding
from unsupervised_keypoints.keypoint_regressor import (find_best_indices, precompute_all_keypoints, return_regressor,
                                                       return_regressor_visible, return_regressor_human36m)
from unsupervised_keypoints.eval import evaluate
from unsupervised_keypoints.visualize import visualize_attn_maps, create_vid
import stablekeypoints_b200.ptp_utils as p, stablekeypoints_b200.eval as e, stablekeypoints_b200.optimize as o
import unsupervised_keypoints.ptp_utils as rp, unsupervised_keypoints.eval as re_, unsupervised_keypoints.optimize as ro
assert rp is p and re_ is e and ro is o
assert load_ldm.__module__ == "stablekeypoints_b200.optimize_token"
assert optimize_embedding.__module__ == "stablekeypoints_b200.optimize"
assert find_best_indices.__module__ == "stablekeypoints_b200.keypoint_regressor"
for fn in (precompute_all_keypoints, return_regressor, return_regressor_visible, return_regressor_human36m):
    assert fn.__module__ == "unsupervised_keypoints.keypoint_regressor", fn
    assert inspect.getsourcefile(inspect.unwrap(fn)).startswith(used), fn
assert inspect.getsourcefile(inspect.unwrap(evaluate)).startswith(used)
assert inspect.getsourcefile(inspect.unwrap(visualize_attn_maps)).startswith(used)
g = inspect.unwrap(evaluate).__globals__
assert g["run_image_with_context_augmented"] is e.run_image_with_context_augmented
assert g["find_max_pixel"] is e.find_max_pixel and g["pixel_from_weighted_avg"] is e.pixel_from_weighted_avg
assert g["ptp_utils"] is p
g = inspect.unwrap(precompute_all_keypoints).__globals__
assert g["run_image_with_context_augmented"] is e.run_image_with_context_augmented and g["ptp_utils"] is p
# names only the reference defines stay reachable through the aliased modules
assert callable(e.find_corresponding_points) and callable(e.swap_points) and callable(o.variance_loss)
print("IMPORTS_OK")
"""
    out = _run(code)
    assert out.returncode == 0 and "IMPORTS_OK" in out.stdout, out.stderr[-3000:]


@needs_ref
def test_reference_main_cli_parses_under_compat():
    code = r"""
import runpy
sys.argv = ["main", "--help"]
try:
    runpy.run_module("unsupervised_keypoints.main", run_name="__main__")
except SystemExit as ex:
    print("EXIT", ex.code)
"""
    out = _run(code)
    assert out.returncode == 0 and "EXIT 0" in out.stdout and "--feature_upsample_res" in out.stdout, out.stderr[-3000:]


@needs_ref
def test_reference_main_stage1_stage2_protocol(tmp_path):
    """The REAL main.py, `--dataset_name custom`: load_ldm -> optimize_embedding -> torch.save -> find_best_indices -> torch.save.
    The three entry points are recorders bound to the B200 signatures (no GPU here): positional/keyword protocol, the
    Namespace attributes the B200 loop reads, and the artefacts."""
    code = r"""
import inspect, runpy, torch
import stablekeypoints_b200.optimize_token as ot, stablekeypoints_b200.optimize as op, stablekeypoints_b200.keypoint_regressor as kr
import unsupervised_keypoints.keypoint_regressor as rkr
calls = []
def recorder(name, real, result):
    sig = inspect.signature(real)
    def fake(*a, **k):
        calls.append((name, sig.bind(*a, **k).arguments))
        return result(*a, **k)
    return fake
READ = ("dataset_name", "dataset_loc", "max_len", "augment_degrees", "augment_scale", "augment_translate", "num_tokens", "batch_size",
        "lr", "num_steps", "layers", "noise_level", "device", "top_k", "furthest_point_num_samples", "sigma", "num_subjects",
        "top_k_strategy", "equivariance_attn_loss_weight", "sharpening_loss_weight", "wandb", "num_indices", "feature_upsample_res")
def fake_opt(ldm, args, controllers, num_gpus, **k):
    missing = [a for a in READ if not hasattr(args, a)]
    assert not missing, missing
    return torch.zeros(1, args.num_tokens, 768)
ot.load_ldm = recorder("load_ldm", ot.load_ldm, lambda *a, **k: ("LDM", {"dev": "CTL"}, 1))
op.optimize_embedding = recorder("optimize_embedding", op.optimize_embedding, fake_opt)
rkr.find_best_indices = recorder("find_best_indices", kr.find_best_indices, lambda *a, **k: torch.arange(10))
sys.argv = ["main", "--my_token", "T", "--dataset_name", "custom", "--dataset_loc", %(tmp)r, "--save_folder", %(out)r,
            "--model_type", "synthetic-small:0", "--num_steps", "2", "--num_tokens", "16"]
runpy.run_module("unsupervised_keypoints.main", run_name="__main__")
names = [c[0] for c in calls]
assert names == ["load_ldm", "optimize_embedding", "find_best_indices"], names
a = calls[0][1]
assert a["device"] == "cuda:0" and a["type"] == "synthetic-small:0" and a["feature_upsample_res"] == 128 and a["my_token"] == "T"
a = calls[1][1]
assert a["ldm"] == "LDM" and a["controllers"] == {"dev": "CTL"} and a["num_gpus"] == 1 and a["args"].num_tokens == 16
a = calls[2][1]
assert a["ldm"] == "LDM" and tuple(a["context"].shape) == (1, 16, 768) and a["num_gpus"] == 1
emb = torch.load(os.path.join(%(out)r, "embedding.pt")); idx = torch.load(os.path.join(%(out)r, "indices.pt"))
assert tuple(emb.shape) == (1, 16, 768) and idx.tolist() == list(range(10))
print("PROTOCOL_OK")
""" % {"tmp": str(tmp_path), "out": str(tmp_path / "outputs")}
    out = _run(code, cwd=str(tmp_path))
    assert out.returncode == 0 and "PROTOCOL_OK" in out.stdout, (out.stdout[-2000:], out.stderr[-3000:])


def test_compat_without_reference_checkout():
    """No checkout on sys.path (the GPU box): the hot-path modules and Stage 2 are importable under the reference's names."""
    code = ("import sys, os; sys.path.insert(0, %r); os.environ.pop('SKP_REFERENCE_ROOT', None);"
            "import stablekeypoints_b200.compat as c; r = c.install(reference_root=None) if not c.find_reference_root() else None;"
            "from unsupervised_keypoints import ptp_utils, optimize, optimize_token, eval, invertable_transform;"
            "from unsupervised_keypoints.keypoint_regressor import find_best_indices;"
            "from unsupervised_keypoints.optimize_token import load_ldm; from unsupervised_keypoints.optimize import optimize_embedding;"
            "import stablekeypoints_b200.ptp_utils as p; assert ptp_utils is p;"
            "assert find_best_indices.__module__ == 'stablekeypoints_b200.keypoint_regressor'; print('ok')" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


@needs_ref
def test_main_args_fixture_matches_reference_parser():
    """The committed Namespace fixture the GPU test uses == what the reference's parser produces today."""
    code = r"""
import json
from tests.golden.make_main_args import capture_args
print("ARGS" + json.dumps(capture_args(["--my_token", "TOKEN"]), sort_keys=True))
"""
    out = _run(code)
    assert out.returncode == 0, out.stderr[-3000:]
    live = json.loads(out.stdout.split("ARGS", 1)[1])
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "main_args_defaults.json")))
    assert live == want


# ----------------------------------------------------------------------------- GPU: main.py:197-228 on the B200 surface
@pytest.mark.gpu
def test_main_stage1_stage2_flow_on_gpu(tmp_path):
    """main.py:197-228 line by line through the compat names, 1/10-width synthetic SD model, the reference CLI's default
    Namespace (batch_size=4 -> B//G = 4 accumulation through the iteration/update CUDA graphs) with the sizes cut down."""
    import argparse
    code_args = json.load(open(os.path.join(ROOT, "tests", "golden", "main_args_defaults.json")))
    code_args.update(model_type="synthetic-small:0", dataset_name="synthetic", synthetic_size=128, max_len=6, num_steps=3,
                     num_tokens=24, feature_upsample_res=16, top_k=4, furthest_point_num_samples=8, num_indices=4,
                     save_folder=str(tmp_path / "outputs"), device="cuda:0")
    args = argparse.Namespace(**code_args)
    import stablekeypoints_b200.compat as compat
    compat.install(reference_root=None) if not HAVE_REF else compat.install()
    from unsupervised_keypoints.optimize_token import load_ldm
    from unsupervised_keypoints.optimize import optimize_embedding
    from unsupervised_keypoints.keypoint_regressor import find_best_indices
    torch.manual_seed(0)
    ldm, controllers, num_gpus = load_ldm(args.device, args.model_type, feature_upsample_res=args.feature_upsample_res,
                                          my_token=args.my_token)
    if not os.path.exists(args.save_folder):
        os.makedirs(args.save_folder)
    args.trace = []
    embedding = optimize_embedding(ldm, args, controllers, num_gpus)
    torch.save(embedding, os.path.join(args.save_folder, "embedding.pt"))
    indices = find_best_indices(ldm, embedding, args, controllers, num_gpus)
    torch.save(indices, os.path.join(args.save_folder, "indices.pt"))
    emb = torch.load(os.path.join(args.save_folder, "embedding.pt"))
    idx = torch.load(os.path.join(args.save_folder, "indices.pt"))
    assert tuple(emb.shape) == (1, 24, 768) and emb.dtype == torch.float32 and not emb.requires_grad
    assert bool(torch.isfinite(emb).all())
    assert len(args.trace) == args.num_steps * (args.batch_size // num_gpus)          # optimize.py:339
    assert all(bool(torch.isfinite(t["loss"])) for t in args.trace)
    assert idx.shape == (4,) and len(set(idx.tolist())) == 4 and all(0 <= i < 24 for i in idx.tolist())
