"""Install this package under the reference's module names so ``unsupervised_keypoints.main`` (and the notebook)
run unchanged on top of it:

    import stablekeypoints_b200.compat as c; c.install(reference_root="/path/to/StableKeypoints")
    import runpy; runpy.run_module("unsupervised_keypoints.main", run_name="__main__")      # the reference's own CLI

What ``install`` does (the reference imports, main.py:7-19):

  * ``unsupervised_keypoints`` becomes a package whose ``__path__`` is the reference checkout's directory, so every module
    this package does NOT replace (``main``, ``keypoint_regressor``, ``visualize``, ``sdxl_monkey_patch`` ...) still
    resolves to the reference's own file.  Without a checkout the package holds the hot-path modules only.
  * the hot-path modules (``ptp_utils``, ``optimize``, ``optimize_token``, ``eval``, ``invertable_transform``) ARE the B200
    modules (``sys.modules["unsupervised_keypoints.ptp_utils"] is stablekeypoints_b200.ptp_utils``).  Names the reference
    module has and ours does not (``eval.evaluate`` -- main.py:18 / eval.py:375, ``eval.find_corresponding_points``,
    the dead helpers ...) are taken from the reference's file, loaded privately, and that private module's globals are
    patched with the B200 functions -- so ``evaluate`` calls OUR ``run_image_with_context_augmented``.
  * ``keypoint_regressor`` stays the reference's module (Stage 3/4 are NumPy ``pinv`` fits and dataset walks) with
    ``find_best_indices`` (Stage 2) replaced by the B200 one.
  * two environment shims the reference's import lines need when run outside its conda env: a ``diffusers`` stub when
    that package is not installed (optimize_token.py:16 imports two names that load_ldm -- replaced here -- was their only
    user), and the reference's ``datasets/`` directory mounted as the top-level ``datasets`` package when the installed
    ``datasets`` is HuggingFace's (the name clash of SURVEY 2.1).
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types
import warnings
from typing import Optional

_ALIASES = ("ptp_utils", "optimize", "optimize_token", "eval", "invertable_transform")
_REF_DATASETS = ("celeba", "custom_images", "cub", "cub_parts", "taichi", "human36m", "unaligned_human36m", "deepfashion")


def find_reference_root(package_name: str = "unsupervised_keypoints") -> Optional[str]:
    """The reference checkout: $SKP_REFERENCE_ROOT, else the first sys.path entry that holds <package>/main.py."""
    cands = [os.environ.get("SKP_REFERENCE_ROOT")] + list(sys.path)
    for c in cands:
        if c is None:
            continue
        c = c or os.getcwd()
        if os.path.isfile(os.path.join(c, package_name, "main.py")) and os.path.isfile(os.path.join(c, package_name, "ptp_utils.py")):
            return os.path.abspath(c)
    return None


def _mount_reference_datasets(root: str) -> None:
    ds_dir = os.path.join(root, "datasets")
    if not os.path.isdir(ds_dir):
        return
    cur = sys.modules.get("datasets")
    if cur is not None and ds_dir in list(getattr(cur, "__path__", [])):
        return
    try:
        if cur is None:
            spec = importlib.util.find_spec("datasets")
            if spec is not None and spec.submodule_search_locations is not None and ds_dir in list(spec.submodule_search_locations):
                return      # already resolves to the reference's directory
    except (ImportError, ValueError):
        pass
    for k in [k for k in sys.modules if k == "datasets" or k.startswith("datasets.")]:
        sys.modules["_skp_shadowed_" + k] = sys.modules.pop(k)      # HuggingFace datasets (if loaded) stays reachable
    pkg = types.ModuleType("datasets")
    pkg.__path__ = [ds_dir]
    pkg.__skp_reference_datasets__ = True
    sys.modules["datasets"] = pkg


def _stub_diffusers_if_missing() -> None:
    if "diffusers" in sys.modules:
        return
    try:
        if importlib.util.find_spec("diffusers") is not None:
            return
    except (ImportError, ValueError):
        pass
    stub = types.ModuleType("diffusers")
    stub.__skp_stub__ = True

    class _Absent:
        def __init__(self, *a, **k):
            raise RuntimeError("diffusers is not installed; stablekeypoints_b200.optimize_token.load_ldm replaces its only user")

        from_pretrained = classmethod(lambda cls, *a, **k: cls())

    stub.StableDiffusionPipeline = _Absent
    stub.DDIMScheduler = _Absent
    sys.modules["diffusers"] = stub


def _load_private(package_name: str, name: str, root: str):
    """Execute the reference's <name>.py under a private module name; its absolute imports resolve to the aliases."""
    path = os.path.join(root, package_name, name + ".py")
    if not os.path.isfile(path):
        return None
    priv = f"{package_name}._reference_{name}"
    if priv in sys.modules:
        return sys.modules[priv]
    spec = importlib.util.spec_from_file_location(priv, path)
    mod = importlib.util.module_from_spec(spec)
    mod.__package__ = package_name
    sys.modules[priv] = mod
    try:
        spec.loader.exec_module(mod)
    except Exception as e:      # an optional dependency of the reference's own file is missing: keep the B200 names only
        del sys.modules[priv]
        warnings.warn(f"stablekeypoints_b200.compat: the reference's {name}.py could not be loaded for its extra names ({e!r})")
        return None
    return mod


def _public(mod):
    return {k: v for k, v in vars(mod).items() if not k.startswith("_")}


def _overlay(ours, ref) -> None:
    """ours gains the names only the reference defines; the reference's functions see OUR names through their globals."""
    mine = _public(ours)
    theirs = _public(ref)
    own = {k: v for k, v in mine.items()
           if getattr(v, "__module__", None) == ours.__name__ and (callable(v) or isinstance(v, type))}
    for k, v in theirs.items():
        if k not in mine:
            setattr(ours, k, v)
    for k, v in own.items():
        setattr(ref, k, v)


def install(package_name: str = "unsupervised_keypoints", reference_root: Optional[str] = None,
            mount_datasets: bool = True) -> Optional[str]:
    """Returns the reference root that was used (None: no checkout found, only the hot-path modules are importable)."""
    root = os.path.abspath(reference_root) if reference_root else find_reference_root(package_name)
    pkg = sys.modules.get(package_name)
    if pkg is None or not hasattr(pkg, "__path__"):
        pkg = types.ModuleType(package_name)
        pkg.__path__ = []
        sys.modules[package_name] = pkg
    if root is not None:
        ref_dir = os.path.join(root, package_name)
        paths = [p for p in list(pkg.__path__) if p != ref_dir]
        pkg.__path__ = [ref_dir] + paths
        if mount_datasets:
            _mount_reference_datasets(root)
        _stub_diffusers_if_missing()
    ours = {}
    for name in _ALIASES:
        mod = importlib.import_module(f"stablekeypoints_b200.{name}")
        sys.modules[f"{package_name}.{name}"] = mod
        setattr(pkg, name, mod)
        ours[name] = mod
    if root is None:
        kr = importlib.import_module("stablekeypoints_b200.keypoint_regressor")
        sys.modules[f"{package_name}.keypoint_regressor"] = kr
        setattr(pkg, "keypoint_regressor", kr)
        return None
    # leaf modules first: the reference files import each other by absolute name and get the aliases
    for name in ("invertable_transform", "eval", "optimize_token", "ptp_utils", "optimize"):
        ref = _load_private(package_name, name, root)
        if ref is not None:
            _overlay(ours[name], ref)
    # Stage 2 (find_best_indices) on the B200 forward; Stages 3/4 stay the reference's
    try:
        kr = importlib.import_module(f"{package_name}.keypoint_regressor")
        mine = importlib.import_module("stablekeypoints_b200.keypoint_regressor")
        kr.find_best_indices = mine.find_best_indices
        kr.vote_top_k = mine.vote_top_k
    except Exception as e:
        warnings.warn(f"stablekeypoints_b200.compat: the reference's keypoint_regressor.py could not be imported ({e!r}); "
                      "using the B200 Stage-2 module only")
        kr = importlib.import_module("stablekeypoints_b200.keypoint_regressor")
        sys.modules[f"{package_name}.keypoint_regressor"] = kr
        setattr(pkg, "keypoint_regressor", kr)
    return root
