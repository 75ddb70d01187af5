"""Install this package under the reference's module names so ``unsupervised_keypoints.main`` (and the notebook)
run unchanged on top of it:

    import stablekeypoints_b200.compat as c; c.install()
    from unsupervised_keypoints import ptp_utils, optimize, optimize_token, eval   # -> the B200 modules

Only the hot-path modules are aliased; everything else of the reference (datasets, visualize, keypoint_regressor,
main) keeps coming from the user's checkout of the reference, which must be importable for those names.
"""
from __future__ import annotations

import importlib
import sys
import types

_ALIASES = ("ptp_utils", "optimize", "optimize_token", "eval", "invertable_transform")
# keypoint_regressor is NOT aliased wholesale (Stage 3/4 stay with the reference); use
# stablekeypoints_b200.keypoint_regressor.find_best_indices explicitly or patch that one attribute.


def install(package_name: str = "unsupervised_keypoints") -> None:
    pkg = sys.modules.get(package_name)
    if pkg is None:
        pkg = types.ModuleType(package_name)
        pkg.__path__ = []
        sys.modules[package_name] = pkg
    for name in _ALIASES:
        mod = importlib.import_module(f"stablekeypoints_b200.{name}")
        sys.modules[f"{package_name}.{name}"] = mod
        setattr(pkg, name, mod)
