"""Mint tests/golden/main_args_defaults.json: the argparse Namespace the reference's own CLI (unsupervised_keypoints/main.py:23-195)
hands to load_ldm / optimize_embedding / find_best_indices, captured by running the REAL main.py (read-only from /root/reference)
up to its parse_args() call.  TEST INFRASTRUCTURE: run in the build container (`python tests/golden/make_main_args.py`); the GPU
box has no /root/reference and uses the committed JSON."""
import argparse
import json
import os
import runpy
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REFERENCE_ROOT = os.environ.get("SKP_REFERENCE_ROOT", "/root/reference")


class _Captured(Exception):
    pass


def stub_optional_imports():
    """Packages only the reference's plotting / dataset readers import and this image lacks."""
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "h5py", "imageio"):
        if m not in sys.modules:
            try:
                __import__(m)
            except ImportError:
                sys.modules[m] = types.ModuleType(m)
    if isinstance(sys.modules.get("matplotlib"), types.ModuleType) and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
        sys.modules["matplotlib"].colors = sys.modules["matplotlib.colors"]


def capture_args(argv):
    import stablekeypoints_b200.compat as compat
    stub_optional_imports()
    compat.install(reference_root=REFERENCE_ROOT)
    box = {}
    real = argparse.ArgumentParser.parse_args

    def fake(self, *a, **k):
        box["args"] = real(self, *a, **k)
        raise _Captured()

    argparse.ArgumentParser.parse_args = fake
    old = sys.argv
    sys.argv = ["main"] + list(argv)
    try:
        runpy.run_module("unsupervised_keypoints.main", run_name="__main__")
    except _Captured:
        pass
    finally:
        argparse.ArgumentParser.parse_args = real
        sys.argv = old
    return vars(box["args"])


if __name__ == "__main__":
    d = capture_args(["--my_token", "TOKEN"])
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "main_args_defaults.json")
    with open(out, "w") as f:
        json.dump(d, f, indent=1, sort_keys=True)
    print("wrote", out, len(d), "flags")
