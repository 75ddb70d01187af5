#!/usr/bin/env python
"""Debug harness: every torch.empty() on the GPU comes back filled with NaN (floats) / 0xFF bytes, so that a kernel reading
memory nobody wrote shows up as NaN instead of depending on what the caching allocator recycled.
    python scripts/poison_empty.py tests/test_gpu_pipeline.py -k cfg5 -m gpu -x -q -s"""
import sys
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pytest
import torch

_empty = torch.empty


def poisoned(*a, **k):
    t = _empty(*a, **k)
    if t.is_cuda and t.numel():
        if t.dtype.is_floating_point:
            t.fill_(float("nan"))
        else:
            t.view(torch.uint8).fill_(0xFF) if t.dtype in (torch.uint8, torch.int8) else t.fill_(-1)
    return t


torch.empty = poisoned
sys.exit(pytest.main(sys.argv[1:]))
