import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_NAMES = ["room_b1_c3_8", "room_b2_c16", "room_b3_c32_64"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a box without CUDA skips the GPU parity tests instead of failing in them."""
    if have_cuda():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (there is no CPU fallback to test)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def unpack(flat, off):
    return [flat[off[k]:off[k + 1]] for k in range(len(off) - 1)]


def rel_err(a, b):
    """max-norm relative error |a-b|_inf / |b|_inf -- the measure every floating-point tolerance in tests/ refers to."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(params=GOLDEN_NAMES)
def golden(request):
    return load_golden(request.param)


def have_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
