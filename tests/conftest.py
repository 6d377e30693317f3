import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def lib_built():
    """Build (if stale) and load the CUDA library; never falls back to anything else."""
    from bsplineinterpolation_b200 import build as _b
    _b.build()
    import bsplineinterpolation_b200 as pkg
    pkg.lib()
    return pkg


def rel_err(a, b):
    """The reference's own metric (test/src/include/rel_err.hpp:11-49)."""
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return float(np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()))
