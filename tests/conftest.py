import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "large-scale-vit-slam_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        with np.load(os.path.join(GOLDEN, name)) as z:
            return {k: (torch.from_numpy(z[k]) if z[k].ndim else z[k].item()) for k in z.files}
    return load


def rnd(seed, *shape, scale=1.0):
    """Same synthetic-input generator as oracle/make_golden.py (numpy PCG64 -> platform independent)."""
    g = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy(g.standard_normal(shape, dtype=np.float32) * np.float32(scale))


def rand01(seed, *shape):
    g = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy(g.random(shape, dtype=np.float32))
