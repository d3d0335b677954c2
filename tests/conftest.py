import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# tolerances of the north star: norm-wise relative error per matrix
TOL = {torch.float32: 1e-5, torch.float64: 1e-12}
TAGS = {torch.float32: "f32", torch.float64: "f64"}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    def __init__(self, path):
        self._z = np.load(path)

    def __call__(self, key, device="cpu"):
        return torch.from_numpy(self._z[key]).to(device)

    def has(self, key):
        return key in self._z.files


@pytest.fixture(scope="session")
def sym_golden():
    return Golden(os.path.join(GOLDEN, "sym_golden.npz"))


@pytest.fixture(scope="session")
def dense_golden():
    return Golden(os.path.join(GOLDEN, "dense_golden.npz"))


@pytest.fixture(scope="session")
def extra_golden():
    return Golden(os.path.join(GOLDEN, "extra_golden.npz"))
