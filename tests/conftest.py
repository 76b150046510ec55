import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def _load(name):
    z = np.load(os.path.join(GOLDEN, name))
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


@pytest.fixture(scope="session")
def golden_retrieval():
    return _load("retrieval.npz")


@pytest.fixture(scope="session")
def golden_model():
    return _load("model.npz")


def golden_checksum(*arrays) -> float:
    """Same formula as oracle/make_golden.py::checksum."""
    s = 0.0
    for a in arrays:
        a = np.asarray(a, np.float64).ravel()
        s += float((a * (1.0 + (np.arange(a.size) % 7))).sum())
    return s


@pytest.fixture(scope="session")
def have_gpu():
    import torch
    return torch.cuda.is_available()
