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
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_models():
    with open(os.path.join(GOLDEN, "models.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle_models(golden_models):
    from oracle import rbd
    return {k: rbd.Model(v) for k, v in golden_models.items()}


def load_npz(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))
