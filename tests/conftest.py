import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = eval(str(z["cfg"]))  # noqa: S307 - our own fixture
    return cfg, z


def relerr(a, b):
    """Relative Frobenius error ||a-b|| / ||b|| (0 when both are empty/zero)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = np.linalg.norm(a - b)
    nb = np.linalg.norm(b)
    return d / nb if nb > 0 else d


@pytest.fixture(scope="session")
def golden_cases():
    return ["c1", "c1_wscal", "c2_cut", "c3_cut", "c5_cut", "edge_odd"]
