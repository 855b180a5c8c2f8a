import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    """Outputs of the unmodified reference (tests/golden/make_golden.py)."""
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz")))


@pytest.fixture(scope="session")
def golden_r2():
    """Round-2 outputs of the unmodified reference (tests/golden/make_golden_r2.py): cos(theta) clamp, lag tails."""
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors_r2.npz")))


def normwise(a, b):
    """Parity metric of SURVEY 7 / BASELINE.md 3: ||a-b||_inf / max(||b||_inf, 1)."""
    a = np.asarray(a, float)
    b = np.asarray(b, float)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1.0))
