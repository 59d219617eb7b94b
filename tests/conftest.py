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
    config.addinivalue_line("markers", "reference: needs /root/reference + numba (build container only)")


@pytest.fixture(scope="session")
def golden_tables():
    return np.load(os.path.join(GOLDEN, "tables.npz"))


@pytest.fixture(scope="session")
def golden_percentiles():
    return np.load(os.path.join(GOLDEN, "percentiles.npz"))


@pytest.fixture(scope="session")
def golden_metrics():
    return np.load(os.path.join(GOLDEN, "metrics.npz"))


def bits_equal(a: np.ndarray, b: np.ndarray) -> bool:
    """Bit-exact equality of float64 arrays, treating every NaN as equal to every NaN and -0.0 as
    equal to +0.0 (which zero a min/selection returns depends on the visiting order of equal keys)."""
    a = np.asarray(a, np.float64) + 0.0
    b = np.asarray(b, np.float64) + 0.0
    if a.shape != b.shape:
        return False
    na, nb = np.isnan(a), np.isnan(b)
    if not np.array_equal(na, nb):
        return False
    return bool(np.array_equal(a[~na].view(np.uint64), b[~nb].view(np.uint64)))
