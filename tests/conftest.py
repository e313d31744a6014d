import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, GOLDEN):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class Golden:
    """Lazy reader of tests/golden/<group>.npz (outputs of the unmodified reference)."""

    def __init__(self):
        self._cache = {}

    def __call__(self, group):
        if group not in self._cache:
            self._cache[group] = np.load(os.path.join(GOLDEN, group + ".npz"))
        return self._cache[group]


@pytest.fixture(scope="session")
def golden():
    return Golden()


def rel_err(a, b):
    """max |a-b| relative to the rms of the reference field b (SURVEY 8c tolerance definition)."""
    dt = np.complex128 if (np.iscomplexobj(a) or np.iscomplexobj(b)) else np.float64
    a = np.asarray(a, dtype=dt)
    b = np.asarray(b, dtype=dt)
    scale = np.sqrt(np.mean(np.abs(b) ** 2))
    return float(np.max(np.abs(a - b)) / (scale if scale > 0 else 1.0))


def assert_close(a, b, tol, what=""):
    assert np.shape(a) == np.shape(b), f"{what}: shape {np.shape(a)} vs {np.shape(b)}"
    e = rel_err(a, b)
    assert e <= tol, f"{what}: max|diff|/rms(ref) = {e:.3e} > {tol:.1e}"
