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


def load_golden(name):
    return np.load(os.path.join(GOLDEN, "golden_%s.npz" % name))


def bits(a):
    """View float64 arrays as int64 so comparisons are bit-exact (and NaN-safe)."""
    a = np.ascontiguousarray(a)
    return a.view(np.int64) if a.dtype == np.float64 else a


def assert_bit_equal(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, "%s shape %s vs %s" % (what, a.shape, b.shape)
    assert a.dtype == b.dtype, "%s dtype %s vs %s" % (what, a.dtype, b.dtype)
    bad = np.flatnonzero(bits(a).ravel() != bits(b).ravel())
    assert bad.size == 0, "%s: %d of %d differ, first at %d: %r vs %r" % (
        what, bad.size, a.size, bad[0], a.ravel()[bad[0]], b.ravel()[bad[0]])


def ulp_diff(a, b):
    """|a-b| in units of the spacing of b (float64)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    sp = np.spacing(np.maximum(np.abs(b), np.finfo(np.float64).tiny))
    return np.abs(a - b) / sp


@pytest.fixture(scope="session")
def oracle_port():
    import oracle
    return oracle.port()


@pytest.fixture(scope="session")
def oracle_ref():
    import oracle
    return oracle.ref
