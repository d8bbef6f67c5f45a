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


class _BestOracle:
    """The UNMODIFIED reference C (oracle/_ref) for every function it exports, the C restatement for the rest
    (upgrade_pixels and pixel_ranges_to_pixels are Python in the reference) or when that build is absent.
    `name` says which one answered, for the assertion messages."""

    _REF_FUNCS = ("angle_to_pixel", "pixel_to_angle", "nest_to_ring", "ring_to_nest", "vector_to_pixel",
                  "pixel_to_vector", "neighbors", "get_interpolation_weights")

    def __init__(self, ref, port):
        self._ref, self._port = ref, port
        self.name = "oracle/_ref (unmodified reference C)" if ref is not None else "oracle port (C restatement)"

    def __getattr__(self, fn):
        if self._ref is not None and fn in self._REF_FUNCS and hasattr(self._ref, fn):
            return getattr(self._ref, fn)
        return getattr(self._port, fn)


@pytest.fixture(scope="session")
def oracle_best(oracle_port):
    import oracle
    return _BestOracle(oracle.ref, oracle_port)
