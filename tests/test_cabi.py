"""The C-ABI shared library and the host-side logic that needs no GPU."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


@pytest.fixture(scope="module")
def L():
    from hpgeom_b200 import _lib
    return _lib.lib()


def test_exports_every_declared_symbol(L):
    hdr = open(os.path.join(ROOT, "include", "hpgb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(hpgb_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 30
    for n in sorted(names):
        assert hasattr(L, n), "libhpgb.so does not export %s" % n
    from hpgeom_b200 import _lib
    assert names == set(_lib.EXPORTS), "ctypes signatures out of sync with include/hpgb.h"


def test_no_oracle_in_product():
    """The product must never import / link the oracle or the host simulation."""
    pkg = os.path.join(ROOT, "hpgeom_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f)).read()
                for word in ("import oracle", "from oracle", "liboracle", "libhostsim", "hostsim"):
                    if word in ("hostsim",) and f.endswith((".h", ".cuh", ".cu")):
                        continue  # headers mention tests/hostsim in comments
                    assert word not in txt, "%s mentions %s" % (f, word)
    out = os.popen("ldd %s" % os.path.join(pkg, "libhpgb.so")).read()
    assert "oracle" not in out and "hostsim" not in out


def test_version_and_strings(L):
    assert L.hpgb_version() >= 100
    assert L.hpgb_error_string(0) == b"ok"
    assert b"power of 2" in L.hpgb_error_string(2)


def test_explain_matches_reference_messages():
    from hpgeom_b200 import _lib
    errs = json.load(open(os.path.join(GOLDEN, "golden_errors.json")))
    msgs = {e["msg"] for e in errs if e["msg"]}
    assert _lib.explain_nside(-1, True) in msgs
    assert _lib.explain_nside(2049, True) in msgs
    assert _lib.explain_nside(2 ** 30, False) in msgs
    assert _lib.explain_nside(2049, False) == ""
    assert _lib.explain_angle(1024, True, 0.0, 100.0, True, True) in msgs
    assert _lib.explain_angle(1024, True, 0.0, -90.00001, True, True) in msgs
    assert _lib.explain_angle(1024, True, 0.0, 1.6, True, False) in msgs
    assert _lib.explain_angle(1024, True, -0.1, 0.0, False, False) in msgs
    assert _lib.explain_angle(1024, True, 3.2, 0.0, False, False) in msgs
    assert _lib.explain_angle(1024, True, 1.0, -0.1, False, False) in msgs
    assert _lib.explain_angle(1024, True, 1.0, 6.3, False, False) in msgs
    assert _lib.explain_angle(1024, True, 10.0, 90.0, True, True) == ""
    assert _lib.explain_pixel(1024, True, -1) in msgs
    assert _lib.explain_pixel(1024, True, 12 * 1024 * 1024) in msgs
    assert _lib.explain_pixel(1024, True, 5) == ""


def test_argument_checks_without_gpu():
    """Everything the reference rejects before its loop must raise the same way here, GPU or not."""
    import hpgeom_b200 as hpg
    with pytest.raises(ValueError, match=r"nside, a, b arrays could not be broadcast together"):
        hpg.angle_to_pixel(1024, np.zeros(3), np.zeros(4))
    with pytest.raises(ValueError, match=r"nside, x, y, z arrays could not be broadcast together"):
        hpg.vector_to_pixel(1024, np.zeros(2), np.zeros(3), np.zeros(2))
    with pytest.raises(ValueError, match=r"nside .* must be positive"):
        hpg.angle_to_pixel(-1, 0.0, 0.0)
    with pytest.raises(ValueError, match=r"nside .* must be power of 2"):
        hpg.angle_to_pixel(2049, 0.0, 0.0, nest=True)
    with pytest.raises(ValueError, match=r"nside .* must not be greater"):
        hpg.angle_to_pixel(2 ** 30, 0.0, 0.0, nest=False)
    with pytest.raises(TypeError, match="Cannot cast"):
        hpg.pixel_to_angle(1024, np.array([0.5]))
    with pytest.raises(ValueError, match=r"pix array must be at most 1D"):
        hpg.neighbors(1024, np.zeros((2, 2), dtype=np.int64))
    with pytest.raises(ValueError, match=r"nside array must be at most 1D"):
        hpg.neighbors(np.full((2, 2), 1024), np.zeros(2, dtype=np.int64))
    with pytest.raises(ValueError, match=r"array must be at most 1D"):
        hpg.get_interpolation_weights(1024, np.zeros((2, 2)), np.zeros((2, 2)))
    with pytest.raises(ValueError, match=r"must have same number of dimensions"):
        hpg.get_interpolation_weights(1024, np.zeros(2), 0.0)
    with pytest.raises(ValueError, match=r"nside_upgrade must be >= nside"):
        hpg.upgrade_pixels(1024, np.array([5]), 512)
    with pytest.raises(ValueError, match=r"Input nside must be a power of two"):
        hpg.upgrade_pixels(1000, np.array([5]), 2000)
    with pytest.raises(ValueError, match=r"nside_upgrade must be a power of two"):
        hpg.upgrade_pixels(1024, np.array([5]), 3000)
    with pytest.raises(ValueError, match=r"must be integers"):
        hpg.upgrade_pixels(1024.0, np.array([1]), 2048)
    # empty inputs never touch the device and never check nside (the reference's loop never runs)
    assert hpg.angle_to_pixel(-5, np.zeros(0), np.zeros(0)).shape == (0,)
    assert hpg.neighbors(1024, np.zeros(0, dtype=np.int64)).shape == (0, 8)
    p, w = hpg.get_interpolation_weights(1024, np.zeros(0), np.zeros(0))
    assert p.shape == (0, 4) and w.shape == (0, 4) and p.dtype == np.int64 and w.dtype == np.float64
    # positional flag order is lonlat, nest, degrees (hpgeom/hpgeom.c:183-184)
    import inspect
    assert list(inspect.signature(hpg.angle_to_pixel).parameters) == ["nside", "a", "b", "lonlat", "nest", "degrees", "n_threads"]
    assert list(inspect.signature(hpg.pixel_to_angle).parameters) == ["nside", "pix", "lonlat", "nest", "degrees", "n_threads"]
    assert list(inspect.signature(hpg.get_interpolation_weights).parameters) == ["nside", "a", "b", "lonlat", "nest", "degrees", "n_threads"]


def test_pixel_ranges_to_pixels():
    import hpgeom_b200 as hpg
    # the expansion itself runs on the GPU (tests/test_query.py, -m gpu); host-side checks only here
    assert hpg.pixel_ranges_to_pixels(np.zeros((0, 2), dtype=np.int64)).shape == (0,)
    with pytest.raises(ValueError, match=r"must be 2D"):
        hpg.pixel_ranges_to_pixels(np.array([1, 2, 3]))
    r = hpg.upgrade_pixel_ranges(32, np.array([[0, 1], [5, 7]]), 64)
    assert r.tolist() == [[0, 4], [20, 28]]


def test_no_cpu_fallback_without_device():
    """On a box without a GPU the conversions must fail loudly, not fall back."""
    import hpgeom_b200 as hpg
    from hpgeom_b200 import _lib
    if _lib.lib().hpgb_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        hpg.nest_to_ring(1024, np.arange(10))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        hpg.angle_to_pixel(1024, np.zeros(4), np.zeros(4))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        hpg.pixel_ranges_to_pixels(np.array([[0, 3], [10, 10], [20, 22]]))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        hpg.query_circle(1024, 10.0, 20.0, 1.0, nest=False)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        hpg.query_circle(1024, 10.0, 20.0, 1.0)


def test_host_copy_pool_is_a_memcpy():
    """hpgb_host_copy -- the parallel staging copy of the hpgb_host_* pipelines -- on sizes around its split points,
    odd sizes and misaligned pointers (host only: no GPU involved)."""
    import ctypes

    import numpy as np

    from hpgeom_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(3)
    for nbytes in (0, 1, 63, 64, 65, (1 << 20) - 1, (1 << 20) + 1, 3 * (1 << 20) + 17, 37 * (1 << 20) + 5):
        for shift_src, shift_dst in ((0, 0), (1, 0), (0, 3), (5, 7)):
            src = rng.integers(0, 255, nbytes + 16, dtype=np.uint8)
            dst = np.full(nbytes + 32, 0xAB, dtype=np.uint8)
            rc = L.hpgb_host_copy(ctypes.c_void_p(dst.ctypes.data + 8 + shift_dst), ctypes.c_void_p(src.ctypes.data + shift_src), nbytes)
            assert rc == 0
            assert np.array_equal(dst[8 + shift_dst:8 + shift_dst + nbytes], src[shift_src:shift_src + nbytes])
            assert (dst[:8 + shift_dst] == 0xAB).all() and (dst[8 + shift_dst + nbytes:] == 0xAB).all()
    assert L.hpgb_host_copy(None, None, -1) != 0
