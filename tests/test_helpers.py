"""Host-side helpers of the drop-in API that need no GPU: the scalar nside utilities the reference itself
implements in numpy (hpgeom/hpgeom.py:190-339), against golden vectors generated from the unmodified reference
(tests/golden/make_golden_next.py).  thetaphi_to_lonlat / angle_to_vector / vector_to_angle / reorder run
through CUDA kernels and are tested in the gpu tier (tests/test_gpu_next.py)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN as GOLDEN_DIR, assert_bit_equal

import hpgeom_b200 as hpg
from hpgeom_b200 import healpy_compat as hpc


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN_DIR, "golden_next.npz"))


def test_scalar_utilities(g):
    ns = (1, 2, 1024, 2 ** 29)
    assert_bit_equal(np.array([hpg.nside_to_pixel_area(n) for n in ns]), g["h_area_deg"])
    assert_bit_equal(np.array([hpg.nside_to_pixel_area(n, degrees=False) for n in ns]), g["h_area_rad"])
    res = np.array([[hpg.nside_to_resolution(n, units=u) for u in ("radians", "degrees", "arcminutes", "arcseconds")]
                    for n in (1, 64, 2 ** 29)])
    assert_bit_equal(res, g["h_res"])
    assert_bit_equal(hpg.nside_to_order(np.array([1, 2, 1024, 2 ** 29])), g["h_order"])
    assert hpg.nside_to_order(1024) == 10 and isinstance(hpg.nside_to_order(1024), int)
    assert_bit_equal(hpg.npixel_to_nside(12 * np.array([1, 4, 1024 ** 2, 2 ** 58])), g["h_npix2nside"])
    assert hpg.npixel_to_nside(12 * 64 * 64) == 64
    assert hpg.order_to_nside(6) == 64
    got = np.array([hpc.nside2npix(64), hpc.npix2nside(12 * 64 * 64), hpc.nside2pixarea(64), hpc.nside2pixarea(64, degrees=True),
                    hpc.nside2resol(64), hpc.nside2resol(64, arcmin=True), hpc.nside2order(64), hpc.order2nside(6)], dtype=np.float64)
    assert_bit_equal(got, g["hp_scalars"][:8])


def test_error_texts():
    with open(os.path.join(GOLDEN_DIR, "golden_next_errors.json")) as f:
        errs = json.load(f)
    calls = {
        "order_bad": lambda: hpg.nside_to_order(3),
        "o2n_bad": lambda: hpg.order_to_nside(30),
        "npix_bad": lambda: hpg.npixel_to_nside(13),
        "npix_neg": lambda: hpg.npixel_to_nside(-12),
        "res_units": lambda: hpg.nside_to_resolution(4, units="furlongs"),
        "reorder_bad": lambda: hpg.reorder(np.zeros(13)),
        "bnd_step0": lambda: hpg.boundaries(32, 1, step=0),
    }
    for name, fn in calls.items():
        kind, msg = errs[name]
        with pytest.raises(ValueError) as e:
            fn()
        assert kind == "ValueError" and str(e.value) == msg, name


def test_region_queries_not_built():
    with pytest.raises(NotImplementedError):
        hpc.query_polygon(32, np.eye(3))
    for name in ("query_polygon", "query_polygon_vec", "query_ellipse", "query_box"):
        with pytest.raises(NotImplementedError, match="outside the accelerated path"):
            getattr(hpg, name)(32, np.eye(3))
