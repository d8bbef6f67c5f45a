"""The oracle (oracle/hpg_oracle.c, a CPU restatement) against the reference's outputs.

(1) bit-for-bit against the committed golden vectors (tests/golden/, generated from the
    unmodified reference C by tests/golden/make_golden.py);
(2) bit-for-bit against oracle/_ref (the reference itself, compiled) on seeded random inputs,
    when that build is present;
(3) the healpy-free invariants the survey measured on the reference.
No GPU involved.
"""
import numpy as np
import pytest

from conftest import assert_bit_equal, load_golden

POW2 = [1, 2, 4, 32, 1024, 4096, 2 ** 20, 2 ** 29]
RING_ONLY = [3, 924, 100003]
CASES = [(n, True) for n in POW2] + [(n, False) for n in POW2 + RING_ONLY]


def tag(nside, nest):
    return "%d_%s" % (nside, "nest" if nest else "ring")


@pytest.mark.parametrize("nside,nest", CASES)
def test_angle_to_pixel_golden(oracle_port, nside, nest):
    g = load_golden("angle_to_pixel")
    t = tag(nside, nest)
    assert_bit_equal(oracle_port.angle_to_pixel(nside, g["lon"], g["lat"], nest=nest), g["deg_" + t], "deg")
    assert_bit_equal(oracle_port.angle_to_pixel(nside, np.radians(g["lon"]), np.radians(g["lat"]), nest=nest,
                                                degrees=False), g["rad_" + t], "rad")
    assert_bit_equal(oracle_port.angle_to_pixel(nside, g["theta"], g["phi"], nest=nest, lonlat=False),
                     g["tp_" + t], "thetaphi")


@pytest.mark.parametrize("nside,nest", CASES)
def test_pixel_to_angle_golden(oracle_port, nside, nest):
    g = load_golden("pixel_to_angle")
    t = tag(nside, nest)
    pix = g["pix_%d" % nside]
    for key, kw in (("deg", {}), ("rad", {"degrees": False}), ("tp", {"lonlat": False})):
        a, b = oracle_port.pixel_to_angle(nside, pix, nest=nest, **kw)
        assert_bit_equal(a, g["%s_a_%s" % (key, t)], key + " a")
        assert_bit_equal(b, g["%s_b_%s" % (key, t)], key + " b")


@pytest.mark.parametrize("nside,nest", CASES)
def test_pixel_to_vector_golden(oracle_port, nside, nest):
    g = load_golden("pixel_to_vector")
    t = tag(nside, nest)
    x, y, z = oracle_port.pixel_to_vector(nside, g["pix_%d" % nside], nest=nest)
    assert_bit_equal(x, g["x_" + t], "x")
    assert_bit_equal(y, g["y_" + t], "y")
    assert_bit_equal(z, g["z_" + t], "z")


@pytest.mark.parametrize("nside,nest", CASES)
def test_vector_to_pixel_golden(oracle_port, nside, nest):
    g = load_golden("vector_to_pixel")
    assert_bit_equal(oracle_port.vector_to_pixel(nside, g["x"], g["y"], g["z"], nest=nest), g[tag(nside, nest)])


@pytest.mark.parametrize("nside", POW2)
def test_nest_ring_golden(oracle_port, nside):
    g = load_golden("nest_ring")
    pix = g["pix_%d" % nside]
    assert_bit_equal(oracle_port.nest_to_ring(nside, pix), g["n2r_%d" % nside], "n2r")
    assert_bit_equal(oracle_port.ring_to_nest(nside, pix), g["r2n_%d" % nside], "r2n")


@pytest.mark.parametrize("nside,nest", CASES)
def test_neighbors_golden(oracle_port, nside, nest):
    g = load_golden("neighbors")
    assert_bit_equal(oracle_port.neighbors(nside, g["pix_%d" % nside], nest=nest), g["nb_" + tag(nside, nest)])


@pytest.mark.parametrize("nside,nest", [(1, True), (1, False), (2, True), (2, False), (32, True), (32, False),
                                        (4096, True), (4096, False), (2 ** 20, True), (2 ** 20, False),
                                        (2 ** 29, True), (2 ** 29, False), (3, False), (924, False)])
def test_interp_golden(oracle_port, nside, nest):
    g = load_golden("interp")
    t = tag(nside, nest)
    p, w = oracle_port.get_interpolation_weights(nside, g["lon"], g["lat"], nest=nest)
    assert_bit_equal(p, g["pix_" + t], "pixels")
    assert_bit_equal(w, g["wgt_" + t], "weights")


@pytest.mark.parametrize("nside,nup", [(1, 1), (1, 4), (32, 64), (4096, 8192), (2 ** 20, 2 ** 22), (2 ** 28, 2 ** 29)])
@pytest.mark.parametrize("nest", [True, False])
def test_upgrade_golden(oracle_port, nside, nup, nest):
    g = load_golden("upgrade")
    out = oracle_port.upgrade_pixels(nside, g["pix_%d_%d" % (nside, nup)], nup, nest=nest)
    assert_bit_equal(out, g["up_%d_%d_%s" % (nside, nup, "nest" if nest else "ring")])


def test_upgrade_last_pixel_quirk(oracle_port):
    # hpgeom/hpgeom.py:505 range-checks p+1, so the last pixel raises in the reference
    with pytest.raises(ValueError, match="out of range"):
        oracle_port.upgrade_pixels(1024, np.array([12 * 1024 * 1024 - 1]), 2048)


def test_lonlat_golden(oracle_port):
    g = load_golden("lonlat")
    th, ph = oracle_port.lonlat_to_thetaphi(g["lon"], g["lat"])
    assert_bit_equal(th, g["deg_theta"])
    assert_bit_equal(ph, g["deg_phi"])
    th, ph = oracle_port.lonlat_to_thetaphi(g["rad_lon"], g["rad_lat"], degrees=False)
    assert_bit_equal(th, g["rad_theta"])
    assert_bit_equal(ph, g["rad_phi"])


def test_pixel_ranges(oracle_port):
    out = oracle_port.pixel_ranges_to_pixels(np.array([[0, 3], [10, 10], [20, 22]]))
    assert out.tolist() == [0, 1, 2, 20, 21]
    out = oracle_port.pixel_ranges_to_pixels(np.array([[0, 3], [10, 10]]), inclusive=True)
    assert out.tolist() == [0, 1, 2, 3, 10]
    with pytest.raises(ValueError):
        oracle_port.pixel_ranges_to_pixels(np.array([[5, 3]]))


# --------------------------------------------------------------------------------------
# against the compiled reference itself, seeded random inputs
# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("nside,nest", [(2, True), (1024, True), (1024, False), (2 ** 29, True), (2 ** 29, False),
                                        (924, False)])
def test_port_vs_compiled_reference(oracle_port, oracle_ref, nside, nest):
    if oracle_ref is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(nside + int(nest))
    n = 200_000
    lon = rng.uniform(0, 360, n)
    lat = np.degrees(np.arcsin(rng.uniform(-1, 1, n)))
    pix = rng.integers(0, 12 * nside * nside, n)
    x, y, z = rng.uniform(-1, 1, (3, n))
    assert_bit_equal(oracle_port.angle_to_pixel(nside, lon, lat, nest=nest),
                     oracle_ref.angle_to_pixel(nside, lon, lat, nest=nest), "a2p")
    for kw in ({}, {"degrees": False}, {"lonlat": False}):
        a, b = oracle_port.pixel_to_angle(nside, pix, nest=nest, **kw)
        ra, rb = oracle_ref.pixel_to_angle(nside, pix, nest=nest, **kw)
        assert_bit_equal(a, ra, "p2a a")
        assert_bit_equal(b, rb, "p2a b")
    assert_bit_equal(oracle_port.vector_to_pixel(nside, x, y, z, nest=nest),
                     oracle_ref.vector_to_pixel(nside, x, y, z, nest=nest), "v2p")
    for u, v in zip(oracle_port.pixel_to_vector(nside, pix, nest=nest), oracle_ref.pixel_to_vector(nside, pix, nest=nest)):
        assert_bit_equal(u, v, "p2v")
    assert_bit_equal(oracle_port.neighbors(nside, pix[:50000], nest=nest),
                     oracle_ref.neighbors(nside, pix[:50000], nest=nest), "neighbors")
    p, w = oracle_port.get_interpolation_weights(nside, lon[:50000], lat[:50000], nest=nest)
    rp, rw = oracle_ref.get_interpolation_weights(nside, lon[:50000], lat[:50000], nest=nest)
    assert_bit_equal(p, rp, "interp pix")
    assert_bit_equal(w, rw, "interp wgt")
    if nest:
        assert_bit_equal(oracle_port.nest_to_ring(nside, pix), oracle_ref.nest_to_ring(nside, pix), "n2r")
        assert_bit_equal(oracle_port.ring_to_nest(nside, pix), oracle_ref.ring_to_nest(nside, pix), "r2n")


# --------------------------------------------------------------------------------------
# healpy-free invariants
# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("nside", [1, 2, 32, 2 ** 15, 2 ** 29])
def test_invariants(oracle_port, nside):
    rng = np.random.default_rng(nside)
    npix = 12 * nside * nside
    pix = np.arange(npix) if npix <= 200_000 else rng.integers(0, npix, 200_000)
    assert np.array_equal(oracle_port.ring_to_nest(nside, oracle_port.nest_to_ring(nside, pix)), pix)
    assert np.array_equal(oracle_port.nest_to_ring(nside, oracle_port.ring_to_nest(nside, pix)), pix)
    for nest in (True, False):
        for kw in ({}, {"degrees": False}, {"lonlat": False}):
            a, b = oracle_port.pixel_to_angle(nside, pix, nest=nest, **kw)
            assert np.array_equal(oracle_port.angle_to_pixel(nside, a, b, nest=nest, **kw), pix)
        x, y, z = oracle_port.pixel_to_vector(nside, pix, nest=nest)
        assert np.array_equal(oracle_port.vector_to_pixel(nside, x, y, z, nest=nest), pix)
    if npix <= 200_000:
        nb = oracle_port.neighbors(nside, pix, nest=True)
        assert np.sum(nb == -1) == (24 if nside > 1 else 24)
