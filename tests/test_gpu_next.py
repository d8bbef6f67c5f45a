"""GPU parity of the rows next to the hot path: boundaries, max_pixel_radius,
reorder, the helpers that run through the kernels, and the healpy_compat shims -- against golden
vectors generated from the unmodified reference (tests/golden/make_golden_next.py)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN as GOLDEN_DIR, assert_bit_equal, ulp_diff

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hpg():
    import hpgeom_b200
    return hpgeom_b200


@pytest.fixture(scope="module")
def hpc():
    from hpgeom_b200 import healpy_compat
    return healpy_compat


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN_DIR, "golden_next.npz"))


def close_angles(got, ref, scale=1.0, frac=0.01):
    """theta-derived outputs: one ulp of theta (pi-sized), mostly bit-identical."""
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1.5 * np.spacing(np.pi) * scale
    assert got.size < 500 or np.mean(got != ref) <= frac


@pytest.mark.parametrize("nside", [1, 2, 32, 1024, 2 ** 20, 2 ** 29, 3, 924])
def test_boundaries_golden(hpg, g, nside):
    pix = g["bnd_pix_%d" % nside]
    for nest in ([True, False] if nside & (nside - 1) == 0 else [False]):
        for step in (1, 2, 5):
            tag = "bnd_%d_%s_%d_" % (nside, "nest" if nest else "ring", step)
            th, ph = hpg.boundaries(nside, pix, step=step, nest=nest, lonlat=False)
            assert_bit_equal(ph, g[tag + "tp_b"], "phi")
            close_angles(th, g[tag + "tp_a"])
            lon, lat = hpg.boundaries(nside, pix, step=step, nest=nest)
            assert_bit_equal(lon, g[tag + "deg_a"], "lon")
            close_angles(lat, g[tag + "deg_b"], 180.0 / np.pi)
            lon, lat = hpg.boundaries(nside, pix, step=step, nest=nest, degrees=False)
            assert_bit_equal(lon, g[tag + "rad_a"], "lon")
            close_angles(lat, g[tag + "rad_b"])


def test_boundaries_shapes_errors_and_device_tensors(hpg, g):
    import torch
    a, b = hpg.boundaries(1024, 77, step=3)
    assert a.shape == (12,) and b.shape == (12,)
    assert_bit_equal(a, g["bnd_scalar_a"])
    close_angles(b, g["bnd_scalar_b"], 180.0 / np.pi)
    with open(os.path.join(GOLDEN_DIR, "golden_next_errors.json")) as f:
        errs = json.load(f)
    for name, fn in (("bnd_2d", lambda: hpg.boundaries(32, np.zeros((2, 2), dtype=np.int64))),
                     ("bnd_badpix", lambda: hpg.boundaries(32, 12 * 32 * 32)),
                     ("bnd_nest_nonpow2", lambda: hpg.boundaries(30, 1)),
                     ("mpr_bad", lambda: hpg.max_pixel_radius(0)),
                     ("mpr_big", lambda: hpg.max_pixel_radius(2 ** 30))):
        with pytest.raises(ValueError) as e:
            fn()
        assert str(e.value) == errs[name][1], name
    pix = g["bnd_pix_1024"]
    ta, tb = hpg.boundaries(1024, torch.from_numpy(pix).cuda(), step=2)
    na, nb = hpg.boundaries(1024, pix, step=2)
    assert ta.is_cuda and np.array_equal(ta.cpu().numpy(), na) and np.array_equal(tb.cpu().numpy(), nb)
    # first bad pixel wins, on a large array
    big = np.zeros(500_000, dtype=np.int64)
    big[400_000] = -5
    big[450_000] = -6
    with pytest.raises(ValueError, match=r"Pixel value -5 out of range"):
        hpg.boundaries(64, big)


def test_max_pixel_radius(hpg, g):
    for key, kw in (("deg", {}), ("rad", {"degrees": False})):
        got = hpg.max_pixel_radius(g["mpr_nside"], **kw)
        assert ulp_diff(got, g["mpr_" + key]).max() <= 1.0
        assert np.mean(got != g["mpr_" + key]) <= 0.1
    assert isinstance(float(hpg.max_pixel_radius(64)), float) and np.ndim(hpg.max_pixel_radius(64)) == 0


def test_reorder_and_helpers_through_the_kernels(hpg, g):
    import torch
    for nside in (1, 4, 32):
        m = g["reorder_in_%d" % nside]
        assert_bit_equal(hpg.reorder(m, ring_to_nest=True), g["reorder_r2n_%d" % nside])
        assert_bit_equal(hpg.reorder(m, ring_to_nest=False), g["reorder_n2r_%d" % nside])
        t = hpg.reorder(torch.from_numpy(m).cuda(), ring_to_nest=True)
        assert t.is_cuda and np.array_equal(t.cpu().numpy(), g["reorder_r2n_%d" % nside])
    # a full-size map on the device: reorder there and back, and against the permutation kernels
    nside = 2048
    m = torch.arange(12 * nside * nside, dtype=torch.float64, device="cuda")
    r2n = hpg.reorder(m, ring_to_nest=True)
    assert torch.equal(hpg.reorder(r2n, ring_to_nest=False), m)
    pixels = torch.arange(12 * nside * nside, dtype=torch.int64, device="cuda")
    want = torch.empty_like(m)
    want[hpg.ring_to_nest(nside, pixels)] = m
    assert torch.equal(r2n, want)
    # every supported element size, small and block-sized faces (nside 64 = one 64 x 64 block per face)
    rng = np.random.default_rng(3)
    for ns in (1, 2, 16, 64, 128):
        npix = 12 * ns * ns
        pix = np.arange(npix)
        for dt in (np.uint8, np.int16, np.float32, np.float64, np.complex128):
            mm = rng.integers(0, 200, npix).astype(dt)
            for r2n_flag in (True, False):
                ref = np.zeros_like(mm)
                ref[(hpg.ring_to_nest if r2n_flag else hpg.nest_to_ring)(ns, pix)] = mm
                assert np.array_equal(hpg.reorder(mm, ring_to_nest=r2n_flag), ref), (ns, dt, r2n_flag)
    assert hpg.reorder(np.arange(12 * 4 * 4).reshape(12, 16), ring_to_nest=True).shape == (12, 16)


def test_angle_vector_helpers_golden(hpg, hpc, g):
    """thetaphi_to_lonlat / angle_to_vector / vector_to_angle kernels against the reference's numpy results
    (golden_next.npz): pure-IEEE outputs bit for bit; sin / cos products bit-identical wherever glibc rounds
    correctly (> 99 %), never more than 1 ulp apart; acos / atan2 within 2 ulp of numpy's SVML routines.
    numpy inputs and CUDA tensors give the same bits."""
    import torch
    for key, kw in (("deg", {}), ("rad", {"degrees": False})):
        a, b = hpg.thetaphi_to_lonlat(g["h_theta"], g["h_phi"], **kw)
        assert_bit_equal(a, g["h_tp2ll_%s_a" % key])
        assert_bit_equal(b, g["h_tp2ll_%s_b" % key])
        ta, tb = hpg.thetaphi_to_lonlat(torch.from_numpy(g["h_theta"]).cuda(), torch.from_numpy(g["h_phi"]).cuda(), **kw)
        assert ta.is_cuda and np.array_equal(ta.cpu().numpy(), a) and np.array_equal(tb.cpu().numpy(), b)

    def near(got, ref, max_ulp, frac):
        assert got.shape == ref.shape
        big = np.abs(ref) > 1e-300
        assert ulp_diff(got[big], ref[big]).max() <= max_ulp
        assert np.abs(got[~big]).max(initial=0.0) < 1e-15
        assert got.size < 200 or np.mean(got != ref) <= frac

    near(hpg.angle_to_vector(g["h_lon"], g["h_lat"]), g["h_a2v_ll"], 1.0, 0.01)
    near(hpg.angle_to_vector(g["h_theta"], g["h_phi"], lonlat=False), g["h_a2v_tp"], 1.0, 0.01)
    near(np.asarray(hpg.angle_to_vector(10.0, 20.0)), g["h_a2v_scalar"], 1.0, 1.0)
    assert hpg.angle_to_vector(10.0, 20.0).shape == (3,)
    near(hpc.ang2vec(g["h_theta"], g["h_phi"]), g["hp_ang2vec"], 1.0, 0.01)
    tv = hpg.angle_to_vector(torch.from_numpy(g["h_lon"]).cuda(), torch.from_numpy(g["h_lat"]).cuda())
    assert tv.is_cuda and np.array_equal(tv.cpu().numpy(), hpg.angle_to_vector(g["h_lon"], g["h_lat"]))
    for kw, ka, kb in (({}, "h_v2a_ll_a", "h_v2a_ll_b"), ({"lonlat": False}, "h_v2a_tp_a", "h_v2a_tp_b")):
        a, b = hpg.vector_to_angle(g["h_vec"], **kw)
        if kw:
            near(a, g[ka], 2.0, 0.25)   # theta = acos
            near(b, g[kb], 2.0, 0.25)   # phi = atan2 mod 2 pi
        else:
            near(a, g[ka], 2.0, 0.25)   # lon
            assert np.abs(b - g[kb]).max() <= 2 * np.spacing(np.pi) * 180 / np.pi  # lat = 90 - theta: ulps of theta
        ta, tb = hpg.vector_to_angle(torch.from_numpy(g["h_vec"]).cuda(), **kw)
        assert np.array_equal(ta.cpu().numpy(), a) and np.array_equal(tb.cpu().numpy(), b)
    a, b = hpc.vec2ang(g["h_vec"])
    near(a, g["hp_vec2ang_a"], 2.0, 0.25)
    near(b, g["hp_vec2ang_b"], 2.0, 0.25)
    # exact cases: the poles and the axes
    ax = np.array([[0, 0, 1.0], [0, 0, -2.0], [1, 0, 0], [0, 3, 0], [-1, 0, 0], [0, -1, 0]])
    th, ph = hpg.vector_to_angle(ax, lonlat=False)
    assert np.array_equal(th, [0.0, np.pi, np.pi / 2, np.pi / 2, np.pi / 2, np.pi / 2])
    assert np.array_equal(ph, [0.0, 0.0, 0.0, np.pi / 2, np.pi, 1.5 * np.pi])
    with open(os.path.join(GOLDEN_DIR, "golden_next_errors.json")) as f:
        errs = json.load(f)
    for name, fn in (("a2v_theta", lambda: hpg.angle_to_vector(4.0, 1.0, lonlat=False)),
                     ("a2v_phi", lambda: hpg.angle_to_vector(1.0, 7.0, lonlat=False)),
                     ("a2v_theta", lambda: hpg.angle_to_vector(np.array([1.0, 1.0, 4.0]), np.array([7.0, 1.0, 1.0]), lonlat=False))):
        with pytest.raises(ValueError) as e:
            fn()
        assert str(e.value) == errs[name][1], name
    with pytest.raises(ValueError, match="Latitude out of range"):
        hpg.angle_to_vector(10.0, 91.0)


@pytest.mark.parametrize("nest", [True, False])
def test_healpy_compat(hpc, g, nest):
    nside = 256
    t = "nest" if nest else "ring"
    pix, theta, phi, lon, lat, vec = g["hp_pix"], g["h_theta"], g["h_phi"], g["h_lon"], g["h_lat"], g["h_vec"]
    assert_bit_equal(hpc.ang2pix(nside, theta, phi, nest=nest), g["hp_ang2pix_" + t])
    assert_bit_equal(hpc.ang2pix(nside, lon, lat, nest=nest, lonlat=True), g["hp_ang2pix_ll_" + t])
    a, b = hpc.pix2ang(nside, pix, nest=nest)
    close_angles(a, g["hp_pix2ang_a_" + t])
    assert_bit_equal(b, g["hp_pix2ang_b_" + t])
    a, b = hpc.pix2ang(nside, pix, nest=nest, lonlat=True)
    assert_bit_equal(a, g["hp_pix2ang_ll_a_" + t])
    close_angles(b, g["hp_pix2ang_ll_b_" + t], 180.0 / np.pi)
    x, y, z = hpc.pix2vec(nside, pix, nest=nest)
    ref = g["hp_pix2vec_" + t]
    assert_bit_equal(z, ref[2])
    assert ulp_diff(x, ref[0]).max() <= 2 and ulp_diff(y, ref[1]).max() <= 2
    assert_bit_equal(hpc.vec2pix(nside, vec[:, 0], vec[:, 1], vec[:, 2], nest=nest), g["hp_vec2pix_" + t])
    assert_bit_equal(hpc.get_all_neighbours(nside, pix, nest=nest), g["hp_neigh_pix_" + t])
    assert_bit_equal(hpc.get_all_neighbours(nside, theta, phi, nest=nest), g["hp_neigh_ang_" + t])
    assert_bit_equal(hpc.get_all_neighbours(nside, 100, nest=nest), g["hp_neigh_scalar_" + t])
    p, w = hpc.get_interp_weights(nside, theta, phi, nest=nest)
    assert_bit_equal(p, g["hp_interp_p_" + t])
    assert np.abs(w - g["hp_interp_w_" + t]).max() <= 4 * np.spacing(np.pi) * nside
    # pixel-centre form: the points sit EXACTLY on ring colatitudes, where ring_above(cos(theta)) is decided by the
    # last bit of cos(theta) (and theta itself is acos(z) to within 1 ulp on both sides).  Every row that differs
    # from the reference's must be explained by that: the reference returns our row for a <= 2 ulp change of theta.
    p, w = hpc.get_interp_weights(nside, pix, nest=nest)
    ref = g["hp_interp_pix_p_" + t]
    assert p.shape == ref.shape
    cols = np.flatnonzero((p != ref).any(axis=0))
    if cols.size:
        import oracle
        port = oracle.port()
        theta, phi = hpc.pix2ang(nside, pix[cols], nest=nest)
        explained = np.zeros(cols.size, dtype=bool)
        for d in (-2, -1, 0, 1, 2):
            th = np.clip(theta + d * np.spacing(theta), 0.0, np.pi)
            rp, _ = port.get_interpolation_weights(nside, th, phi, nest=nest, lonlat=False, degrees=False)
            explained |= (rp.T == p[:, cols]).all(axis=0)
        assert explained.all(), "%d interpolation rows differ without sitting on a ring boundary" % int((~explained).sum())
    assert cols.size <= 0.05 * pix.size
    b3 = hpc.boundaries(nside, pix[:20], step=2, nest=nest)
    assert b3.shape == g["hp_bnd_" + t].shape and np.abs(b3 - g["hp_bnd_" + t]).max() <= 4e-16
    b1 = hpc.boundaries(nside, 5, step=1, nest=nest)
    assert b1.shape == g["hp_bnd_scalar_" + t].shape and np.abs(b1 - g["hp_bnd_scalar_" + t]).max() <= 4e-16
    if nest:
        assert_bit_equal(hpc.nest2ring(nside, pix), g["hp_nest2ring"])
        assert_bit_equal(hpc.ring2nest(nside, pix), g["hp_ring2nest"])
        v = hpc.ang2vec(lon, lat, lonlat=True)  # sin / cos: glibc itself mis-rounds ~0.1 % of arguments
        assert ulp_diff(v, g["hp_ang2vec_ll"]).max() <= 1.0 and np.mean(v != g["hp_ang2vec_ll"]) <= 0.01
        assert ulp_diff(np.array([hpc.max_pixrad(64), hpc.max_pixrad(64, degrees=True)]), g["hp_scalars"][8:]).max() <= 1
