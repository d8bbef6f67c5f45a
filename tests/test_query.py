"""query_circle (RING) and pixel_ranges_to_pixels: SURVEY.md 8(a) row a18 and 8(f) row 4.

CPU tier: the per-ring element function + host plan (compiled for the host, tests/hostsim) against
golden vectors frozen from the unmodified reference (tests/golden/make_golden_query.py), and against
the live reference where oracle/_ref is present.  GPU tier: the public API through the C ABI.
"""
import numpy as np
import pytest

from conftest import assert_bit_equal, load_golden


def cases():
    g = load_golden("query")
    ends = np.cumsum(g["qc_sizes"])
    for k, p in enumerate(g["qc_params"]):
        nside, a, b, radius, fact, lonlat, degrees = p
        yield (int(nside), float(a), float(b), float(radius), int(fact), bool(lonlat), bool(degrees)), \
            g["qc_pix"][ends[k] - g["qc_sizes"][k]:ends[k]]


def nest_cases():
    g = load_golden("query")
    pe, re_ = np.cumsum(g["qn_sizes"]), np.cumsum(g["qn_rsizes"])
    for k, p in enumerate(g["qn_params"]):
        nside, theta, phi, radius, fact = p
        yield (int(nside), float(theta), float(phi), float(radius), int(fact)), \
            g["qn_pix"][pe[k] - g["qn_sizes"][k]:pe[k]], g["qn_ranges"][re_[k] - g["qn_rsizes"][k]:re_[k]]


def expand(table):
    parts = [np.arange(lo, hi, dtype=np.int64) for lo, hi in table if hi > lo]
    return np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64)


@pytest.fixture(scope="module")
def sim():
    from hostsim import HostSim
    return HostSim()


def test_hostsim_ring_tables_match_golden(sim):
    """theta/phi cases only (the lonlat forms go through the C ABI plan, GPU tier)."""
    n = 0
    for (nside, a, b, radius, fact, lonlat, degrees), want in cases():
        if lonlat:
            continue
        table = sim.query_circle_ring_ranges(nside, a, b, radius, fact)
        assert np.all(table[:, 1] >= table[:, 0])
        nz = table[table[:, 1] > table[:, 0]]
        assert np.all(nz[1:, 0] >= nz[:-1, 1]), "rows must be ascending and disjoint"
        assert_bit_equal(expand(table), want, "nside=%d fact=%d" % (nside, fact))
        n += 1
    assert n > 500


def test_hostsim_ring_vs_live_reference(sim, oracle_ref):
    if oracle_ref is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(5)
    for nside in (4, 37, 512, 2 ** 14):
        for _ in range(25):
            theta, phi = float(np.arccos(rng.uniform(-1, 1))), float(rng.uniform(0, 2 * np.pi))
            radius = float(10 ** rng.uniform(-3.5, 0.3)) if nside < 2 ** 12 else float(10 ** rng.uniform(-4, -2))
            for fact in (0, 1, 4, 5):
                want = oracle_ref.query_circle(nside, theta, phi, radius, inclusive=fact > 0, fact=max(fact, 1),
                                               nest=False, lonlat=False)
                assert_bit_equal(expand(sim.query_circle_ring_ranges(nside, theta, phi, radius, fact)), want)


def test_hostsim_nest_walk_matches_golden(sim):
    """The breadth-first walk with hpgb_qn_visit gives the reference's (depth-first) range set."""
    n = 0
    for (nside, theta, phi, radius, fact), pix, ranges in nest_cases():
        got = sim.query_circle_nest_ranges(nside, theta, phi, radius, fact)
        assert_bit_equal(got, ranges, "nside=%d fact=%d" % (nside, fact))
        assert_bit_equal(expand(got), pix)
        n += 1
    assert n > 400


def test_plan_argument_checks():
    """The plan is host-only code: the wrapper's checks, in its order (hpgeom/hpgeom.c:793-838)."""
    import hpgeom_b200 as hpg
    with pytest.raises(ValueError, match=r"lat = 100 out of range \[-90, 90\]"):
        hpg.query_circle(1024, 0.0, 100.0, -1.0, nest=False)
    with pytest.raises(ValueError, match=r"colatitude \(theta\) = -0.1 out of range"):
        hpg.query_circle(1024, -0.1, 0.0, 1.0, nest=False, lonlat=False)
    with pytest.raises(ValueError, match="Radius must be positive."):
        hpg.query_circle(0, 0.0, 0.0, 0.0, nest=False)
    with pytest.raises(ValueError, match="nside"):
        hpg.query_circle(0, 0.0, 0.0, 1.0, nest=False)
    with pytest.raises(ValueError, match="power of 2"):
        hpg.query_circle(1000, 0.0, 0.0, 1.0, nest=True)
    with pytest.raises(ValueError, match="Inclusive factor 0 must be positive."):
        hpg.query_circle(1024, 0.0, 0.0, 1.0, inclusive=True, fact=0, nest=False)
    with pytest.raises(ValueError, match=r"Inclusive factor \* nside must be <= 536870912"):
        hpg.query_circle(2 ** 29, 0.0, 0.0, 1.0, inclusive=True, fact=2, nest=False)
    with pytest.raises(ValueError, match="Inclusive factor 3 must be power of 2 for nest."):
        hpg.query_circle(1024, 0.0, 0.0, 1.0, inclusive=True, fact=3, nest=True)
    with pytest.raises(RuntimeError, match="return_pixel_ranges"):
        hpg.query_circle(1024, 0.0, 0.0, 1.0, nest=False, return_pixel_ranges=True)
    with pytest.raises(ValueError, match="vec must be 3 elements."):
        hpg.query_circle_vec(1024, [1.0, 0.0], 0.1)
    with pytest.raises(ValueError, match=r"pixel_ranges must be 2D, with shape \(M, 2\)."):
        hpg.pixel_ranges_to_pixels(np.zeros(4, dtype=np.int64))
    assert hpg.pixel_ranges_to_pixels(np.zeros((0, 2), dtype=np.int64)).shape == (0,)


def test_batch_planning_is_host_only():
    """hpgb_query_circle_ring_plan_many / _batch_prepare run on the host: per-disc plans equal the single-disc
    plan, the first rejected disc is reported, whole-sphere discs are normalised, ring_start is the prefix sum."""
    import ctypes
    from hpgeom_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(4)
    D, nside = 50, 256
    a, b = 360 * rng.random(D), np.degrees(np.arcsin(2 * rng.random(D) - 1))
    r = 10 ** rng.uniform(-1, 1.5, D)
    r[7] = 200.0  # degrees: covers the sphere
    plans = (_lib.Disc * D)()
    bad = ctypes.c_int64(-1)
    assert L.hpgb_query_circle_ring_plan_many(nside, a.ctypes.data, b.ctypes.data, r.ctypes.data, D, 1, 2, 1, 1, plans,
                                              ctypes.byref(bad)) == 0 and bad.value == -1
    one = _lib.Disc()
    for d in (0, 7, 31):
        assert L.hpgb_query_circle_ring_plan(nside, a[d], b[d], r[d], 1, 2, 1, 1, ctypes.byref(one)) == 0
        assert bytes(one) == bytes(plans[d])
    assert plans[7].whole_sphere == 1 and plans[7].m == 1
    ring_start = np.zeros(D + 1, dtype=np.int64)
    total = L.hpgb_query_circle_ring_batch_prepare(plans, D, ring_start.ctypes.data)
    counts = [max(p.irmax - p.irmin + 1, 0) for p in plans]
    assert counts[7] == 0 and plans[7].north_hi == 12 * nside * nside and plans[7].south_lo == -1
    assert ring_start.tolist() == [0] + np.cumsum(counts).tolist() and total == sum(counts)
    # the first offending disc and the reference's error class for it
    b2 = b.copy()
    b2[13] = 91.0
    r2 = r.copy()
    r2[5] = -1.0
    assert L.hpgb_query_circle_ring_plan_many(nside, a.ctypes.data, b2.ctypes.data, r2.ctypes.data, D, 0, 4, 1, 1, plans,
                                              ctypes.byref(bad)) == 7 and bad.value == 5   # radius
    r2[5] = 1.0
    assert L.hpgb_query_circle_ring_plan_many(nside, a.ctypes.data, b2.ctypes.data, r2.ctypes.data, D, 0, 4, 1, 1, plans,
                                              ctypes.byref(bad)) == 6 and bad.value == 13  # latitude


# ------------------------------------------------------------------------------------------------
# GPU tier
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def hpg():
    import hpgeom_b200
    return hpgeom_b200


@pytest.mark.gpu
def test_query_circle_golden(hpg):
    n = 0
    for (nside, a, b, radius, fact, lonlat, degrees), want in cases():
        got = hpg.query_circle(nside, a, b, radius, inclusive=fact > 0, fact=max(fact, 1), nest=False, lonlat=lonlat,
                               degrees=degrees)
        assert_bit_equal(got, want, "nside=%d fact=%d lonlat=%d" % (nside, fact, lonlat))
        n += 1
    assert n > 900


@pytest.mark.gpu
def test_query_circle_device_resident_vec_and_healpy(hpg):
    import torch
    from hpgeom_b200 import healpy_compat as hpc
    k = 0
    for (nside, a, b, radius, fact, lonlat, degrees), want in cases():
        k += 1
        if k % 25:
            continue
        t = hpg.query_circle(nside, a, b, radius, inclusive=fact > 0, fact=max(fact, 1), nest=False, lonlat=lonlat,
                             degrees=degrees, device_resident=True)
        assert isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.int64
        assert_bit_equal(t.cpu().numpy(), want)
    vec = hpg.angle_to_vector(1.1, 2.2, lonlat=False)
    a = hpg.query_circle_vec(512, vec, 0.05, nest=False)
    b = hpg.query_circle(512, 1.1, 2.2, 0.05, nest=False, lonlat=False)
    c = hpc.query_disc(512, vec, 0.05)
    assert a.size > 100 and np.array_equal(a, b) and np.array_equal(a, c)


@pytest.mark.gpu
def test_query_circle_nest_golden(hpg):
    import torch
    n = 0
    for (nside, theta, phi, radius, fact), pix, ranges in nest_cases():
        kw = dict(inclusive=fact > 0, fact=max(fact, 1), nest=True, lonlat=False)
        assert_bit_equal(hpg.query_circle(nside, theta, phi, radius, **kw), pix, "nside=%d fact=%d" % (nside, fact))
        n += 1
        if n % 5 == 0:
            assert_bit_equal(hpg.query_circle(nside, theta, phi, radius, return_pixel_ranges=True, **kw), ranges)
        if n % 40 == 0:
            t = hpg.query_circle(nside, theta, phi, radius, device_resident=True, **kw)
            assert isinstance(t, torch.Tensor) and t.is_cuda
            assert_bit_equal(t.cpu().numpy(), pix)
            assert_bit_equal(hpg.query_circle(np.int64(nside), np.degrees(phi), 90.0 - np.degrees(theta), np.degrees(radius),
                                              inclusive=fact > 0, fact=max(fact, 1)).shape, pix.shape)
    assert n > 400


@pytest.mark.gpu
def test_query_circle_nest_big_and_whole_sphere(hpg, oracle_ref):
    """2e8 pixels at nside 2^14 (same set as the RING query, up to pixel-centre round-off), default
    arguments = NEST, a whole-sphere disc, and the live reference where present."""
    nside, theta, phi, radius = 2 ** 14, 0.7, 4.0, 0.5
    got = hpg.query_circle(nside, theta, phi, radius, lonlat=False)
    assert got.size > 1.9e8 and np.all(np.diff(got) > 0)
    ring = np.sort(hpg.ring_to_nest(nside, hpg.query_circle(nside, theta, phi, radius, nest=False, lonlat=False)))
    assert abs(ring.size - got.size) < 50 and np.setxor1d(ring, got, assume_unique=True).size < 50
    if oracle_ref is not None:
        assert_bit_equal(got, oracle_ref.query_circle(nside, theta, phi, radius, lonlat=False))
        for args in ((2 ** 20, 33.0, -41.0, 0.02), (2 ** 12, 200.0, 89.9, 3.0), (64, 0.0, 0.0, 200.0), (2 ** 25, 10.0, 10.0, 1e-4)):
            for kw in (dict(), dict(inclusive=True), dict(inclusive=True, fact=16), dict(return_pixel_ranges=True)):
                assert_bit_equal(hpg.query_circle(*args, **kw), oracle_ref.query_circle(*args, **kw), "%r %r" % (args, kw))
    assert hpg.query_circle(8, 0.0, 0.0, 3.2, lonlat=False).tolist() == list(range(768))
    assert hpg.query_circle(8, 0.0, 0.0, 3.2, lonlat=False, return_pixel_ranges=True).tolist() == [[0, 768]]


@pytest.mark.gpu
def test_query_circle_big_disc_properties(hpg, oracle_ref):
    """A disc of 2.1e8 pixels at nside 2^14: sorted, unique, complete against pixel_to_angle distances;
    identical to the reference where it is present."""
    nside, theta, phi, radius = 2 ** 14, 0.7, 4.0, 0.5
    got = hpg.query_circle(nside, theta, phi, radius, nest=False, lonlat=False)
    assert got.size > 1.9e8 and np.all(np.diff(got) > 0)
    if oracle_ref is not None:
        assert_bit_equal(got, oracle_ref.query_circle(nside, theta, phi, radius, nest=False, lonlat=False))
    rng = np.random.default_rng(3)
    samp = rng.integers(0, 12 * nside * nside, 2_000_000)
    th, ph = hpg.pixel_to_angle(nside, samp, nest=False, lonlat=False)
    cosd = np.cos(th) * np.cos(theta) + np.sin(th) * np.sin(theta) * np.cos(ph - phi)
    inside = cosd > np.cos(radius) + 1e-9
    outside = cosd < np.cos(radius) - 1e-9
    member = np.isin(samp, got)
    assert np.all(member[inside]) and not np.any(member[outside])


@pytest.mark.gpu
def test_pixel_ranges_to_pixels_golden_and_torch(hpg):
    import torch
    g = load_golden("query")
    rr = g["pr_ranges"]
    assert hpg.pixel_ranges_to_pixels(np.array([[0, 3], [10, 10], [20, 22]])).tolist() == [0, 1, 2, 20, 21]
    assert hpg.pixel_ranges_to_pixels(np.array([[0, 3], [10, 10]]), inclusive=True).tolist() == [0, 1, 2, 3, 10]
    with pytest.raises(ValueError, match=r"must all be <="):
        hpg.pixel_ranges_to_pixels(np.array([[5, 3]]))
    assert_bit_equal(hpg.pixel_ranges_to_pixels(rr), g["pr_excl"])
    assert_bit_equal(hpg.pixel_ranges_to_pixels(rr, inclusive=True), g["pr_incl"])
    r32 = (rr % 1000 + np.array([0, 1000])).astype(np.int32)  # safe cast int32 -> int64, like the reference
    assert_bit_equal(hpg.pixel_ranges_to_pixels(r32), np.concatenate([np.arange(a, b, dtype=np.int64) for a, b in r32]))
    t = hpg.pixel_ranges_to_pixels(torch.from_numpy(rr).cuda(), inclusive=True)
    assert t.is_cuda
    assert_bit_equal(t.cpu().numpy(), g["pr_incl"])
    bad = rr.copy()
    bad[1234, 1] = bad[1234, 0] - 1
    for x in (bad, torch.from_numpy(bad).cuda()):
        with pytest.raises(ValueError, match=r"pixel_ranges\[:, 0\] must all be <= pixel_ranges\[:, 1\]"):
            hpg.pixel_ranges_to_pixels(x)
    with pytest.raises(TypeError):
        hpg.pixel_ranges_to_pixels(rr.astype(np.float64))


@pytest.mark.gpu
@pytest.mark.parametrize("shape", ["long", "short", "mixed", "empties", "single"])
def test_pixel_ranges_to_pixels_shapes(hpg, oracle_port, shape):
    """Every regime of the expand kernel: one range per many tiles, many ranges per tile, more than
    one 1024-offset window per tile (runs of empty rows), and a scan over several scan tiles."""
    rng = np.random.default_rng(11)
    if shape == "long":
        m, ln = 37, rng.integers(1, 3_000_000, 37)
    elif shape == "short":
        m, ln = 700_001, rng.integers(0, 7, 700_001)
    elif shape == "mixed":
        m = 50_000
        ln = np.where(rng.random(m) < 0.01, rng.integers(1000, 200_000, m), rng.integers(0, 4, m))
    elif shape == "empties":
        m = 300_000
        ln = np.zeros(m, dtype=np.int64)
        ln[rng.integers(0, m, 40)] = rng.integers(1, 5000, 40)
        ln[0] = ln[-1] = 0
    else:
        m, ln = 1, np.array([12345])
    lo = rng.integers(0, 2 ** 58, m)
    rr = np.stack([lo, lo + ln], axis=1).astype(np.int64)
    for incl in (False, True):
        want = oracle_port.pixel_ranges_to_pixels(rr, inclusive=incl)
        assert_bit_equal(hpg.pixel_ranges_to_pixels(rr, inclusive=incl), want, "%s incl=%d" % (shape, incl))
    import torch
    t = torch.from_numpy(rr).cuda()
    assert_bit_equal(hpg.pixel_ranges_to_pixels(t).cpu().numpy(), oracle_port.pixel_ranges_to_pixels(rr))
    # a misaligned output window (odd out_begin) through the raw C ABI
    from hpgeom_b200 import _lib
    import ctypes
    L = _lib.lib()
    offs = torch.empty(m + 1, dtype=torch.int64, device="cuda")
    st = torch.full((2,), -1, dtype=torch.int64, device="cuda")
    assert L.hpgb_ranges_offsets(t.data_ptr(), m, 0, offs.data_ptr(), st.data_ptr(), None) == 0
    total = int(offs[m].item())
    want = oracle_port.pixel_ranges_to_pixels(rr)
    assert total == want.size
    if total > 10:
        begin, count = 3, total - 7
        buf = torch.zeros(count + 1, dtype=torch.int64, device="cuda")
        assert L.hpgb_ranges_expand(t.data_ptr(), offs.data_ptr(), m, begin, count, buf.data_ptr() + 8, None) == 0
        torch.cuda.synchronize()
        assert_bit_equal(buf[1:].cpu().numpy(), want[begin:begin + count])


@pytest.mark.gpu
def test_query_circle_batch_matches_single_goldens(hpg):
    """Batched RING query: every list of a batch equals the golden single-disc result (batches = the golden
    cases grouped by nside and fact), host and device-resident forms; an empty batch; error reporting."""
    import collections
    import torch
    groups = collections.defaultdict(list)
    for (nside, a, b, radius, fact, lonlat, degrees), want in cases():
        if not lonlat:
            groups[(nside, fact)].append((a, b, radius, want))
    assert len(groups) > 30
    for k, ((nside, fact), items) in enumerate(sorted(groups.items())):
        a = np.array([it[0] for it in items]); b = np.array([it[1] for it in items]); r = np.array([it[2] for it in items])
        kw = dict(inclusive=fact > 0, fact=max(fact, 1), lonlat=False)
        pix, offs = hpg.query_circle_batch(nside, a, b, r, **kw)
        assert offs.shape == (len(items) + 1,) and offs[0] == 0 and offs[-1] == pix.size
        for d, it in enumerate(items):
            assert_bit_equal(pix[offs[d]:offs[d + 1]], it[3], "nside=%d fact=%d disc %d" % (nside, fact, d))
        if k % 4 == 0:
            tp, to = hpg.query_circle_batch(nside, a, b, r, device_resident=True, **kw)
            assert isinstance(tp, torch.Tensor) and tp.is_cuda
            assert_bit_equal(tp.cpu().numpy(), pix)
            assert_bit_equal(to.cpu().numpy(), offs)
    pix, offs = hpg.query_circle_batch(64, np.zeros(0), np.zeros(0), 0.1)
    assert pix.shape == (0,) and offs.tolist() == [0]
    with pytest.raises(ValueError, match=r"lat = 95 out of range \[-90, 90\]"):
        hpg.query_circle_batch(1024, [10.0, 20.0], [5.0, 95.0], 1.0)
    # scalar radius broadcast, many small discs, against the single call
    rng = np.random.default_rng(8)
    lon, lat = 360 * rng.random(2000), np.degrees(np.arcsin(2 * rng.random(2000) - 1))
    pix, offs = hpg.query_circle_batch(2048, lon, lat, 0.2)
    for d in (0, 17, 1999):
        assert_bit_equal(pix[offs[d]:offs[d + 1]], hpg.query_circle(2048, lon[d], lat[d], 0.2, nest=False))


@pytest.mark.gpu
def test_query_circle_batch_nest_matches_single_goldens(hpg):
    """Batched NEST search: every list of a batch equals the golden single-disc result."""
    import collections
    import torch
    groups = collections.defaultdict(list)
    for (nside, theta, phi, radius, fact), pix, ranges in nest_cases():
        groups[(nside, fact)].append((theta, phi, radius, pix))
    assert len(groups) > 30
    for k, ((nside, fact), items) in enumerate(sorted(groups.items())):
        a = np.array([it[0] for it in items]); b = np.array([it[1] for it in items]); r = np.array([it[2] for it in items])
        kw = dict(inclusive=fact > 0, fact=max(fact, 1), nest=True, lonlat=False)
        pix, offs = hpg.query_circle_batch(nside, a, b, r, **kw)
        assert offs.shape == (len(items) + 1,) and offs[0] == 0 and offs[-1] == pix.size
        for d, it in enumerate(items):
            assert_bit_equal(pix[offs[d]:offs[d + 1]], it[3], "nside=%d fact=%d disc %d" % (nside, fact, d))
        if k % 5 == 0:
            tp, to = hpg.query_circle_batch(nside, a, b, r, device_resident=True, **kw)
            assert isinstance(tp, torch.Tensor) and tp.is_cuda
            assert_bit_equal(tp.cpu().numpy(), pix)
            assert_bit_equal(to.cpu().numpy(), offs)
    pix, offs = hpg.query_circle_batch(64, np.zeros(0), np.zeros(0), 0.1, nest=True)
    assert pix.shape == (0,) and offs.tolist() == [0]
    with pytest.raises(ValueError, match=r"power of 2"):
        hpg.query_circle_batch(1000, [10.0], [5.0], 1.0, nest=True)
    with pytest.raises(ValueError, match=r"Radius must be positive."):
        hpg.query_circle_batch(1024, [10.0, 20.0], [5.0, 5.0], [1.0, 0.0], nest=True)
    # whole-sphere discs mixed with small ones, default-argument single calls as the check
    lon, lat, rad = np.array([10.0, 200.0, 33.0, 300.0]), np.array([5.0, -60.0, 89.0, 0.0]), np.array([1.0, 190.0, 0.5, 181.0])
    pix, offs = hpg.query_circle_batch(32, lon, lat, rad, nest=True)
    for d in range(4):
        assert_bit_equal(pix[offs[d]:offs[d + 1]], hpg.query_circle(32, lon[d], lat[d], rad[d]))
