"""query_circle / pixel_ranges_to_pixels end to end (numpy out) against the reference's CPU path."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpgeom_b200 as hpg
import oracle
ref = oracle.ref


def best(fn, k=3):
    fn()
    b = 1e9
    for _ in range(k):
        t0 = time.perf_counter(); r = fn(); b = min(b, time.perf_counter() - t0)
    return b, r


for nside, radius, kw in ((2 ** 14, 0.5, dict(nest=False)), (2 ** 14, 0.5, dict(nest=True)),
                          (2 ** 12, 0.3, dict(nest=False, inclusive=True, fact=4)), (2 ** 12, 0.3, dict(nest=True, inclusive=True, fact=4)),
                          (2 ** 20, 0.003, dict(nest=True)), (2 ** 10, 0.05, dict(nest=True)), (2 ** 10, 0.05, dict(nest=False))):
    args = (nside, 0.7, 4.0, radius)
    t, got = best(lambda: hpg.query_circle(*args, lonlat=False, **kw))
    td, gd = best(lambda: hpg.query_circle(*args, lonlat=False, device_resident=True, **kw))
    line = "nside=2^%-2d r=%-6g %-44s %11d pix  ours %8.2f ms (device-resident %7.2f ms)" % (
        int(np.log2(nside)), radius, str(kw), got.size, t * 1e3, td * 1e3)
    if ref is not None:
        tr, want = best(lambda: ref.query_circle(*args, lonlat=False, **kw), k=2)
        assert np.array_equal(got, want)
        line += "  reference %8.2f ms  (x%.1f / x%.1f)" % (tr * 1e3, tr / t, tr / td)
    print(line, flush=True)

rng = np.random.default_rng(0)
for name, lens in (("1e6-pixel ranges", np.full(200, 1_000_000)), ("ragged 0..63", rng.integers(0, 64, 4_000_000))):
    lo = np.cumsum(lens + 3) - lens
    rr = np.stack([lo, lo + lens], 1).astype(np.int64)
    t, got = best(lambda: hpg.pixel_ranges_to_pixels(rr))
    line = "pixel_ranges_to_pixels %-18s %11d pix  ours %8.2f ms" % (name, got.size, t * 1e3)
    if ref is not None:
        tr, want = best(lambda: ref.pixel_ranges_to_pixels(rr), k=2)
        assert np.array_equal(got, want)
        line += "  reference %8.2f ms  (x%.1f)" % (tr * 1e3, tr / t)
    print(line, flush=True)

# many small discs: the batched RING query against a loop over the reference
rng = np.random.default_rng(3)
for D, nside, rdeg in ((20000, 2048, 0.2), (200000, 2048, 0.2), (20000, 2 ** 16, 0.005)):
    lon, lat = 360 * rng.random(D), np.degrees(np.arcsin(2 * rng.random(D) - 1))
    t, (pix, offs) = best(lambda: hpg.query_circle_batch(nside, lon, lat, rdeg))
    line = "batch of %6d discs nside=%d r=%g deg: %10d pix  ours %8.2f ms (%.2f Mdiscs/s)" % (D, nside, rdeg, pix.size, t * 1e3, D / t / 1e6)
    if ref is not None:
        nref = min(D, 20000)
        t0 = time.perf_counter()
        for d in range(nref):
            r = ref.query_circle(nside, lon[d], lat[d], rdeg, nest=False)
        tr = (time.perf_counter() - t0) * D / nref
        assert np.array_equal(r, pix[offs[nref - 1]:offs[nref]])
        line += "  reference loop %9.2f ms (x%.1f)" % (tr * 1e3, tr / t)
    print(line, flush=True)

for D, nside, rdeg in ((20000, 2048, 0.2), (200000, 2048, 0.2)):
    lon, lat = 360 * rng.random(D), np.degrees(np.arcsin(2 * rng.random(D) - 1))
    t, (pix, offs) = best(lambda: hpg.query_circle_batch(nside, lon, lat, rdeg, nest=True))
    line = "NEST batch of %6d discs nside=%d r=%g deg: %10d pix  ours %8.2f ms (%.2f Mdiscs/s)" % (D, nside, rdeg, pix.size, t * 1e3, D / t / 1e6)
    if ref is not None:
        nref = min(D, 20000)
        t0 = time.perf_counter()
        for d in range(nref):
            r = ref.query_circle(nside, lon[d], lat[d], rdeg)
        tr = (time.perf_counter() - t0) * D / nref
        assert np.array_equal(r, pix[offs[nref - 1]:offs[nref]])
        line += "  reference loop %9.2f ms (x%.1f)" % (tr * 1e3, tr / t)
    print(line, flush=True)
