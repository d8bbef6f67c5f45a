"""Host stress: interpolation fast path (cheap ring choice, table-driven colatitudes, tame division)
must equal the reference-shaped path bit for bit."""
import sys
import numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from hostsim import HostSim
sim = HostSim()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 400_000
rng = np.random.default_rng(11)
tot = 0
for nside in [2 ** 29, 2 ** 20, 4096, 1024, 64, 8, 2, 1]:
    lon = rng.uniform(0, 360, n)
    lat = np.degrees(np.arcsin(rng.uniform(-1, 1, n)))
    k = n // 8
    lat[:k] = np.sign(rng.uniform(-1, 1, k)) * (90 - 10 ** rng.uniform(-9, 1, k))
    lat[k:2 * k] = np.sign(rng.uniform(-1, 1, k)) * (41.810314895778596 + rng.normal(0, 1, k) * 10 ** rng.uniform(-12, 0, k))
    # ring colatitudes themselves (points sitting on a ring: the ring_above step)
    pix = rng.integers(0, 12 * nside * nside, k)
    import oracle
    plon, plat = oracle.port().pixel_to_angle(nside, pix, nest=False)
    lat[2 * k:3 * k] = plat * (1 + rng.integers(-3, 4, k) * 2.0 ** -52)
    lat[3 * k:4 * k] = rng.normal(0, 1, k) * 10 ** rng.uniform(-12, 0, k)
    lat = np.clip(lat, -90, 90)
    for nest in (True, False):
        for name, (a, b, kw) in {"deg": (lon, lat, dict(lonlat=True, degrees=True)),
                                 "tp": (np.clip(np.pi / 2 - np.radians(lat), 0, np.pi), np.radians(lon), dict(lonlat=False, degrees=False))}.items():
            p1, w1 = sim.get_interpolation_weights_k(nside, a, b, nest=nest, **kw)
            slow = sim.last_slow
            p0, w0 = sim.get_interpolation_weights(nside, a, b, nest=nest, **kw)
            bad = np.flatnonzero((p1 != p0).any(axis=1) | (w1.view(np.int64) != w0.view(np.int64)).any(axis=1))
            tot += bad.size
            print("nside 2^%-4.1f %s %-3s: fallback fraction %.2e, differing rows %d %s" % (
                np.log2(nside), "nest" if nest else "ring", name, slow / n, bad.size, (a[bad[:2]], b[bad[:2]]) if bad.size else ""))
print("TOTAL differing rows:", tot)
