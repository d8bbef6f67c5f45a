"""Host stress test: angle_to_pixel fast path + fallback must equal the exact path on every point
(both use our own correctly rounded trig, so the comparison is exact, no libm noise)."""
import sys
import numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from hostsim import HostSim
sim = HostSim()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000
rng = np.random.default_rng(2024)
tot_bad = 0
for nside in [2 ** 29, 2 ** 28, 2 ** 20, 2 ** 16, 2 ** 15, 2 ** 10, 8, 1]:
    lon = rng.uniform(0, 360, n)
    lat = np.degrees(np.arcsin(rng.uniform(-1, 1, n)))
    k = n // 8
    # concentrate: poles, the +-41.81 deg transition, face edges in longitude, equator
    lat[:k] = np.sign(rng.uniform(-1, 1, k)) * (90 - 10 ** rng.uniform(-9, 1, k))
    lat[k:2 * k] = np.sign(rng.uniform(-1, 1, k)) * (41.810314895778596 + rng.normal(0, 1, k) * 10 ** rng.uniform(-12, 0, k))
    lon[2 * k:3 * k] = 45.0 * rng.integers(0, 9, k) + rng.normal(0, 1, k) * 10 ** rng.uniform(-12, -1, k)
    lat[3 * k:4 * k] = rng.normal(0, 1, k) * 10 ** rng.uniform(-12, 0, k)
    lon[4 * k:5 * k] = rng.uniform(-720, 720, k)
    lat = np.clip(lat, -90, 90)
    for kw, (a, b) in {"deg": (lon, lat), "rad": (np.radians(lon), np.radians(lat)),
                       "tp": (np.clip(np.pi / 2 - np.radians(lat), 0, np.pi), np.radians(lon % 360))}.items():
        flags = dict(lonlat=kw != "tp", degrees=kw == "deg")
        for nest in (True, False):
            fast = sim.angle_to_pixel(nside, a, b, nest=nest, **flags)
            slow_n = sim.last_slow
            exact = sim.angle_to_pixel(nside, a, b, nest=nest, exact_only=True, **flags)
            bad = np.flatnonzero(fast != exact)
            tot_bad += bad.size
            print("nside 2^%-4.1f %-3s %s: %d points, exact-path fraction %.2e, fast != exact: %d %s" % (
                np.log2(nside), kw, "nest" if nest else "ring", n, slow_n / n, bad.size, (a[bad[:3]], b[bad[:3]]) if bad.size else ""))
print("TOTAL mismatches:", tot_bad)
