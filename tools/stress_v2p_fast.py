"""Host stress test: vector_to_pixel fast path + fallback must equal the exact path on every vector."""
import sys
import numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from hostsim import HostSim
sim = HostSim()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
rng = np.random.default_rng(77)
tot = 0
for nside in [2 ** 29, 2 ** 28, 2 ** 20, 2 ** 16, 2 ** 15, 1024, 8, 1]:
    lon = rng.uniform(0, 2 * np.pi, n)
    zz = rng.uniform(-1, 1, n)
    k = n // 8
    zz[:k] = np.sign(rng.uniform(-1, 1, k)) * (1 - 10 ** rng.uniform(-16, -1, k))                   # poles
    zz[k:2 * k] = np.sign(rng.uniform(-1, 1, k)) * (2 / 3 + rng.normal(0, 1, k) * 10 ** rng.uniform(-14, -2, k))  # transition
    lon[2 * k:3 * k] = (np.pi / 4) * rng.integers(0, 9, k) + rng.normal(0, 1, k) * 10 ** rng.uniform(-14, -2, k)  # face edges, octant seams
    zz[3 * k:4 * k] = rng.normal(0, 1, k) * 10 ** rng.uniform(-14, -1, k)                           # equator
    zz = np.clip(zz, -1, 1)
    st = np.sqrt((1 - zz) * (1 + zz))
    scale = 10.0 ** rng.uniform(-100, 100, n)                                                    # un-normalised vectors
    scale[: n // 2] = 1.0
    x, y, z = st * np.cos(lon) * scale, st * np.sin(lon) * scale, zz * scale
    x[4 * k:4 * k + 100] = 0.0; y[4 * k + 50:4 * k + 150] = 0.0
    for nest in (True, False):
        fast = sim.vector_to_pixel(nside, x, y, z, nest=nest)
        slow_n = sim.last_slow
        exact = sim.vector_to_pixel(nside, x, y, z, nest=nest, exact_only=True)
        bad = np.flatnonzero(fast != exact)
        tot += bad.size
        print("nside 2^%-4.1f %s: %d vectors, exact-path fraction %.2e, fast != exact: %d %s" % (
            np.log2(nside), "nest" if nest else "ring", n, slow_n / n, bad.size, (x[bad[:3]], y[bad[:3]], z[bad[:3]]) if bad.size else ""))
print("TOTAL mismatches:", tot)
