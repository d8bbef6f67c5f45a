"""End-to-end rate of the Python drop-in with plain (pageable) numpy arrays vs the reference's CPU path."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpgeom_b200 as hpg
import oracle
for n in (10_000_000, 100_000_000):
    rng = np.random.default_rng(12345)
    lon = 360.0 * rng.random(n)
    lat = np.degrees(np.arcsin(2.0 * rng.random(n) - 1.0))
    hpg.angle_to_pixel(1024, lon[:1000], lat[:1000])
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter(); p = hpg.angle_to_pixel(1024, lon, lat); best = min(best, time.perf_counter() - t0)
    line = "n=%9d  hpgeom_b200 (numpy in/out): %7.1f ms  %.2f Gpts/s" % (n, best * 1e3, n / best / 1e9)
    if oracle.ref is not None:
        for th in (1, os.cpu_count()):
            b2 = 1e9
            for _ in range(3):
                t0 = time.perf_counter(); r = oracle.ref.angle_to_pixel(1024, lon, lat, n_threads=th); b2 = min(b2, time.perf_counter() - t0)
            line += "   reference %d threads: %7.1f ms" % (th, b2 * 1e3)
        assert np.array_equal(p, r)
    print(line)
