#!/usr/bin/env python
"""CPU baseline of every function on the path: the reference's own C (oracle/_ref) on this host,
n_threads = 1 and all cores, BASELINE.json's configs (bounded samples).  Prints a markdown table."""
import os
import subprocess
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle

ref = oracle.ref
assert ref is not None, "oracle/_ref not built"
T = os.cpu_count() or 1
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 24
rng = np.random.default_rng(12345)
lon = 360.0 * rng.random(n)
lat = np.degrees(np.arcsin(2.0 * rng.random(n) - 1.0))


def rate(fn, m=n, reps=3):
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return m / best / 1e6


def both(name, f):
    r1, rt = rate(lambda: f(1)), rate(lambda: f(T))
    print("| %s | %.1f | %.1f |" % (name, r1, rt))


try:
    model = subprocess.check_output("lscpu | grep 'Model name' | head -1", shell=True, text=True).split(":")[1].strip()
except Exception:
    model = "?"
print("host: %s, %d logical CPUs; reference C (oracle/_ref), %d elements per call, best of 3, Melem/s" % (model, T, n))
print("| function / config | 1 thread | %d threads |\n|---|---|---|" % T)
for ns, tag in ((1024, "1024 (config 1)"), (2 ** 29, "2^29 (config 5)")):
    both("angle_to_pixel nside=%s nest lonlat deg" % tag, lambda t: ref.angle_to_pixel(ns, lon, lat, nest=True, lonlat=True, degrees=True, n_threads=t))
both("angle_to_pixel nside=2^29 ring", lambda t: ref.angle_to_pixel(2 ** 29, lon, lat, nest=False, lonlat=True, degrees=True, n_threads=t))
ns = 2 ** 29
pix = rng.integers(0, 12 * ns * ns, n)
both("nest_to_ring nside=2^29 (config 3)", lambda t: ref.nest_to_ring(ns, pix, n_threads=t))
both("ring_to_nest nside=2^29 (config 3)", lambda t: ref.ring_to_nest(ns, pix, n_threads=t))
ns = 2 ** 20
pix = rng.integers(0, 12 * ns * ns, n)
for nest in (True, False):
    s = "nest" if nest else "ring"
    both("pixel_to_angle nside=2^20 %s (config 2)" % s, lambda t: ref.pixel_to_angle(ns, pix, nest=nest, lonlat=True, degrees=True, n_threads=t))
    both("pixel_to_vector nside=2^20 %s (config 2)" % s, lambda t: ref.pixel_to_vector(ns, pix, nest=nest, n_threads=t))
x, y, z = ref.pixel_to_vector(ns, pix, nest=True, n_threads=T)
both("vector_to_pixel nside=2^29 nest", lambda t: ref.vector_to_pixel(2 ** 29, x, y, z, nest=True, n_threads=t))
ns = 4096
m = n // 4
pix = rng.integers(0, 12 * ns * ns - 1, m)
for nest in (True, False):
    s = "nest" if nest else "ring"
    r1, rt = rate(lambda: ref.neighbors(ns, pix, nest=nest, n_threads=1), m), rate(lambda: ref.neighbors(ns, pix, nest=nest, n_threads=T), m)
    print("| neighbors nside=4096 %s (config 4) | %.1f | %.1f |" % (s, r1, rt))
    r1 = rate(lambda: ref.get_interpolation_weights(ns, lon[:m], lat[:m], nest=nest, lonlat=True, degrees=True, n_threads=1), m)
    rt = rate(lambda: ref.get_interpolation_weights(ns, lon[:m], lat[:m], nest=nest, lonlat=True, degrees=True, n_threads=T), m)
    print("| get_interpolation_weights nside=4096 %s (config 4) | %.1f | %.1f |" % (s, r1, rt))
