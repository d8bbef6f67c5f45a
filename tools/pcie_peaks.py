"""Host<->device copy ceilings of this box with pinned memory, on 1 .. N GPUs driven from ONE process, next to the
single-process multi-GPU e2e rate of the drop-in (hpgb_host_angle_to_pixel with HPGB_DEVICE_ALL).

usage: python tools/pcie_peaks.py [--gpus N] [--log2 27]
For every G in (1, 2, 4, ... N): concurrent H2D (16 B/point) + D2H (8 B/point) on G GPUs at the path's 2:1 ratio
-> the e2e ceiling in Gpoints/s; then the measured e2e rate with pinned and with pageable host arrays.  If the
ceiling stops scaling with G the host side (memory bandwidth / PCIe root complex) is the limiter, not the pipeline."""
import argparse
import ctypes
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hpgeom_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=torch.cuda.device_count())
ap.add_argument("--log2", type=int, default=27, help="points per GPU")
args = ap.parse_args()
L = _lib.lib()
n = 1 << args.log2
N = min(args.gpus, torch.cuda.device_count())
print("host: %d cpus; GPUs visible: %d; points per GPU 2^%d" % (os.cpu_count(), torch.cuda.device_count(), args.log2))
try:
    import subprocess
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
except Exception:
    pass

h_in = [torch.empty(2 * n, dtype=torch.float64).pin_memory() for _ in range(N)]
h_out = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(N)]
d_in = [torch.empty(2 * n, dtype=torch.float64, device="cuda:%d" % g) for g in range(N)]
d_out = [torch.empty(n, dtype=torch.float64, device="cuda:%d" % g) for g in range(N)]
s_in = [torch.cuda.Stream(device=g) for g in range(N)]
s_out = [torch.cuda.Stream(device=g) for g in range(N)]


def sync(G):
    for g in range(G):
        torch.cuda.synchronize(g)


def raw(G, h2d, d2h, reps=4):
    sync(G)
    t0 = time.perf_counter()
    for _ in range(reps):
        for g in range(G):
            if h2d:
                with torch.cuda.stream(s_in[g]):
                    d_in[g].copy_(h_in[g], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s_out[g]):
                    h_out[g].copy_(d_out[g], non_blocking=True)
    sync(G)
    return (time.perf_counter() - t0) / reps


Gs = [g for g in (1, 2, 4, 8) if g <= N]
raw(N, True, True, 1)
print("%-4s %-12s %-12s %-34s" % ("GPUs", "H2D alone", "D2H alone", "H2D+D2H at 2:1 (total) -> ceiling"))
ceil = {}
for G in Gs:
    t1, t2, t3 = raw(G, True, False), raw(G, False, True), raw(G, True, True)
    ceil[G] = G * n / t3 / 1e9
    print("%-4d %7.1f GB/s %7.1f GB/s   H2D %6.1f + D2H %6.1f GB/s -> %6.2f Gpoints/s" % (
        G, G * 16 * n / t1 / 1e9, G * 8 * n / t2 / 1e9, G * 16 * n / t3 / 1e9, G * 8 * n / t3 / 1e9, ceil[G]))

# e2e through the C ABI, one process, G GPUs: pinned and pageable host arrays
nside = 2 ** 29
fb = ctypes.c_int64(-1)
for G in Gs:
    m = G * n
    devs = (ctypes.c_int * G)(*range(G))
    assert L.hpgb_host_set_devices(devs, G) == 0
    for kind in ("pinned", "pageable"):
        if kind == "pinned":
            lon = torch.empty(m, dtype=torch.float64).pin_memory()
            lat = torch.empty(m, dtype=torch.float64).pin_memory()
            pix = torch.empty(m, dtype=torch.int64).pin_memory()
            lon.uniform_(0, 360)
            lat.uniform_(-89, 89)
            pl, pb, pp = lon.data_ptr(), lat.data_ptr(), pix.data_ptr()
        else:
            a = np.random.default_rng(1).uniform(0, 360, m)
            b = np.random.default_rng(2).uniform(-89, 89, m)
            p = np.zeros(m, dtype=np.int64)
            pl, pb, pp = a.ctypes.data, b.ctypes.data, p.ctypes.data
        best = 1e9
        for it in range(4):
            t0 = time.perf_counter()
            rc = L.hpgb_host_angle_to_pixel(pl, pb, pp, m, nside, None, 1, 1, 1, -1 if G > 1 else 0, ctypes.byref(fb))
            dt = time.perf_counter() - t0
            assert rc == 0 and fb.value == -1, (rc, fb.value)
            if it:
                best = min(best, dt)
        print("e2e one process, %d GPU(s), %-8s host arrays: %6.2f Gpoints/s  (%.2f of the copy ceiling %.2f)" % (
            G, kind, m / best / 1e9, m / best / 1e9 / ceil[G], ceil[G]))
L.hpgb_host_set_devices(None, 0)
