"""A/B timing of kernel variants IN ONE PROCESS: several builds of libhpgb.so are loaded side by side and the same
kernel is launched from each in turn (round robin), so that clock / power drift of the box hits every variant alike.
usage: python tools/ab_multi.py <log2n> <kernel> <reps> <rounds> <lib> [<lib> ...]     (lib = path, or 'default')
kernels: a2p a2p_ring n2r r2n p2a p2a_ring p2v v2p interp interp_ring neigh neigh_ring upg upg_ring"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, ".")
from hpgeom_b200 import _lib

log2, kern, reps, rounds = int(sys.argv[1]), sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
paths = [(_lib.LIB_PATH if p == "default" else p) for p in sys.argv[5:]]
libs = []
for p in paths:
    L = ctypes.CDLL(os.path.abspath(p))
    for name, args in _lib._SIGS.items():
        if hasattr(L, name):
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = _lib._RESTYPES.get(name, ctypes.c_int)
    libs.append(L)
n = 1 << log2
nside = int(os.environ.get("HPGB_AB_NSIDE", 2 ** 29))
g = torch.Generator(device="cuda").manual_seed(1)
P = lambda t: ctypes.c_void_p(t.data_ptr())
st = torch.empty(2, dtype=torch.int64, device="cuda")
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
outs = []
if kern.startswith("a2p"):
    lon = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 360.0
    lat = torch.rad2deg(torch.asin(torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1))
    out = torch.empty(n, dtype=torch.int64, device="cuda")
    nest = 0 if kern.endswith("ring") else 1
    call = lambda L: L.hpgb_angle_to_pixel(P(lon), P(lat), P(out), n, nside, None, nest, 1, 1, P(st), s)
    outs, bpe, m = [out], 24, n
elif kern in ("n2r", "r2n", "p2a", "p2a_ring", "p2v", "v2p"):
    pix = torch.randint(0, 12 * nside * nside, (n,), dtype=torch.int64, device="cuda", generator=g)
    out = torch.empty(n, dtype=torch.int64, device="cuda")
    a, b, c = (torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(3))
    m = n
    if kern == "n2r":
        call, outs, bpe = (lambda L: L.hpgb_nest_to_ring(P(pix), P(out), n, nside, None, P(st), s)), [out], 16
    elif kern == "r2n":
        call, outs, bpe = (lambda L: L.hpgb_ring_to_nest(P(pix), P(out), n, nside, None, P(st), s)), [out], 16
    elif kern.startswith("p2a"):
        nest = 0 if kern.endswith("ring") else 1
        call, outs, bpe = (lambda L: L.hpgb_pixel_to_angle(P(pix), P(a), P(b), n, nside, None, nest, 1, 1, P(st), s)), [a, b], 24
    elif kern == "p2v":
        call, outs, bpe = (lambda L: L.hpgb_pixel_to_vector(P(pix), P(a), P(b), P(c), n, nside, None, 1, P(st), s)), [a, b, c], 32
    else:
        libs[0].hpgb_pixel_to_vector(P(pix), P(a), P(b), P(c), n, nside, None, 1, P(st), s)
        call, outs, bpe = (lambda L: L.hpgb_vector_to_pixel(P(a), P(b), P(c), P(out), n, nside, None, 1, P(st), s)), [out], 32
else:
    m = n // 8
    nest = 0 if kern.endswith("ring") else 1
    lon = torch.rand(m, dtype=torch.float64, device="cuda", generator=g) * 360.0
    lat = torch.rad2deg(torch.asin(torch.rand(m, dtype=torch.float64, device="cuda", generator=g) * 2 - 1))
    p4 = torch.randint(0, 12 * 4096 * 4096, (m,), dtype=torch.int64, device="cuda", generator=g)
    o8 = torch.empty((m, 8), dtype=torch.int64, device="cuda")
    w4 = torch.empty((m, 4), dtype=torch.float64, device="cuda")
    if kern.startswith("upg"):
        o4 = torch.empty((m, 4), dtype=torch.int64, device="cuda")
        p4 = p4 % (12 * 4096 * 4096 - 1)
        call, outs, bpe = (lambda L: L.hpgb_upgrade_pixels(P(p4), P(o4), m, 4096, 8192, nest, P(st), s)), [o4], 40
    elif kern.startswith("interp"):
        call, outs, bpe = (lambda L: L.hpgb_interp_weights(P(lon), P(lat), P(o8), P(w4), m, 4096, None, nest, 1, 1, P(st), s)), [o8[: m // 2], w4], 80
    else:
        call, outs, bpe = (lambda L: L.hpgb_neighbors(P(p4), P(o8), m, 4096, None, nest, P(st), s)), [o8], 72
libs[0].hpgb_status_reset(P(st), s)
ref = None
for L, p in zip(libs, paths):  # warm-up + result check against the first library
    for _ in range(2):
        assert call(L) == 0
    torch.cuda.synchronize()
    snap = [o.clone() for o in outs]
    if ref is None:
        ref = snap
    else:
        same = all(torch.equal(x.view(torch.int64), y.view(torch.int64)) for x, y in zip(ref, snap))
        print("%-40s result %s the first library's" % (os.path.basename(p), "==" if same else "DIFFERS FROM"))
tot = [0.0] * len(libs)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for r in range(rounds):
    for k, L in enumerate(libs):
        e0.record()
        for _ in range(reps):
            call(L)
        e1.record()
        torch.cuda.synchronize()
        tot[k] += e0.elapsed_time(e1) / reps
for k, p in enumerate(paths):
    ms = tot[k] / rounds
    print("%-12s %-40s %8.3f ms  %7.1f Gelem/s  %6.0f GB/s  (%.3f of 6533.5)" % (kern, os.path.basename(p), ms, m / ms / 1e6, bpe * m / ms / 1e6, bpe * m / ms / 1e6 / 6533.5))
assert st[0].item() == -1
