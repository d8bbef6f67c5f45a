"""Device-resident timing of individual kernels through the C ABI (CUDA events, 2^LOG2 points)."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from hpgeom_b200 import _lib

L = _lib.lib()
log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 28
which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["a2p", "n2r", "r2n"]
n = 1 << log2
nside = 2 ** 29
g = torch.Generator(device="cuda").manual_seed(1)
lon = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 360.0
lat = torch.rad2deg(torch.asin(torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1))
pix = torch.randint(0, 12 * nside * nside, (n,), dtype=torch.int64, device="cuda", generator=g)
out = torch.empty(n, dtype=torch.int64, device="cuda")
st = torch.empty(2, dtype=torch.int64, device="cuda")
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: ctypes.c_void_p(t.data_ptr())


def bench(name, fn, nbytes, K=10):
    L.hpgb_status_reset(P(st), s)
    for _ in range(3):
        assert fn() == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print("%-22s %8.3f ms  %7.1f Gpts/s  %7.1f GB/s  (%.1f%% of 6533.5)  status %d" % (
        name, ms, n / ms / 1e6, nbytes * n / ms / 1e6, nbytes * n / ms / 1e6 / 65.335, st[0].item()))


if "a2p" in which:
    bench("a2p nest deg", lambda: L.hpgb_angle_to_pixel(P(lon), P(lat), P(out), n, nside, None, 1, 1, 1, P(st), s), 24)
if "a2p_ring" in which:
    bench("a2p ring deg", lambda: L.hpgb_angle_to_pixel(P(lon), P(lat), P(out), n, nside, None, 0, 1, 1, P(st), s), 24)
if "a2p_rad" in which:
    lr, br = torch.deg2rad(lon), torch.deg2rad(lat)
    bench("a2p nest rad", lambda: L.hpgb_angle_to_pixel(P(lr), P(br), P(out), n, nside, None, 1, 1, 0, P(st), s), 24)
    th = (3.141592653589793 / 2 - br).clamp_(0, 3.141592653589793)
    bench("a2p nest thetaphi", lambda: L.hpgb_angle_to_pixel(P(th), P(lr), P(out), n, nside, None, 1, 0, 0, P(st), s), 24)
    del lr, br, th
if "n2r" in which:
    bench("n2r", lambda: L.hpgb_nest_to_ring(P(pix), P(out), n, nside, None, P(st), s), 16)
if "r2n" in which:
    bench("r2n", lambda: L.hpgb_ring_to_nest(P(pix), P(out), n, nside, None, P(st), s), 16)
a = b = c = None
if any(w in which for w in ("p2a", "p2v", "v2p", "lonlat")):
    a = torch.empty(n, dtype=torch.float64, device="cuda")
    b = torch.empty(n, dtype=torch.float64, device="cuda")
    c = torch.empty(n, dtype=torch.float64, device="cuda")
if "p2a" in which:
    bench("p2a nest deg", lambda: L.hpgb_pixel_to_angle(P(pix), P(a), P(b), n, nside, None, 1, 1, 1, P(st), s), 24)
    bench("p2a ring deg", lambda: L.hpgb_pixel_to_angle(P(pix), P(a), P(b), n, nside, None, 0, 1, 1, P(st), s), 24)
if "p2v" in which:
    bench("p2v nest", lambda: L.hpgb_pixel_to_vector(P(pix), P(a), P(b), P(c), n, nside, None, 1, P(st), s), 32)
if "v2p" in which:
    L.hpgb_pixel_to_vector(P(pix), P(a), P(b), P(c), n, nside, None, 1, P(st), s)
    bench("v2p nest", lambda: L.hpgb_vector_to_pixel(P(a), P(b), P(c), P(out), n, nside, None, 1, P(st), s), 32)
if "neigh" in which:
    m = n // 8
    o8 = torch.empty((m, 8), dtype=torch.int64, device="cuda")
    p4 = pix[:m] % (12 * 4096 * 4096)
    def f():
        return L.hpgb_neighbors(P(p4), P(o8), m, 4096, None, 1, P(st), s)
    n_save = n
    n = m
    bench("neighbors nest 4096", f, 72)
    bench("neighbors ring 4096", lambda: L.hpgb_neighbors(P(p4), P(o8), m, 4096, None, 0, P(st), s), 72)
    n = n_save
    del o8
if "interp" in which:
    m = n // 8
    o4 = torch.empty((m, 4), dtype=torch.int64, device="cuda")
    w4 = torch.empty((m, 4), dtype=torch.float64, device="cuda")
    n_save = n
    n = m
    bench("interp nest 4096", lambda: L.hpgb_interp_weights(P(lon), P(lat), P(o4), P(w4), m, 4096, None, 1, 1, 1, P(st), s), 80)
    bench("interp ring 4096", lambda: L.hpgb_interp_weights(P(lon), P(lat), P(o4), P(w4), m, 4096, None, 0, 1, 1, P(st), s), 80)
    n = n_save
    del o4, w4
if "upgrade" in which:
    m = n // 4
    p4 = pix[:m] % (12 * 4096 * 4096 - 1)
    n_save = n
    n = m
    bench("upgrade nest 4096->8192", lambda: L.hpgb_upgrade_pixels(P(p4), P(out), m, 4096, 8192, 1, P(st), s), 40)
    bench("upgrade ring 4096->8192", lambda: L.hpgb_upgrade_pixels(P(p4), P(out), m, 4096, 8192, 0, P(st), s), 40)
    n = n_save
if "bounds" in which:
    m = n // 8
    p4 = pix[:m] % (12 * 4096 * 4096)
    ba = torch.empty((m, 4), dtype=torch.float64, device="cuda")
    bb = torch.empty((m, 4), dtype=torch.float64, device="cuda")
    n_save = n
    n = m
    bench("boundaries nest 4096 step 1", lambda: L.hpgb_boundaries(P(p4), P(ba), P(bb), m, 4096, None, 1, 1, 1, 1, P(st), s), 72)
    bench("boundaries ring 4096 step 1", lambda: L.hpgb_boundaries(P(p4), P(ba), P(bb), m, 4096, None, 1, 0, 1, 1, P(st), s), 72)
    n = n_save
    del ba, bb
if "lonlat" in which:
    bench("lonlat_to_thetaphi deg", lambda: L.hpgb_lonlat_to_thetaphi(P(lon), P(lat), P(a), P(b), n, 1, P(st), s), 32)
