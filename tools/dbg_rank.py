import ctypes, sys, torch
sys.path.insert(0, ".")
from hpgeom_b200 import _lib
import hpgeom_b200 as hpg
L = _lib.lib()
NSIDE = 2 ** 29
rank = int(sys.argv[1])
dev = torch.device("cuda", 0)
n = 1 << 30
gen = torch.Generator(device=dev).manual_seed(12345 + rank)
pool = torch.empty(3 * n, dtype=torch.float64, device=dev)
lon, lat, pix = pool[:n], pool[n:2 * n], pool[2 * n:].view(torch.int64)
torch.rand(n, dtype=torch.float64, device=dev, generator=gen, out=lon).mul_(360.0)
torch.rand(n, dtype=torch.float64, device=dev, generator=gen, out=lat).mul_(2.0).sub_(1.0).asin_().mul_(180.0 / 3.141592653589793)
status = torch.full((2,), -1, dtype=torch.int64, device=dev)
stp = ctypes.c_void_p(status.data_ptr())
L.hpgb_status_reset(stp, None)
assert L.hpgb_angle_to_pixel(lon.data_ptr(), lat.data_ptr(), pix.data_ptr(), n, NSIDE, None, 1, 1, 1, stp, None) == 0
torch.cuda.synchronize(); print("a2p status", int(status[0]))
npix = 12 * NSIDE * NSIDE
print("pix range", int(pix.min()), int(pix.max()), npix)
keep = pix.clone()
ring = lon.view(torch.int64)
L.hpgb_status_reset(stp, None)
assert L.hpgb_nest_to_ring(pix.data_ptr(), ring.data_ptr(), n, NSIDE, None, stp, None) == 0
torch.cuda.synchronize(); print("n2r status", int(status[0]), "ring range", int(ring.min()), int(ring.max()))
m = 1 << 27
L.hpgb_status_reset(stp, None)
assert L.hpgb_ring_to_nest(ring.data_ptr(), pix.data_ptr(), m, NSIDE, None, stp, None) == 0
torch.cuda.synchronize(); print("r2n status", int(status[0]))
bad = torch.nonzero(pix[:m] != keep[:m]).flatten()
print("round trip mismatches", bad.numel(), bad[:5].tolist())
if bad.numel():
    i = int(bad[0]); print("idx", i, "orig", int(keep[i]), "ring", int(ring[i]), "back", int(pix[i]))
    import oracle
    port = oracle.port()
    import numpy as np
    print("oracle n2r", port.nest_to_ring(NSIDE, np.array([int(keep[i])])), "oracle r2n", port.ring_to_nest(NSIDE, np.array([int(ring[i])])))
f1 = torch.empty(m, dtype=torch.float64, device=dev); f2 = torch.empty(m, dtype=torch.float64, device=dev)
L.hpgb_status_reset(stp, None)
assert L.hpgb_pixel_to_angle(pix.data_ptr(), f1.data_ptr(), f2.data_ptr(), m, NSIDE, None, 1, 1, 1, stp, None) == 0
torch.cuda.synchronize(); print("p2a status", int(status[0]))
