"""ncu target for the two headline kernels AT THE METRIC'S SIZE: angle_to_pixel and nest_to_ring on 2^LOG2 points
(default 2^30), operands either carved from one allocation (bench.py round 1 layout) or three separate tensors.
usage: python tools/ncu_headline.py [log2=30] [carved|separate] [reps=2]"""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from hpgeom_b200 import _lib

L = _lib.lib()
log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 30
layout = sys.argv[2] if len(sys.argv) > 2 else "separate"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
n = 1 << log2
nside = 2 ** 29
g = torch.Generator(device="cuda").manual_seed(1)
if layout == "carved":
    pool = torch.empty(3 * n, dtype=torch.float64, device="cuda")
    lon, lat, pix = pool[:n], pool[n:2 * n], pool[2 * n:].view(torch.int64)
else:
    lon = torch.empty(n, dtype=torch.float64, device="cuda")
    pix = torch.empty(n, dtype=torch.int64, device="cuda")
    lat = torch.empty(n, dtype=torch.float64, device="cuda")
torch.rand(n, dtype=torch.float64, device="cuda", generator=g, out=lon).mul_(360.0)
torch.rand(n, dtype=torch.float64, device="cuda", generator=g, out=lat).mul_(2.0).sub_(1.0).asin_().mul_(57.29577951308232)
st = torch.empty(2, dtype=torch.int64, device="cuda")
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
L.hpgb_status_reset(st.data_ptr(), s)
print("layout %s: lon %x lat %x pix %x" % (layout, lon.data_ptr(), lat.data_ptr(), pix.data_ptr()))
for _ in range(reps):
    assert L.hpgb_angle_to_pixel(lon.data_ptr(), lat.data_ptr(), pix.data_ptr(), n, nside, None, 1, 1, 1, st.data_ptr(), s) == 0
ring = lon.view(torch.int64)
for _ in range(reps):
    assert L.hpgb_nest_to_ring(pix.data_ptr(), ring.data_ptr(), n, nside, None, st.data_ptr(), s) == 0
torch.cuda.synchronize()
assert st[0].item() == -1
print("ok", n)
