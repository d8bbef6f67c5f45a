"""ncu workload: interpolation weights (nest, ring) and NEST neighbours at nside 4096, 2^23 points."""
import sys
import torch
sys.path.insert(0, ".")
from hpgeom_b200 import _lib
L = _lib.lib()
st = torch.full((2,), -1, dtype=torch.int64, device="cuda")
g = torch.Generator(device="cuda").manual_seed(1)
n = 1 << 23
nside = 4096
lon = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 360.0
lat = torch.rad2deg(torch.asin(torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1))
pix = torch.randint(0, 12 * nside * nside - 1, (n,), dtype=torch.int64, device="cuda", generator=g)
o8 = torch.empty(n * 8, dtype=torch.int64, device="cuda")
w4 = torch.empty(n * 4, dtype=torch.float64, device="cuda")
for rep in range(2):
    for nest in (1, 0):
        assert L.hpgb_interp_weights(lon.data_ptr(), lat.data_ptr(), o8.data_ptr(), w4.data_ptr(), n, nside, None, nest, 1, 1, st.data_ptr(), None) == 0
    assert L.hpgb_neighbors(pix.data_ptr(), o8.data_ptr(), n, nside, None, 1, st.data_ptr(), None) == 0
torch.cuda.synchronize()
print("ok")
