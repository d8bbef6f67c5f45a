"""ncu workload: boundaries (step 1) at nside 4096, 2^22 pixels."""
import sys
import torch
sys.path.insert(0, ".")
from hpgeom_b200 import _lib
L = _lib.lib()
st = torch.full((2,), -1, dtype=torch.int64, device="cuda")
g = torch.Generator(device="cuda").manual_seed(1)
n = 1 << 22
nside = 4096
pix = torch.randint(0, 12 * nside * nside, (n,), dtype=torch.int64, device="cuda", generator=g)
a = torch.empty(n * 4, dtype=torch.float64, device="cuda")
b = torch.empty(n * 4, dtype=torch.float64, device="cuda")
for rep in range(2):
    for nest in (1, 0):
        assert L.hpgb_boundaries(pix.data_ptr(), a.data_ptr(), b.data_ptr(), n, nside, None, 1, nest, 1, 1, st.data_ptr(), None) == 0
torch.cuda.synchronize()
print("ok")
