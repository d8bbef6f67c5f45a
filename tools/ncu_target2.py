"""ncu workload, round r01k: the kernels added / rewritten this session (one launch each after a warm-up)."""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from hpgeom_b200 import _lib

L = _lib.lib()
st = torch.full((2,), -1, dtype=torch.int64, device="cuda")
g = torch.Generator(device="cuda").manual_seed(1)
n = 1 << 24
nside = 4096
pix = torch.randint(0, 12 * nside * nside - 1, (n,), dtype=torch.int64, device="cuda", generator=g)
o8 = torch.empty(n * 8, dtype=torch.int64, device="cuda")
for rep in range(2):
    assert L.hpgb_neighbors(pix.data_ptr(), o8.data_ptr(), n, nside, None, 0, st.data_ptr(), None) == 0
    assert L.hpgb_upgrade_pixels(pix.data_ptr(), o8.data_ptr(), n, nside, 2 * nside, 0, st.data_ptr(), None) == 0
    for lens in (np.full(512, 250_000), np.random.default_rng(0).integers(0, 64, 4_000_000)):
        m = lens.size
        lo = np.cumsum(lens + 3) - lens
        rr = torch.from_numpy(np.stack([lo, lo + lens], 1).astype(np.int64)).cuda()
        offs = torch.empty(m + 1, dtype=torch.int64, device="cuda")
        assert L.hpgb_ranges_offsets(rr.data_ptr(), m, 0, offs.data_ptr(), st.data_ptr(), None) == 0
        total = int(offs[m].item())
        assert L.hpgb_ranges_expand(rr.data_ptr(), offs.data_ptr(), m, 0, total, o8.data_ptr(), None) == 0
torch.cuda.synchronize()
print("ok")
