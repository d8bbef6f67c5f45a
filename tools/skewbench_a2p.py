"""angle_to_pixel at 2^LOG2: throughput vs relative placement of lon / lat / pix in HBM."""
import ctypes, sys, torch
sys.path.insert(0, ".")
from hpgeom_b200 import _lib
L = _lib.lib()
log2 = int(sys.argv[1]); n = 1 << log2; nside = 2 ** 29
g = torch.Generator(device="cuda").manual_seed(1)
pad = 1 << 25
pool = torch.empty(3 * (n + pad), dtype=torch.float64, device="cuda")
st = torch.empty(2, dtype=torch.int64, device="cuda")
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
lon0 = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 360.0
lat0 = torch.rad2deg(torch.asin(torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1))
print("pool %x" % pool.data_ptr())
for (s1, s2) in [(0, 0), (512, 1024), (0, 4096), (4096, 0), (1 << 17, 1 << 18), (1 << 20, 3 << 20), ((1 << 24), (1 << 23)), (12345 * 2, 54321 * 2)]:
    lon = pool[:n]; lat = pool[n + pad + s1: n + pad + s1 + n]; pix = pool[2 * (n + pad) + s2: 2 * (n + pad) + s2 + n].view(torch.int64)
    lon.copy_(lon0); lat.copy_(lat0)
    L.hpgb_status_reset(st.data_ptr(), s)
    f = lambda: L.hpgb_angle_to_pixel(lon.data_ptr(), lat.data_ptr(), pix.data_ptr(), n, nside, None, 1, 1, 1, st.data_ptr(), s)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("lat skew %9d pix skew %9d: %.3f ms %.1f Gpts/s %.0f GB/s" % (s1, s2, ms, n / ms / 1e6, 24 * n / ms / 1e6))
