"""angle_to_pixel at 2^LOG2: throughput vs the gap (pad, in elements) between the back-to-back operand arrays."""
import ctypes, sys, torch
sys.path.insert(0, ".")
from hpgeom_b200 import _lib
L = _lib.lib()
log2 = int(sys.argv[1]); n = 1 << log2; nside = 2 ** 29
g = torch.Generator(device="cuda").manual_seed(1)
maxpad = 1 << 27
pool = torch.empty(3 * n + 2 * maxpad, dtype=torch.float64, device="cuda")
st = torch.empty(2, dtype=torch.int64, device="cuda")
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
print("pool %x" % pool.data_ptr())
for pad in [0, 1 << 11, 1 << 14, 1 << 17, 1 << 20, 1 << 21, 1 << 22, 1 << 23, 1 << 24, 1 << 25, 1 << 26, 1 << 27, 3 << 24, 5 << 22]:
    lon = pool[:n]; lat = pool[n + pad: 2 * n + pad]; pix = pool[2 * (n + pad): 2 * (n + pad) + n].view(torch.int64)
    torch.rand(n, dtype=torch.float64, device="cuda", generator=g, out=lon).mul_(360.0)
    torch.rand(n, dtype=torch.float64, device="cuda", generator=g, out=lat).mul_(2.0).sub_(1.0).asin_().mul_(57.29577951308232)
    L.hpgb_status_reset(st.data_ptr(), s)
    f = lambda: L.hpgb_angle_to_pixel(lon.data_ptr(), lat.data_ptr(), pix.data_ptr(), n, nside, None, 1, 1, 1, st.data_ptr(), s)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    # n2r pix -> lon storage
    ring = lon.view(torch.int64)
    f2 = lambda: L.hpgb_nest_to_ring(pix.data_ptr(), ring.data_ptr(), n, nside, None, st.data_ptr(), s)
    for _ in range(3): f2()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10): f2()
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / 10
    print("pad %10d elems (%8.1f MiB): a2p %.3f ms %.1f Gpts/s   n2r %.3f ms %.1f Gpix/s" % (pad, pad * 8 / 2 ** 20, ms, n / ms / 1e6, ms2, n / ms2 / 1e6))
