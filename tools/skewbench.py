"""Does the relative placement of the input/output arrays matter? n2r at 2^LOG2 with the output skewed."""
import ctypes, sys, torch
sys.path.insert(0, ".")
from hpgeom_b200 import _lib
L = _lib.lib()
log2 = int(sys.argv[1]); n = 1 << log2; nside = 2 ** 29
g = torch.Generator(device="cuda").manual_seed(1)
pix = torch.randint(0, 12 * nside * nside, (n,), dtype=torch.int64, device="cuda", generator=g)
st = torch.empty(2, dtype=torch.int64, device="cuda")
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
big = torch.empty(n + (1 << 24), dtype=torch.int64, device="cuda")
print("pix %x big %x" % (pix.data_ptr(), big.data_ptr()))
for skew in [0, 2, 512, 4096, 1 << 15, 1 << 17, (1 << 18) + 512, 1 << 20, (1 << 22) + 4096 + 32]:
    out = big[skew:skew + n]
    L.hpgb_status_reset(st.data_ptr(), s)
    for _ in range(3):
        L.hpgb_nest_to_ring(pix.data_ptr(), out.data_ptr(), n, nside, None, st.data_ptr(), s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        L.hpgb_nest_to_ring(pix.data_ptr(), out.data_ptr(), n, nside, None, st.data_ptr(), s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("skew %10d elems: %.3f ms %.1f Gpix/s %.0f GB/s" % (skew, ms, n / ms / 1e6, 16 * n / ms / 1e6))
# plain copy for reference
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
out = big[:n]
for _ in range(3): out.copy_(pix)
torch.cuda.synchronize(); e0.record()
for _ in range(10): out.copy_(pix)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("torch copy: %.3f ms %.0f GB/s" % (ms, 16 * n / ms / 1e6))
