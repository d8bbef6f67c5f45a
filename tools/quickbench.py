import ctypes, torch, time, sys
sys.path.insert(0, '.')
from hpgeom_b200 import _lib
L = _lib.lib()
nside = 2**29
n = 1 << 28
g = torch.Generator(device="cuda").manual_seed(1)
pix = torch.randint(0, 12*nside*nside, (n,), dtype=torch.int64, device="cuda", generator=g)
out = torch.empty_like(pix)
st = torch.empty(2, dtype=torch.int64, device="cuda")
s = torch.cuda.current_stream().cuda_stream
for name in ["hpgb_nest_to_ring", "hpgb_ring_to_nest"]:
    f = getattr(L, name)
    L.hpgb_status_reset(st.data_ptr(), s)
    for _ in range(3): f(pix.data_ptr(), out.data_ptr(), n, nside, None, st.data_ptr(), s)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 10
    for _ in range(K): f(pix.data_ptr(), out.data_ptr(), n, nside, None, st.data_ptr(), s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/K
    print(name, "ms", ms, "Gpts/s", n/ms/1e6, "GB/s", 16*n/ms/1e6, "status", st[0].item())
# copy baseline
for _ in range(3): out.copy_(pix)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): out.copy_(pix)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/10
print("copy", "ms", ms, "GB/s", 16*n/ms/1e6)
