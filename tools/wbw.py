"""Write-only / read-only HBM bandwidth with torch primitives (context for the write-heavy kernels)."""
import torch
n = 1 << 30
a = torch.empty(n, dtype=torch.int64, device="cuda")
def t(f, nb, name, k=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / k
    print("%-28s %.3f ms  %.0f GB/s" % (name, ms, nb / ms / 1e6))
t(lambda: a.fill_(7), 8 * n, "fill_ (write only)")
t(lambda: a.zero_(), 8 * n, "zero_ (memset)")
t(lambda: a.sum(), 8 * n, "sum (read only)")
b = torch.empty(n // 2, dtype=torch.int64, device="cuda")
t(lambda: b.copy_(a[: n // 2]), 8 * n, "copy (1:1 read:write)")
c = a[: n // 8]
import sys; sys.stdout.flush()
o = a[n // 2: n // 2 + n // 8 * 4].view(-1, 4)
t(lambda: o.copy_(c[:, None].expand(-1, 4)), 8 * (n // 8) * 5, "broadcast 1 read : 4 write")
