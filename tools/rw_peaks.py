import torch
n = 1 << 29  # 4 GiB of int64
x = torch.empty(n, dtype=torch.int64, device="cuda")
y = torch.empty(n, dtype=torch.int64, device="cuda")
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
ms = t(lambda: x.fill_(7)); print("fill_  (write only): %.3f ms  %.0f GB/s" % (ms, n * 8 / ms / 1e6))
ms = t(lambda: x.zero_()); print("zero_  (memset)    : %.3f ms  %.0f GB/s" % (ms, n * 8 / ms / 1e6))
ms = t(lambda: y.copy_(x)); print("copy_  (read+write): %.3f ms  %.0f GB/s" % (ms, 2 * n * 8 / ms / 1e6))
ms = t(lambda: x.sum()); print("sum    (read only) : %.3f ms  %.0f GB/s" % (ms, n * 8 / ms / 1e6))
