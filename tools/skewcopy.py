"""torch copy_ (the MEASURED_PEAKS.json method) vs relative placement of source and destination."""
import sys, torch
log2 = int(sys.argv[1]); n = 1 << log2
src = torch.randint(0, 1 << 40, (n,), dtype=torch.int64, device="cuda")
big = torch.empty(n + (1 << 24), dtype=torch.int64, device="cuda")
print("src %x dst %x" % (src.data_ptr(), big.data_ptr()))
for skew in [0, 512, 4096, 1 << 17, 1 << 20, (1 << 22) + 4096 + 32]:
    out = big[skew:skew + n]
    for _ in range(3): out.copy_(src)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): out.copy_(src)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("skew %10d elems: %.3f ms %.0f GB/s" % (skew, ms, 16 * n / ms / 1e6))
