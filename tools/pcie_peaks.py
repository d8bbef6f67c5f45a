"""Host<->device copy rates of this box (pinned memory), alone and concurrently: the ceiling of the e2e path."""
import time
import torch
n = 1 << 28  # 2 GiB of float64
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n // 2, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.empty(n // 2, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
run(True, True, 2)
t = run(True, False); print("H2D alone      : %.1f GB/s" % (n * 8 / t / 1e9))
t = run(False, True); print("D2H alone      : %.1f GB/s" % (n * 4 / t / 1e9))
t = run(True, True); print("H2D + D2H (2:1): H2D %.1f GB/s, D2H %.1f GB/s -> angle_to_pixel e2e ceiling %.2f Gpts/s" % (n * 8 / t / 1e9, n * 4 / t / 1e9, n / 2 / t / 1e9))
