"""Per-launch time series of angle_to_pixel under sustained load, next to the SM clock / power nvidia-smi reports.
usage: python tools/sustained_a2p.py [log2n=30] [launches=300] [kernel=a2p|n2r|copy]"""
import ctypes
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, ".")
from hpgeom_b200 import _lib

L = _lib.lib()
log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 30
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 300
kern = sys.argv[3] if len(sys.argv) > 3 else "a2p"
n = 1 << log2
nside = 2 ** 29
g = torch.Generator(device="cuda").manual_seed(1)
lon = torch.empty(n, dtype=torch.float64, device="cuda")
pix = torch.empty(n, dtype=torch.int64, device="cuda")
lat = torch.empty(n, dtype=torch.float64, device="cuda")
torch.rand(n, dtype=torch.float64, device="cuda", generator=g, out=lon).mul_(360.0)
torch.rand(n, dtype=torch.float64, device="cuda", generator=g, out=lat).mul_(2.0).sub_(1.0).asin_().mul_(57.29577951308232)
st = torch.empty(2, dtype=torch.int64, device="cuda")
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
L.hpgb_status_reset(st.data_ptr(), s)
if kern == "a2p":
    step = lambda: L.hpgb_angle_to_pixel(lon.data_ptr(), lat.data_ptr(), pix.data_ptr(), n, nside, None, 1, 1, 1, st.data_ptr(), s)
    bpe = 24
elif kern == "n2r":
    L.hpgb_angle_to_pixel(lon.data_ptr(), lat.data_ptr(), pix.data_ptr(), n, nside, None, 1, 1, 1, st.data_ptr(), s)
    ring = lon.view(torch.int64)
    step = lambda: L.hpgb_nest_to_ring(pix.data_ptr(), ring.data_ptr(), n, nside, None, st.data_ptr(), s)
    bpe = 16
else:
    step = lambda: lat.copy_(lon)
    bpe = 16
samples = []
stop = False


def poll():
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_throttle_reasons.active",
                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
    t0 = time.perf_counter()
    while not stop:
        line = p.stdout.readline()
        if not line:
            break
        samples.append((time.perf_counter() - t0, line.strip()))
    p.kill()


th = threading.Thread(target=poll, daemon=True)
th.start()
time.sleep(0.5)
for _ in range(3):
    step()
torch.cuda.synchronize()
time.sleep(0.3)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(nl + 1)]
t_start = time.perf_counter()
ev[0].record()
for i in range(nl):
    step()
    ev[i + 1].record()
torch.cuda.synchronize()
t_end = time.perf_counter()
stop = True
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(nl)]
print("%s 2^%d x %d launches, total %.2f s" % (kern, log2, nl, t_end - t_start))
blk_n = max(nl // 15, 1)
print("  first 12 launches (ms): " + " ".join("%.3f" % v for v in ms[:12]))
for lo in range(0, nl, blk_n):
    blk = ms[lo:lo + blk_n]
    m = sum(blk) / len(blk)
    print("  launches %4d-%4d: %.3f ms  %.1f Gelem/s  %.0f GB/s" % (lo, lo + len(blk) - 1, m, n / m / 1e6, bpe * n / m / 1e6))
print("nvidia-smi (t since poll start, sm MHz, mem MHz, W, C, throttle reasons) -- load starts at ~0.8 s:")
for t, l in samples[::max(len(samples) // 40, 1)]:
    print("  %.2f s  %s" % (t, l))
