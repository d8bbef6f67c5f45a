"""Rate of the host pipeline's staging copy (hpgb_host_copy: the pool's parallel memcpy) against numpy's copy, on this
host.  No GPU needed.  usage: python tools/host_copy_rate.py [MiB=512] [reps=5]; HPGB_HOST_THREADS=n sets the pool size."""
import ctypes
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hpgeom_b200 import _lib

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
L = _lib.lib()
n = mib << 20
src = np.random.default_rng(1).integers(0, 255, n, dtype=np.uint8)
dst = np.zeros(n, dtype=np.uint8)
P = lambda a: ctypes.c_void_p(a.ctypes.data)
best = 1e9
for _ in range(reps):
    t = time.perf_counter()
    rc = L.hpgb_host_copy(P(dst), P(src), n)
    best = min(best, time.perf_counter() - t)
    assert rc == 0
assert np.array_equal(dst, src)
bnp = 1e9
for _ in range(reps):
    t = time.perf_counter()
    np.copyto(dst, src)
    bnp = min(bnp, time.perf_counter() - t)
print("hpgb_host_copy %d MiB: %.1f GB/s   (numpy copyto, one thread: %.1f GB/s)   HPGB_HOST_THREADS=%s cpus=%d"
      % (mib, n / best / 1e9, n / bnp / 1e9, os.environ.get("HPGB_HOST_THREADS", "default"), os.cpu_count()))
