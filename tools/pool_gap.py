import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hpgeom_b200 import _lib
L = _lib.lib()
P = lambda a: ctypes.c_void_p(a.ctypes.data)
n = 10 << 20
src = np.ones(n, dtype=np.uint8); dst = np.zeros(n, dtype=np.uint8)
L.hpgb_host_copy(P(dst), P(src), n)
for gap_us in (0, 100, 300, 1000, 3000):
    tot = 0.0
    reps = 100
    for _ in range(reps):
        t1 = time.perf_counter()
        while time.perf_counter() - t1 < gap_us * 1e-6: pass
        t = time.perf_counter(); L.hpgb_host_copy(P(dst), P(src), n); tot += time.perf_counter() - t
    print("gap %5d us: 10 MiB job %.3f ms" % (gap_us, tot / reps * 1e3))
