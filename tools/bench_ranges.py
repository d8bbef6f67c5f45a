"""Throughput of the pixel_ranges_to_pixels / query_circle kernels (device resident, CUDA events)."""
import ctypes
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpgeom_b200 as hpg
from hpgeom_b200 import _lib

L = _lib.lib()
PEAK = 6533.5


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / reps


def expand_case(name, lens, rng):
    m = lens.size
    lo = np.cumsum(lens + 3) - lens
    rr = torch.from_numpy(np.stack([lo, lo + lens], 1).astype(np.int64)).cuda()
    offs = torch.empty(m + 1, dtype=torch.int64, device="cuda")
    st = torch.full((2,), -1, dtype=torch.int64, device="cuda")
    t_off = timeit(lambda: L.hpgb_ranges_offsets(rr.data_ptr(), m, 0, offs.data_ptr(), st.data_ptr(), None))
    total = int(offs[m].item())
    out = torch.empty(total, dtype=torch.int64, device="cuda")
    t_exp = timeit(lambda: L.hpgb_ranges_expand(rr.data_ptr(), offs.data_ptr(), m, 0, total, out.data_ptr(), None))
    gbs = total * 8 / t_exp / 1e6
    print("%-28s m=%9d pixels=%11d  offsets %.3f ms  expand %.3f ms  %.1f Gpix/s  %.0f GB/s (%.1f%% of %.1f)"
          % (name, m, total, t_off, t_exp, total / t_exp / 1e6, gbs, 100 * gbs / PEAK, PEAK))


rng = np.random.default_rng(0)
expand_case("long ranges (1e6 each)", np.full(512, 1_000_000), rng)
expand_case("query-like (4096 each)", np.full(131072, 4096), rng)
expand_case("short ragged (0..63)", rng.integers(0, 64, 16_000_000), rng)
expand_case("upgrade-like (4 each)", np.full(64_000_000, 4), rng)

for nside, radius in ((2 ** 14, 0.5), (2 ** 16, 0.1), (2 ** 20, 0.005)):
    plan = _lib.Disc()
    L.hpgb_query_circle_ring_plan(nside, 0.7, 4.0, radius, 0, 1, 0, 0, ctypes.byref(plan))
    m = int(plan.m)
    table = torch.empty((m, 2), dtype=torch.int64, device="cuda")
    t = timeit(lambda: L.hpgb_query_circle_ring_ranges(ctypes.byref(plan), table.data_ptr(), None, None))
    pix = hpg.pixel_ranges_to_pixels(table)
    print("query_circle ring nside=2^%d radius=%g: %d rings in %.3f ms (%.2f Grings/s), %d pixels"
          % (int(np.log2(nside)), radius, (m - 2) // 2, t, (m - 2) / 2 / t / 1e6, pix.numel()))
