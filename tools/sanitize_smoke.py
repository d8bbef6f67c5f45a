"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck): odd sizes, both schemes,
ring-table and per-pixel forms.  usage: compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import hpgeom_b200 as hpg

rng = np.random.default_rng(3)
for nside, n in ((2 ** 29, 70_001), (4096, 1_100_003), (1024, 150_001), (3, 20_001)):
    npix = 12 * nside * nside
    pow2 = nside & (nside - 1) == 0
    lon = torch.from_numpy(360.0 * rng.random(n)).cuda()
    lat = torch.from_numpy(np.degrees(np.arcsin(2.0 * rng.random(n) - 1.0))).cuda()
    pix = torch.from_numpy(rng.integers(0, npix, n)).cuda()
    for nest in ((True, False) if pow2 else (False,)):
        p = hpg.angle_to_pixel(nside, lon, lat, nest=nest)
        a, b = hpg.pixel_to_angle(nside, pix, nest=nest)
        x, y, z = hpg.pixel_to_vector(nside, pix, nest=nest)
        q = hpg.vector_to_pixel(nside, x, y, z, nest=nest)
        assert torch.equal(q, pix)
        nb = hpg.neighbors(nside, pix[:50_001], nest=nest)
        ip, iw = hpg.get_interpolation_weights(nside, lon, lat, nest=nest)
        assert float((iw.sum(dim=1) - 1).abs().max()) < 1e-9
        if pow2 and nside < 2 ** 29:
            up = hpg.upgrade_pixels(nside, pix[:50_001] % (npix - 1), 2 * nside, nest=nest)
    if pow2:
        r = hpg.nest_to_ring(nside, pix)
        assert torch.equal(hpg.ring_to_nest(nside, r), pix) or nside == 2 ** 29
    torch.cuda.synchronize()
    print("nside", nside, "ok")
lonh, lath = 360.0 * rng.random(300_001), np.degrees(np.arcsin(2.0 * rng.random(300_001) - 1.0))
ph = hpg.angle_to_pixel(2 ** 20, lonh, lath)  # host pipeline, pageable buffers
assert np.array_equal(ph, hpg.angle_to_pixel(2 ** 20, torch.from_numpy(lonh).cuda(), torch.from_numpy(lath).cuda()).cpu().numpy())
print("sanitize smoke ok")
