#!/usr/bin/env python
"""Generate the golden vectors in tests/golden/ from the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference):

    make -C oracle ref && python tests/golden/make_golden.py

The C functions come from oracle/_ref/_hpgeom*.so (the reference's four C files compiled by
oracle/Makefile); the Python-level functions (lonlat_to_thetaphi, upgrade_pixels) are executed
from /root/reference/hpgeom/hpgeom.py itself, bound to that same compiled module.  Nothing from
hpgeom_b200 or from oracle/hpg_oracle.c is involved, so the vectors pin both of them.

Outputs: golden_<function>.npz (inputs + reference outputs, keyed by case) and
golden_errors.json (exception type + message for the reference's argument checks).
"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402

REF_SRC = "/root/reference/hpgeom/hpgeom.py"
N = 1024
SEED = 12345


def load_reference_python():
    """Execute the reference's hpgeom.py against the compiled reference extension."""
    pkg = types.ModuleType("hpgeom")
    pkg.__path__ = []
    sys.modules["hpgeom"] = pkg
    sys.modules["hpgeom._hpgeom"] = oracle.ref
    mod = types.ModuleType("hpgeom.hpgeom")
    mod.__package__ = "hpgeom"
    with open(REF_SRC) as f:
        code = compile(f.read(), REF_SRC, "exec")
    exec(code, mod.__dict__)
    return mod


def sky(rng, n):
    lon = rng.uniform(0.0, 360.0, n)
    lat = np.degrees(np.arcsin(rng.uniform(-1.0, 1.0, n)))
    # edge cases the reference accepts
    edge_lon = np.array([0.0, 360.0, -10.0, 725.5, 90.0, 180.0, 270.0, 45.0, 1e-300, 359.99999999999994])
    edge_lat = np.array([90.0, -90.0, 0.0, 41.810314895778596, -41.810314895778596, 89.999999, -89.999999,
                         1e-12, -1e-12, 66.44353569089877])
    lon[:edge_lon.size] = edge_lon
    lat[:edge_lat.size] = edge_lat
    # a band near the poles where the sin(theta) branch is live (|z| >= 0.99)
    lat[20:60] = 90.0 - rng.uniform(0.0, 8.2, 40)
    lat[60:100] = -90.0 + rng.uniform(0.0, 8.2, 40)
    lat[100:120] = 90.0 - 10.0 ** rng.uniform(-9, -1, 20)
    return lon, lat


def pixels(rng, nside, n):
    npix = 12 * nside * nside
    pix = rng.integers(0, npix, n)
    ncap = 2 * nside * (nside - 1)
    edge = np.array([0, 1, 2, 3, npix - 1, npix - 2, npix - 4, ncap - 1, ncap, ncap + 1,
                     npix - ncap - 1, npix - ncap, npix // 2, npix // 2 - 1, 4 * nside - 1,
                     nside * nside - 1, nside * nside, 4 * nside * nside, 8 * nside * nside - 1], dtype=np.int64)
    edge = edge[(edge >= 0) & (edge < npix)]
    pix[:edge.size] = edge
    if npix <= n:
        pix[:npix] = np.arange(npix)
    return pix.astype(np.int64)


def main():
    r = oracle.ref
    assert r is not None, "build oracle/_ref first (make -C oracle ref)"
    hp = load_reference_python()
    rng = np.random.default_rng(SEED)

    pow2 = [1, 2, 4, 32, 1024, 4096, 2 ** 20, 2 ** 29]
    ring_only = [3, 924, 100003]

    # ---------------- angle_to_pixel ----------------
    out = {}
    lon, lat = sky(rng, N)
    theta = np.arccos(rng.uniform(-1, 1, N))
    phi = rng.uniform(0, 2 * np.pi, N)
    theta[:6] = [0.0, np.pi, 0.005, np.pi - 0.005, 0.01, 3.14159 - 0.01]
    phi[:6] = [0.0, 2 * np.pi, np.pi / 2, np.pi, 3 * np.pi / 2, 6.283185307179586]
    out["lon"], out["lat"], out["theta"], out["phi"] = lon, lat, theta, phi
    for nside in pow2 + ring_only:
        for nest in ([True, False] if nside in pow2 else [False]):
            tag = "%d_%s" % (nside, "nest" if nest else "ring")
            out["deg_" + tag] = r.angle_to_pixel(nside, lon, lat, nest=nest)
            out["rad_" + tag] = r.angle_to_pixel(nside, np.radians(lon), np.radians(lat), nest=nest, degrees=False)
            out["tp_" + tag] = r.angle_to_pixel(nside, theta, phi, nest=nest, lonlat=False)
    np.savez_compressed(os.path.join(HERE, "golden_angle_to_pixel.npz"), **out)

    # ---------------- pixel_to_angle / pixel_to_vector / nest<->ring / neighbors ----------------
    p2a, p2v, conv, nb = {}, {}, {}, {}
    for nside in pow2 + ring_only:
        pix = pixels(rng, nside, N)
        p2a["pix_%d" % nside] = pix
        for nest in ([True, False] if nside in pow2 else [False]):
            tag = "%d_%s" % (nside, "nest" if nest else "ring")
            a, b = r.pixel_to_angle(nside, pix, nest=nest)
            p2a["deg_a_" + tag], p2a["deg_b_" + tag] = a, b
            a, b = r.pixel_to_angle(nside, pix, nest=nest, degrees=False)
            p2a["rad_a_" + tag], p2a["rad_b_" + tag] = a, b
            a, b = r.pixel_to_angle(nside, pix, nest=nest, lonlat=False)
            p2a["tp_a_" + tag], p2a["tp_b_" + tag] = a, b
            x, y, z = r.pixel_to_vector(nside, pix, nest=nest)
            p2v["x_" + tag], p2v["y_" + tag], p2v["z_" + tag] = x, y, z
            nb["nb_" + tag] = r.neighbors(nside, pix[:256], nest=nest)
        if nside in pow2:
            conv["n2r_%d" % nside] = r.nest_to_ring(nside, pix)
            conv["r2n_%d" % nside] = r.ring_to_nest(nside, pix)
    np.savez_compressed(os.path.join(HERE, "golden_pixel_to_angle.npz"), **p2a)
    p2v.update({k: v for k, v in p2a.items() if k.startswith("pix_")})
    np.savez_compressed(os.path.join(HERE, "golden_pixel_to_vector.npz"), **p2v)
    conv.update({k: v for k, v in p2a.items() if k.startswith("pix_")})
    np.savez_compressed(os.path.join(HERE, "golden_nest_ring.npz"), **conv)
    nb.update({k: v[:256] for k, v in p2a.items() if k.startswith("pix_")})
    np.savez_compressed(os.path.join(HERE, "golden_neighbors.npz"), **nb)

    # ---------------- vector_to_pixel ----------------
    v2p = {}
    x, y, z = rng.uniform(-1, 1, N), rng.uniform(-1, 1, N), rng.uniform(-1, 1, N)
    x[:6] = [0.0, 0.0, 1.0, -1.0, 1e-3, 0.0]
    y[:6] = [0.0, 0.0, 0.0, 0.0, 1e-3, -2.0]
    z[:6] = [1.0, -3.0, 0.0, 0.0, 1.0, 0.0]
    x[10:40] *= 0.05
    y[10:40] *= 0.05  # |nz| > 0.99 branch
    v2p["x"], v2p["y"], v2p["z"] = x, y, z
    for nside in pow2 + ring_only:
        for nest in ([True, False] if nside in pow2 else [False]):
            v2p["%d_%s" % (nside, "nest" if nest else "ring")] = r.vector_to_pixel(nside, x, y, z, nest=nest)
    np.savez_compressed(os.path.join(HERE, "golden_vector_to_pixel.npz"), **v2p)

    # ---------------- get_interpolation_weights ----------------
    it = {}
    lon, lat = sky(rng, 512)
    it["lon"], it["lat"] = lon, lat
    for nside in [1, 2, 32, 4096, 2 ** 20, 2 ** 29, 3, 924]:
        for nest in ([True, False] if (nside & (nside - 1)) == 0 else [False]):
            tag = "%d_%s" % (nside, "nest" if nest else "ring")
            p, w = r.get_interpolation_weights(nside, lon, lat, nest=nest)
            it["pix_" + tag], it["wgt_" + tag] = p, w
    np.savez_compressed(os.path.join(HERE, "golden_interp.npz"), **it)

    # ---------------- upgrade_pixels / lonlat_to_thetaphi (reference Python) ----------------
    up = {}
    for nside, nup in [(1, 1), (1, 4), (32, 64), (4096, 8192), (2 ** 20, 2 ** 22), (2 ** 28, 2 ** 29)]:
        npix = 12 * nside * nside
        pix = rng.integers(0, npix - 1, 200).astype(np.int64)  # last pixel raises in the reference (quirk)
        pix[:3] = [0, npix - 2, npix // 3]
        up["pix_%d_%d" % (nside, nup)] = pix
        for nest in [True, False]:
            up["up_%d_%d_%s" % (nside, nup, "nest" if nest else "ring")] = hp.upgrade_pixels(nside, pix, nup, nest=nest)
    np.savez_compressed(os.path.join(HERE, "golden_upgrade.npz"), **up)

    ll = {}
    lon = rng.uniform(-720, 720, N)
    lat = rng.uniform(-90, 90, N)
    lon[:5] = [0.0, 360.0, -360.0, -1e-20, 719.9999]
    lat[:3] = [90.0, -90.0, 0.0]
    ll["lon"], ll["lat"] = lon, lat
    ll["deg_theta"], ll["deg_phi"] = hp.lonlat_to_thetaphi(lon, lat, degrees=True)
    ll["rad_theta"], ll["rad_phi"] = hp.lonlat_to_thetaphi(np.radians(lon), np.radians(lat) * (1 - 1e-16), degrees=False)
    ll["rad_lon"], ll["rad_lat"] = np.radians(lon), np.radians(lat) * (1 - 1e-16)
    np.savez_compressed(os.path.join(HERE, "golden_lonlat.npz"), **ll)

    # ---------------- error behaviour ----------------
    errs = []

    def rec(name, fn, *args, **kw):
        try:
            fn(*args, **kw)
            errs.append({"call": name, "args": repr((args, kw)), "type": None, "msg": None})
        except Exception as e:  # noqa: BLE001
            errs.append({"call": name, "args": repr((args, kw)), "type": type(e).__name__, "msg": str(e)})

    rec("angle_to_pixel", r.angle_to_pixel, -1, 0.0, 0.0)
    rec("angle_to_pixel", r.angle_to_pixel, 0, 0.0, 0.0)
    rec("angle_to_pixel", r.angle_to_pixel, 2049, 0.0, 0.0, nest=True)
    rec("angle_to_pixel", r.angle_to_pixel, 2 ** 30, 0.0, 0.0, nest=False)
    rec("angle_to_pixel", r.angle_to_pixel, 2 ** 29 + 1, 0.0, 0.0, nest=False)
    rec("angle_to_pixel", r.angle_to_pixel, 1024, 0.0, 100.0)
    rec("angle_to_pixel", r.angle_to_pixel, 1024, 0.0, -90.00001)
    rec("angle_to_pixel", r.angle_to_pixel, 1024, 0.0, 1.6, degrees=False)
    rec("angle_to_pixel", r.angle_to_pixel, 1024, -0.1, 0.0, lonlat=False)
    rec("angle_to_pixel", r.angle_to_pixel, 1024, 3.2, 0.0, lonlat=False)
    rec("angle_to_pixel", r.angle_to_pixel, 1024, 1.0, -0.1, lonlat=False)
    rec("angle_to_pixel", r.angle_to_pixel, 1024, 1.0, 6.3, lonlat=False)
    rec("angle_to_pixel", r.angle_to_pixel, 1024, np.zeros(3), np.zeros(4))
    rec("angle_to_pixel", r.angle_to_pixel, 1024, np.array([0.0, 1.0, 2.0]), np.array([0.0, 95.5, -123.25]))
    rec("angle_to_pixel", r.angle_to_pixel, -5, np.zeros(0), np.zeros(0))
    rec("pixel_to_angle", r.pixel_to_angle, 1024, -1)
    rec("pixel_to_angle", r.pixel_to_angle, 1024, 12 * 1024 * 1024)
    rec("pixel_to_angle", r.pixel_to_angle, 1000, 5, nest=True)
    rec("pixel_to_angle", r.pixel_to_angle, 1024, np.array([0.5]))
    rec("nest_to_ring", r.nest_to_ring, 1000, 5)
    rec("nest_to_ring", r.nest_to_ring, 1024, np.array([1, 2, -7, 12 * 1024 * 1024]))
    rec("ring_to_nest", r.ring_to_nest, 1024, 12 * 1024 * 1024)
    rec("ring_to_nest", r.ring_to_nest, 0, 1)
    rec("vector_to_pixel", r.vector_to_pixel, 1000, 1.0, 0.0, 0.0, nest=True)
    rec("vector_to_pixel", r.vector_to_pixel, 1024, np.zeros(2), np.zeros(3), np.zeros(2))
    rec("pixel_to_vector", r.pixel_to_vector, 1024, -3)
    rec("neighbors", r.neighbors, 1024, -3)
    rec("neighbors", r.neighbors, 1024, np.zeros((2, 2), dtype=np.int64))
    rec("neighbors", r.neighbors, np.full((2, 2), 1024), np.zeros(2, dtype=np.int64))
    rec("neighbors", r.neighbors, np.array([1024, 2048, 4096]), np.zeros(2, dtype=np.int64))
    rec("get_interpolation_weights", r.get_interpolation_weights, 1024, np.zeros((2, 2)), np.zeros((2, 2)))
    rec("get_interpolation_weights", r.get_interpolation_weights, 1024, np.zeros(2), 0.0)
    rec("get_interpolation_weights", r.get_interpolation_weights, 1024, 0.0, 91.0)
    rec("upgrade_pixels", hp.upgrade_pixels, 1024, np.array([5]), 512)
    rec("upgrade_pixels", hp.upgrade_pixels, 1000, np.array([5]), 2000)
    rec("upgrade_pixels", hp.upgrade_pixels, 1024, np.array([5]), 3000)
    rec("upgrade_pixels", hp.upgrade_pixels, 1024, np.array([12 * 1024 * 1024 - 1]), 2048)
    rec("upgrade_pixels", hp.upgrade_pixels, 1024, np.array([-1]), 2048)
    rec("upgrade_pixels", hp.upgrade_pixels, 1024.0, np.array([1]), 2048)
    rec("lonlat_to_thetaphi", hp.lonlat_to_thetaphi, np.array([0.0]), np.array([90.1]))
    rec("pixel_ranges_to_pixels", r.pixel_ranges_to_pixels, np.array([[5, 3]]))
    rec("pixel_ranges_to_pixels", r.pixel_ranges_to_pixels, np.array([1, 2, 3]))
    with open(os.path.join(HERE, "golden_errors.json"), "w") as f:
        json.dump(errs, f, indent=1)

    # scalar / shape behaviour (types recorded as strings)
    shapes = {
        "a2p_scalar": type(r.angle_to_pixel(1024, 1.0, 2.0)).__name__,
        "a2p_2d": list(r.angle_to_pixel(1024, np.zeros((3, 2)), np.zeros((3, 2))).shape),
        "a2p_bcast": list(r.angle_to_pixel(1024, np.zeros((3, 1)), np.zeros((1, 4))).shape),
        "p2a_scalar": type(r.pixel_to_angle(1024, 5)[0]).__name__,
        "nb_scalar": list(r.neighbors(1024, 5).shape),
        "nb_1d": list(r.neighbors(1024, np.arange(3)).shape),
        "nb_nside_vec": r.neighbors(np.array([1024, 2048, 4096]), 1000).tolist(),
        "interp_scalar": [list(v.shape) for v in r.get_interpolation_weights(1024, 1.0, 2.0)],
        "empty": list(r.angle_to_pixel(1024, np.zeros(0), np.zeros(0)).shape),
        "a2p_nside_vec": r.angle_to_pixel(np.array([1, 2, 4, 8]), 10.0, 20.0).tolist(),
        "p2a_int32": [float(v) for v in r.pixel_to_angle(32, np.int32(7))],
    }
    with open(os.path.join(HERE, "golden_shapes.json"), "w") as f:
        json.dump(shapes, f, indent=1)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
