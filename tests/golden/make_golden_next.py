#!/usr/bin/env python
"""Golden vectors for the rows next to the hot path: boundaries, max_pixel_radius,
the numpy helpers of hpgeom.py and the healpy_compat shims -- generated from the UNMODIFIED
reference (its compiled C in oracle/_ref, its own hpgeom.py / healpy_compat.py executed from
/root/reference).  Run in the authoring container only:

    make -C oracle ref && python tests/golden/make_golden_next.py
"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402


def load_reference():
    pkg = types.ModuleType("hpgeom")
    pkg.__path__ = []
    sys.modules["hpgeom"] = pkg
    sys.modules["hpgeom._hpgeom"] = oracle.ref
    mods = {}
    for name in ("hpgeom", "healpy_compat"):
        path = "/root/reference/hpgeom/%s.py" % name
        mod = types.ModuleType("hpgeom." + name)
        mod.__package__ = "hpgeom"
        sys.modules["hpgeom." + name] = mod
        with open(path) as f:
            exec(compile(f.read(), path, "exec"), mod.__dict__)
        setattr(pkg, name, mod)
        mods[name] = mod
    return mods["hpgeom"], mods["healpy_compat"]


def main():
    ref, hp = load_reference()
    rng = np.random.default_rng(2468)
    out = {}
    # ---- boundaries ----
    for nside in (1, 2, 32, 1024, 2 ** 20, 2 ** 29, 3, 924):
        npix = 12 * nside * nside
        pix = np.unique(np.concatenate([rng.integers(0, npix, 200), np.arange(min(npix, 12)), npix - 1 - np.arange(min(npix, 12))]))
        out["bnd_pix_%d" % nside] = pix
        for nest in ([True, False] if nside & (nside - 1) == 0 else [False]):
            for step in (1, 2, 5):
                for key, kw in (("deg", {}), ("rad", {"degrees": False}), ("tp", {"lonlat": False})):
                    a, b = ref.boundaries(nside, pix, step=step, nest=nest, **kw)
                    tag = "bnd_%d_%s_%d_%s" % (nside, "nest" if nest else "ring", step, key)
                    out[tag + "_a"], out[tag + "_b"] = a, b
    a, b = ref.boundaries(1024, 77, step=3)
    out["bnd_scalar_a"], out["bnd_scalar_b"] = a, b
    # ---- max_pixel_radius ----
    nsides = np.array([1, 2, 3, 4, 7, 32, 924, 1024, 4096, 100003, 2 ** 20, 2 ** 29 - 1, 2 ** 29])
    out["mpr_nside"] = nsides
    out["mpr_deg"] = ref.max_pixel_radius(nsides)
    out["mpr_rad"] = ref.max_pixel_radius(nsides, degrees=False)
    # ---- numpy helpers ----
    n = 500
    lon = rng.uniform(0, 360, n)
    lat = np.degrees(np.arcsin(rng.uniform(-1, 1, n)))
    theta, phi = ref.lonlat_to_thetaphi(lon, lat)
    out["h_lon"], out["h_lat"], out["h_theta"], out["h_phi"] = lon, lat, theta, phi
    out["h_tp2ll_deg_a"], out["h_tp2ll_deg_b"] = ref.thetaphi_to_lonlat(theta, phi)
    out["h_tp2ll_rad_a"], out["h_tp2ll_rad_b"] = ref.thetaphi_to_lonlat(theta, phi, degrees=False)
    out["h_a2v_ll"] = ref.angle_to_vector(lon, lat)
    out["h_a2v_tp"] = ref.angle_to_vector(theta, phi, lonlat=False)
    out["h_a2v_scalar"] = ref.angle_to_vector(10.0, 20.0)
    vec = rng.normal(size=(n, 3))
    out["h_vec"] = vec
    out["h_v2a_ll_a"], out["h_v2a_ll_b"] = ref.vector_to_angle(vec)
    out["h_v2a_tp_a"], out["h_v2a_tp_b"] = ref.vector_to_angle(vec, lonlat=False)
    out["h_area_deg"] = np.array([ref.nside_to_pixel_area(ns) for ns in (1, 2, 1024, 2 ** 29)])
    out["h_area_rad"] = np.array([ref.nside_to_pixel_area(ns, degrees=False) for ns in (1, 2, 1024, 2 ** 29)])
    out["h_res"] = np.array([[ref.nside_to_resolution(ns, units=u) for u in ("radians", "degrees", "arcminutes", "arcseconds")]
                             for ns in (1, 64, 2 ** 29)])
    out["h_order"] = ref.nside_to_order(np.array([1, 2, 1024, 2 ** 29]))
    out["h_npix2nside"] = ref.npixel_to_nside(12 * np.array([1, 4, 1024 ** 2, 2 ** 58]))
    for nside in (1, 4, 32):
        m = rng.normal(size=12 * nside * nside)
        out["reorder_in_%d" % nside] = m
        out["reorder_r2n_%d" % nside] = ref.reorder(m, ring_to_nest=True)
        out["reorder_n2r_%d" % nside] = ref.reorder(m, ring_to_nest=False)
    # ---- healpy_compat ----
    nside = 256
    pix = rng.integers(0, 12 * nside * nside, 300)
    out["hp_pix"] = pix
    for nest in (True, False):
        t = "nest" if nest else "ring"
        out["hp_ang2pix_" + t] = hp.ang2pix(nside, theta, phi, nest=nest)
        out["hp_ang2pix_ll_" + t] = hp.ang2pix(nside, lon, lat, nest=nest, lonlat=True)
        a, b = hp.pix2ang(nside, pix, nest=nest)
        out["hp_pix2ang_a_" + t], out["hp_pix2ang_b_" + t] = a, b
        a, b = hp.pix2ang(nside, pix, nest=nest, lonlat=True)
        out["hp_pix2ang_ll_a_" + t], out["hp_pix2ang_ll_b_" + t] = a, b
        x, y, z = hp.pix2vec(nside, pix, nest=nest)
        out["hp_pix2vec_" + t] = np.array([x, y, z])
        out["hp_vec2pix_" + t] = hp.vec2pix(nside, vec[:, 0], vec[:, 1], vec[:, 2], nest=nest)
        out["hp_neigh_pix_" + t] = hp.get_all_neighbours(nside, pix, nest=nest)
        out["hp_neigh_ang_" + t] = hp.get_all_neighbours(nside, theta, phi, nest=nest)
        out["hp_neigh_scalar_" + t] = hp.get_all_neighbours(nside, 100, nest=nest)
        p, w = hp.get_interp_weights(nside, theta, phi, nest=nest)
        out["hp_interp_p_" + t], out["hp_interp_w_" + t] = p, w
        p, w = hp.get_interp_weights(nside, pix, nest=nest)
        out["hp_interp_pix_p_" + t], out["hp_interp_pix_w_" + t] = p, w
        out["hp_bnd_" + t] = hp.boundaries(nside, pix[:20], step=2, nest=nest)
        out["hp_bnd_scalar_" + t] = hp.boundaries(nside, 5, step=1, nest=nest)
    out["hp_nest2ring"] = hp.nest2ring(nside, pix)
    out["hp_ring2nest"] = hp.ring2nest(nside, pix)
    out["hp_ang2vec"] = hp.ang2vec(theta, phi)
    out["hp_ang2vec_ll"] = hp.ang2vec(lon, lat, lonlat=True)
    a, b = hp.vec2ang(vec)
    out["hp_vec2ang_a"], out["hp_vec2ang_b"] = a, b
    out["hp_scalars"] = np.array([hp.nside2npix(64), hp.npix2nside(12 * 64 * 64), hp.nside2pixarea(64), hp.nside2pixarea(64, degrees=True),
                                  hp.nside2resol(64), hp.nside2resol(64, arcmin=True), hp.nside2order(64), hp.order2nside(6),
                                  hp.max_pixrad(64), hp.max_pixrad(64, degrees=True)], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "golden_next.npz"), **out)
    # error texts
    errs = {}

    def rec(name, fn):
        try:
            fn()
            errs[name] = None
        except Exception as e:  # noqa: BLE001
            errs[name] = [type(e).__name__, str(e)]

    rec("bnd_step0", lambda: ref.boundaries(32, 1, step=0))
    rec("bnd_2d", lambda: ref.boundaries(32, np.zeros((2, 2), dtype=np.int64)))
    rec("bnd_badpix", lambda: ref.boundaries(32, 12 * 32 * 32))
    rec("bnd_nest_nonpow2", lambda: ref.boundaries(30, 1))
    rec("mpr_bad", lambda: ref.max_pixel_radius(0))
    rec("mpr_big", lambda: ref.max_pixel_radius(2 ** 30))
    rec("order_bad", lambda: ref.nside_to_order(3))
    rec("o2n_bad", lambda: ref.order_to_nside(30))
    rec("npix_bad", lambda: ref.npixel_to_nside(13))
    rec("npix_neg", lambda: ref.npixel_to_nside(-12))
    rec("res_units", lambda: ref.nside_to_resolution(4, units="furlongs"))
    rec("a2v_theta", lambda: ref.angle_to_vector(4.0, 1.0, lonlat=False))
    rec("a2v_phi", lambda: ref.angle_to_vector(1.0, 7.0, lonlat=False))
    rec("reorder_bad", lambda: ref.reorder(np.zeros(13)))
    with open(os.path.join(HERE, "golden_next_errors.json"), "w") as f:
        json.dump(errs, f, indent=1)
    print("wrote golden_next.npz (%d arrays), golden_next_errors.json" % len(out))


if __name__ == "__main__":
    main()
