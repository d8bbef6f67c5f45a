"""Parity REPORT for the floating-point outputs (BASELINE north_star: pixel_to_angle, pixel_to_vector and the
interpolation weights "within 2 ulp").  For nside in {1024, 4096, 2^20, 2^29}, NEST and RING, 2M random elements
each, against the reference's own C (oracle/_ref; the C restatement if that build is absent) this test measures and
writes to gpurun_out/parity_report.json (committed copy: profiles/parity_report.json):

  * the fraction of outputs that are bit-identical,
  * the largest difference of the quantity the libm call produces (theta for pixel_to_angle, cos / sin(phi) scaled
    by sin(theta) for pixel_to_vector) in ulps,
  * the largest difference of the FINAL value in its own ulps.

Outputs with no cancellation after the transcendental (theta, x, y) must be within 2 ulp of the reference.
Latitude (90 deg - theta) and the theta-weights of the interpolation subtract nearly equal angles: a 1-ulp difference
of a correctly-rounded-vs-glibc acos becomes thousands of ulps OF THE SMALL RESULT; for those the
bound is on theta -- <= 1 ulp of an angle of size pi feeding the subtraction -- and the report states both numbers.
"""
import json
import os

import numpy as np
import pytest

from conftest import ROOT, ulp_diff

pytestmark = pytest.mark.gpu

NSIDES = [1024, 4096, 2 ** 20, 2 ** 29]
N = 2_000_000


@pytest.fixture(scope="module")
def hpg():
    import hpgeom_b200
    return hpgeom_b200


@pytest.fixture(scope="module")
def ref_api():
    """The reference's functions with its keyword names; oracle/_ref when present (it travels to the GPU box)."""
    import oracle
    if oracle.ref is not None:
        return oracle.ref, "oracle/_ref (unmodified reference C, glibc libm)"
    return oracle.port(), "oracle port (C restatement, glibc libm)"


def stats(got, ref):
    big = np.abs(ref) > 1e-300
    d = ulp_diff(got[big], ref[big])
    return {"identical_fraction": float(np.mean(got == ref)), "max_ulp_of_value": float(d.max()) if d.size else 0.0}


def test_parity_report(hpg, ref_api):
    ref, ref_name = ref_api
    report = {"reference": ref_name, "n_per_case": N, "cases": {}}
    ulp_pi = float(np.spacing(np.pi))
    for nside in NSIDES:
        rng = np.random.default_rng(12345 + nside % 1000)
        pix = rng.integers(0, 12 * nside * nside, N)
        lon = 360.0 * rng.random(N)
        lat = np.degrees(np.arcsin(2.0 * rng.random(N) - 1.0))
        for nest in (True, False):
            case = {}
            # ---- pixel_to_angle: theta (the libm output), phi (pure IEEE), latitude (cancels) ----
            th, ph = hpg.pixel_to_angle(nside, pix, nest=nest, lonlat=False)
            rth, rph = ref.pixel_to_angle(nside, pix, nest=nest, lonlat=False)
            lo, la = hpg.pixel_to_angle(nside, pix, nest=nest)
            rlo, rla = ref.pixel_to_angle(nside, pix, nest=nest)
            c = {"theta": stats(th, rth), "phi": stats(ph, rph), "lon_deg": stats(lo, rlo), "lat_deg": stats(la, rla)}
            c["lat_deg"]["max_abs_diff_in_ulps_of_theta"] = float(np.abs(np.radians(la) - np.radians(rla)).max() / ulp_pi)
            case["pixel_to_angle"] = c
            assert c["theta"]["max_ulp_of_value"] <= 2.0 and c["phi"]["identical_fraction"] == 1.0
            assert c["lon_deg"]["identical_fraction"] == 1.0
            assert c["theta"]["identical_fraction"] >= 0.99 and c["lat_deg"]["identical_fraction"] >= 0.99
            assert np.abs(la - rla).max() <= 1.5 * ulp_pi * 180.0 / np.pi  # one ulp of theta, in degrees
            # ---- pixel_to_vector: z pure IEEE; x, y = sin(theta) * cos / sin(phi) ----
            x, y, z = hpg.pixel_to_vector(nside, pix, nest=nest)
            rx, ry, rz = ref.pixel_to_vector(nside, pix, nest=nest)
            c = {"x": stats(x, rx), "y": stats(y, ry), "z": stats(z, rz)}
            case["pixel_to_vector"] = c
            assert c["z"]["identical_fraction"] == 1.0
            assert c["x"]["max_ulp_of_value"] <= 2.0 and c["y"]["max_ulp_of_value"] <= 2.0
            assert c["x"]["identical_fraction"] >= 0.99 and c["y"]["identical_fraction"] >= 0.99
            # ---- interpolation: pixels exact away from ring boundaries; phi-weights pure IEEE; theta-weights cancel ----
            p, w = hpg.get_interpolation_weights(nside, lon, lat, nest=nest)
            rp, rw = ref.get_interpolation_weights(nside, lon, lat, nest=nest)
            same = (p == rp).all(axis=1)
            wd = np.abs(w[same] - rw[same])
            ring_spacing = np.pi / (4.0 * nside)  # smallest colatitude step between rings is ~ this (belt: ~2/(3 nside))
            c = {"pixel_rows_identical_fraction": float(np.mean(same)),
                 "weights": stats(w[same], rw[same]),
                 "weights_max_abs_diff": float(wd.max()),
                 "weights_max_abs_diff_in_units_of_ulp_theta_over_ring_spacing": float(wd.max() / (ulp_pi / ring_spacing))}
            case["get_interpolation_weights"] = c
            assert c["pixel_rows_identical_fraction"] >= 1.0 - 1e-5       # random points never sit on a ring
            assert c["weights"]["identical_fraction"] >= 0.99
            assert wd.max() <= 2.0 * ulp_pi / ring_spacing * 4.0          # <= ~2 ulp of a ring's theta, amplified by 1 / spacing
            report["cases"]["nside_%d_%s" % (nside, "nest" if nest else "ring")] = case
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "parity_report.json"), "w") as f:
        json.dump(report, f, indent=1, sort_keys=True)
