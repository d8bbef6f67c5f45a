"""Golden vectors for query_circle (RING and NEST) and pixel_ranges_to_pixels, generated from the UNMODIFIED
reference C (oracle/_ref, compiled from /root/reference by oracle/Makefile).

    python tests/golden/make_golden_query.py        # writes tests/golden/golden_query.npz

Run in the authoring container only (needs /root/reference); the fixture travels to the GPU box.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402


def main():
    ref = oracle.ref
    assert ref is not None, "build oracle/_ref first (make -C oracle ref)"
    rng = np.random.default_rng(97531)
    params, sizes, pix = [], [], []

    def add(nside, a, b, radius, fact, lonlat, degrees):
        out = ref.query_circle(nside, a, b, radius, inclusive=fact > 0, fact=max(fact, 1), nest=False,
                               lonlat=bool(lonlat), degrees=bool(degrees))
        if out.size > 60000:
            return
        params.append((nside, a, b, radius, fact, lonlat, degrees))
        sizes.append(out.size)
        pix.append(out)

    for nside in (1, 2, 3, 8, 32, 100, 1024, 4096, 2 ** 15, 2 ** 22, 2 ** 29):
        res = np.sqrt(4 * np.pi / (12. * nside * nside))  # pixel size, radians
        for it in range(14):
            theta = float(np.arccos(rng.uniform(-1, 1)))
            phi = float(rng.uniform(0, 2 * np.pi))
            if it == 0:
                theta = 0.0
            if it == 1:
                theta = np.pi
            if it == 2:
                theta, phi = 1e-4, 0.0
            if it == 3:
                phi = 2 * np.pi
            if it == 4:
                theta = np.pi / 2
            radius = float(res * 10 ** rng.uniform(-0.5, 1.7))
            if nside <= 8 and it % 3 == 0:
                radius = float(rng.uniform(0.5, 3.3))
            for fact in (0, 1, 2, 4, 3):
                if fact * nside > 2 ** 29:
                    continue
                add(nside, theta, phi, radius, fact, 0, 0)
            # lonlat degrees form of the same disc
            add(nside, float(np.degrees(phi)), float(90.0 - np.degrees(theta)), float(np.degrees(radius)), 0, 1, 1)
            add(nside, float(phi), float(np.pi / 2 - theta) if abs(np.pi / 2 - theta) <= np.pi / 2 else 0.0, radius, 2 if 2 * nside <= 2 ** 29 else 0, 1, 0)
    # ---- NEST scheme: pixel lists and the merged range lists (return_pixel_ranges=True) ----
    nparams, nsizes, npixl, nrsizes, nranges = [], [], [], [], []
    for nside in (1, 2, 8, 32, 128, 1024, 4096, 2 ** 15, 2 ** 22, 2 ** 29):
        res = np.sqrt(4 * np.pi / (12. * nside * nside))
        for it in range(12):
            theta = float(np.arccos(rng.uniform(-1, 1)))
            phi = float(rng.uniform(0, 2 * np.pi))
            if it == 0:
                theta = 0.0
            if it == 1:
                theta = np.pi
            if it == 2:
                theta, phi = np.pi / 2, 0.0
            radius = float(res * 10 ** rng.uniform(-0.5, 1.7))
            if nside <= 8 and it % 3 == 0:
                radius = float(rng.uniform(0.5, 3.3))
            for fact in (0, 1, 2, 4, 16):
                if fact * nside > 2 ** 29:
                    continue
                kw = dict(inclusive=fact > 0, fact=max(fact, 1), nest=True, lonlat=False)
                px = ref.query_circle(nside, theta, phi, radius, **kw)
                if px.size > 60000:
                    continue
                rg = ref.query_circle(nside, theta, phi, radius, return_pixel_ranges=True, **kw)
                nparams.append((nside, theta, phi, radius, fact))
                nsizes.append(px.size)
                npixl.append(px)
                nrsizes.append(rg.shape[0])
                nranges.append(rg.reshape(-1, 2))
    out = {
        "qn_params": np.array(nparams, dtype=np.float64),
        "qn_sizes": np.array(nsizes, dtype=np.int64),
        "qn_pix": np.concatenate(npixl).astype(np.int64),
        "qn_rsizes": np.array(nrsizes, dtype=np.int64),
        "qn_ranges": np.concatenate(nranges).astype(np.int64),
        "qc_params": np.array(params, dtype=np.float64),
        "qc_sizes": np.array(sizes, dtype=np.int64),
        "qc_pix": np.concatenate(pix).astype(np.int64),
    }
    # pixel_ranges_to_pixels: ragged ranges with empty rows, both `inclusive` settings
    lo = np.sort(rng.integers(0, 2 ** 40, 3000))
    ln = rng.integers(0, 40, 3000)
    ln[rng.integers(0, 3000, 600)] = 0
    ln[7] = 5000
    rr = np.stack([lo, lo + ln], axis=1).astype(np.int64)
    out["pr_ranges"] = rr
    out["pr_excl"] = ref.pixel_ranges_to_pixels(rr, inclusive=False)
    out["pr_incl"] = ref.pixel_ranges_to_pixels(rr, inclusive=True)
    np.savez_compressed(os.path.join(HERE, "golden_query.npz"), **out)
    print("nest cases:", len(nparams), "pixels:", out["qn_pix"].size, "ranges:", out["qn_ranges"].shape[0])
    print("cases:", len(params), "pixels:", out["qc_pix"].size, "pr:", out["pr_excl"].size, out["pr_incl"].size)


if __name__ == "__main__":
    main()
