// tests/hostsim/hostsim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the per-element device functions of hpgeom_b200/csrc (hpgb_core.h, hpgb_math.h,
// hpgb_fp.h -- all `__host__ __device__`) with g++ and exposes array loops over them, so the
// bit-level logic of the CUDA kernels can be checked against the oracle in the `-m "not gpu"`
// test tier (no GPU in the authoring container).  Built with -ffp-contract=off: fused
// operations happen only where the device code calls fma() explicitly, exactly as under
// nvcc -fmad=false.  The product (hpgeom_b200) never loads this library.
#include <stdint.h>

#include "../../hpgeom_b200/csrc/hpgb_core.h"
#include "../../hpgeom_b200/csrc/hpgb_reorder.h"
#if __has_include("../../hpgeom_b200/csrc/hpgb_fp.h")
#include "../../hpgeom_b200/csrc/hpgb_fp.h"
#define HAVE_FP 1
#endif

extern "C" {

int64_t sim_nest_to_ring(int64_t nside, const int64_t *in, int64_t n, int64_t *out) {
    hpgb_geom g = hpgb_make_geom(nside, 1);
    for (int64_t i = 0; i < n; i++) {
        if (!hpgb_pixel_ok(g, in[i])) return i;
        out[i] = hpgb_nest2ring(g, in[i]);
    }
    return -1;
}

int64_t sim_ring_to_nest(int64_t nside, const int64_t *in, int64_t n, int64_t *out) {
    hpgb_geom g = hpgb_make_geom(nside, 1);
    for (int64_t i = 0; i < n; i++) {
        if (!hpgb_pixel_ok(g, in[i])) return i;
        out[i] = hpgb_ring2nest(g, in[i]);
    }
    return -1;
}

// the minimum-instruction cores the scalar-nside kernels use (hpgb_reorder.h); hi = 1 selects the
// nside >= 2^16 instantiation (only valid for such nside), 0 the general power-of-two one
int64_t sim_nest_to_ring_k(int64_t nside, const int64_t *in, int64_t n, int64_t *out, int hi) {
    hpgb_geom g = hpgb_make_geom(nside, 1);
    hpgb_rk c = hpgb_make_rk(g);
    for (int64_t i = 0; i < n; i++) {
        if (!hpgb_pixel_ok(g, in[i])) return i;
        const uint32_t lo = (uint32_t)(uint64_t)in[i], h = (uint32_t)((uint64_t)in[i] >> 32);
        out[i] = hi ? hpgb_nest2ring_k<true>(c, lo, h) : hpgb_nest2ring_k<false>(c, lo, h);
    }
    return -1;
}

int64_t sim_ring_to_nest_k(int64_t nside, const int64_t *in, int64_t n, int64_t *out, int hi) {
    hpgb_geom g = hpgb_make_geom(nside, 1);
    hpgb_rk c = hpgb_make_rk(g);
    for (int64_t i = 0; i < n; i++) {
        if (!hpgb_pixel_ok(g, in[i])) return i;
        out[i] = hi ? hpgb_ring2nest_k<true>(c, g, in[i]) : hpgb_ring2nest_k<false>(c, g, in[i]);
    }
    return -1;
}

// the folded-coordinate form of the streaming kernel's tile routine (lookup table built exactly as the kernel's prologue does)
int64_t sim_ring_to_nest_fold(int64_t nside, const int64_t *in, int64_t n, int64_t *out, int hi) {
    hpgb_geom g = hpgb_make_geom(nside, 1);
    hpgb_rk c = hpgb_make_rk(g);
    hpgb_u32x2 lut[16];
    for (uint32_t i = 0; i < 16; i++) hpgb_r2n_lut_entry(c, i, lut[i].x, lut[i].y);
    for (int64_t i = 0; i < n; i++) {
        if (!hpgb_pixel_ok(g, in[i])) return i;
        out[i] = hi ? hpgb_ring2nest_fold<true>(c, g, lut, in[i]) : hpgb_ring2nest_fold<false>(c, g, lut, in[i]);
    }
    return -1;
}

// ring -> (ix,iy,face) -> ring through the general-nside path (non powers of two)
int64_t sim_ring_roundtrip_any(int64_t nside, const int64_t *in, int64_t n, int64_t *out) {
    hpgb_geom g = hpgb_make_geom(nside, 0);
    for (int64_t i = 0; i < n; i++) {
        int32_t ix, iy;
        uint32_t f;
        hpgb_ring2xyf_any(g, in[i], ix, iy, f);
        out[i] = hpgb_xyf2ring(g, ix, iy, f);
    }
    return -1;
}

}  // extern "C"

#ifdef HAVE_FP
#include "hostsim_fp.inc"
#endif

// upgrade_pixels, RING scheme, as k_upgrade<false, C> does it (hpgb_upgrade_ring_group)
extern "C" int64_t sim_upgrade_ring(int64_t nside, int64_t nside_up, const int64_t *pix, int64_t n, int group4, int64_t *out) {
    const hpgb_geom g = hpgb_make_geom(nside, 1), gu = hpgb_make_geom(nside_up, 1);
    const hpgb_rk c = hpgb_make_rk(g), cu = hpgb_make_rk(gu);
    const int64_t ratio = (nside_up / nside) * (nside_up / nside);
    for (int64_t i = 0; i < n; i++) {
        if (!hpgb_pixel_ok(g, pix[i])) return i;
        bool last;
        if (group4 && ratio >= 4) {
            for (int64_t f = 0; f < ratio; f += 4) hpgb_upgrade_ring_group<4>(c, cu, g, pix[i], (uint64_t)f, out + i * ratio + f, last);
        } else {
            for (int64_t f = 0; f < ratio; f++) hpgb_upgrade_ring_group<1>(c, cu, g, pix[i], (uint64_t)f, out + i * ratio + f, last);
        }
        if (last) return i;
    }
    return -1;
}
