// k_interp.cu -- get_interpolation_weights kernels
// (one Op struct per API function, see hpgb_launch.cuh for the kernel skeleton they plug into;
// templated on the flags the reference tests per element -- nest / lonlat / degrees,
// hpgeom/hpgeom.c:143-153 -- so those branches are compile-time here.  A scalar instantiation
// (VAR) serves per-element nside arrays.)
#include <cuda_runtime.h>
#include <stdint.h>

#include "hpgb_core.h"
#include "hpgb_fp.h"
#include "hpgb_host.h"
#include "hpgb_internal.h"
#include "hpgb_launch.cuh"

namespace {

// per-element geometry for nside arrays; flags the element and returns false on a bad nside
// (the reference re-checks nside whenever it changes, hpgeom/hpgeom.c:133-140)
__device__ __forceinline__ bool var_geom(const int64_t *nside_v, int64_t i, int nest, hpgb_geom &g) {
    const int64_t ns = hpgb_ld1(nside_v + i);
    if (hpgb_check_nside(ns, nest)) return false;
    g = hpgb_make_geom(ns, nest);
    return true;
}

// -------------------------------------------------------------------------------------------
// get_interpolation_weights   (hpgeom/hpgeom.c:3491-3523)   16 B in + 64 B out per point
// -------------------------------------------------------------------------------------------
// acos / atan Taylor tables of hpgb_math.h (72 + 46 KB), staged into dynamic shared memory: ONE
// 512-thread CTA per SM shares them.  The sin/cos table of the reference-shaped fallback is static.
static __device__ const unsigned long long g_hpgb_acos_tab[HPGB_ACOS_ROWS * HPGB_ACOS_ROW_WORDS] = HPGB_ACOS_TABLE_INIT;
static __device__ const double g_hpgb_atanq_tab[HPGB_ATANQ_ROWS * HPGB_ATANQ_ROW_DOUBLES] = HPGB_ATANQ_TABLE_INIT;
#define HPGB_INTERP_TAB_WORDS (HPGB_ACOS_ROWS * HPGB_ACOS_ROW_WORDS + HPGB_ATANQ_ROWS * HPGB_ATANQ_ROW_DOUBLES)
#ifndef HPGB_INTERP_BLOCK
#define HPGB_INTERP_BLOCK 512
#endif

template <bool NEST, bool POW2, bool LONLAT, bool DEGREES, bool VAR>
struct OpInterp {
    const double *a, *b;
    int64_t *pix;
    double *wgt;
    const int64_t *nside_v;
    int64_t base;
    hpgb_geom g;
    hpgb_rk rk;       // FAST: RING -> NEST constants
    hpgb_a2p_fast f;  // FAST: polynomial of the cheap ring choice
    struct In2 {
        double2 a, b;
    };
    struct Ctx {
        const double *T;     // sin/cos rows (reference-shaped path)
        const uint64_t *TA;  // acos rows
        const double *TQ;    // atan rows
    };
    static constexpr bool FAST = POW2 && !VAR;
    static constexpr int DYN_SMEM = FAST ? HPGB_INTERP_TAB_WORDS * 8 : 0;
    HPGB_DEFAULT_RUN4
    __device__ __forceinline__ Ctx prologue() const {
        Ctx c{nullptr, nullptr, nullptr};
        if (FAST) {
            extern __shared__ __align__(16) unsigned long long hpgb_interp_tabs[];
            for (int i = threadIdx.x; i < HPGB_ACOS_ROWS * HPGB_ACOS_ROW_WORDS; i += blockDim.x) hpgb_interp_tabs[i] = g_hpgb_acos_tab[i];
            for (int i = threadIdx.x; i < HPGB_ATANQ_ROWS * HPGB_ATANQ_ROW_DOUBLES; i += blockDim.x)
                hpgb_interp_tabs[HPGB_ACOS_ROWS * HPGB_ACOS_ROW_WORDS + i] = (unsigned long long)__double_as_longlong(g_hpgb_atanq_tab[i]);
            c.TA = reinterpret_cast<const uint64_t *>(hpgb_interp_tabs);
            c.TQ = reinterpret_cast<const double *>(hpgb_interp_tabs + HPGB_ACOS_ROWS * HPGB_ACOS_ROW_WORDS);
        }
        __shared__ double sT[4 * HPGB_SINCOS_ROWS];
        for (int i = threadIdx.x; i < 4 * HPGB_SINCOS_ROWS; i += blockDim.x) sT[i] = g_hpgb_sincos[i];
        __syncthreads();
        c.T = sT;
        return c;
    }
    __device__ __forceinline__ In2 load2(int64_t i) const { return In2{hpgb_ld2(a + i), hpgb_ld2(b + i)}; }
    static __device__ __noinline__ void slow(const hpgb_geom &gg, const double *T, double theta, double phi, int64_t *p,
                                             double *w) {
        hpgb_interpol<NEST, POW2>(gg, T, theta, phi, p, w);
    }
    __device__ __forceinline__ void row(const hpgb_geom &gg, double va, double vb, int64_t i, const Ctx &c, hpgb_bad_t &bad) const {
        int64_t p[4] = {-1, -1, -1, -1};
        double w[4] = {0.0, 0.0, 0.0, 0.0};
        double theta, phi;
        if (!hpgb_angles_exact<LONLAT, DEGREES>(va, vb, theta, phi)) {
            bad.flag(base + i);
        } else {
            int64_t ir1;
            if (FAST && hpgb_ring_above_fast(gg, f, theta, ir1)) {
                const hpgb_mp_tab mp{c.TA, c.TQ};
                hpgb_interpol_rings<NEST, true>(gg, rk, mp, theta, phi, ir1, p, w);
            } else {
                slow(gg, c.T, theta, phi, p, w);
            }
        }
        hpgb_st2(pix + 4 * i, p[0], p[1]);
        hpgb_st2(pix + 4 * i + 2, p[2], p[3]);
        hpgb_st2(wgt + 4 * i, w[0], w[1]);
        hpgb_st2(wgt + 4 * i + 2, w[2], w[3]);
    }
    __device__ __forceinline__ void run2(int64_t i, const In2 &v, const Ctx &c, hpgb_bad_t &bad) const {
        row(g, v.a.x, v.b.x, i, c, bad);
        row(g, v.a.y, v.b.y, i + 1, c, bad);
    }
    __device__ __forceinline__ void run1(int64_t i, const Ctx &c, hpgb_bad_t &bad) const {
        hpgb_geom gg = g;
        if (VAR && !var_geom(nside_v, i, NEST, gg)) {
            bad.flag(base + i);
            for (int m = 0; m < 4; m++) {
                pix[4 * i + m] = -1;
                wgt[4 * i + m] = 0.0;
            }
            return;
        }
        row(gg, hpgb_ld1(a + i), hpgb_ld1(b + i), i, c, bad);
    }
};

template <bool NEST, bool LONLAT, bool DEGREES>
int launch_interp_t(const double *a, const double *b, int64_t *pix, double *wgt, int64_t n, int64_t nside,
                    const int64_t *nside_v, int64_t base, hpgb_status_t *st, cudaStream_t s) {
    if (nside_v) {
        OpInterp<NEST, NEST, LONLAT, DEGREES, true> op{a, b, pix, wgt, nside_v, base, hpgb_geom(), hpgb_rk(), hpgb_a2p_fast()};
        return hpgb_launch_map1(op, n, st, s);
    }
    const hpgb_geom g = hpgb_make_geom(nside, NEST);
    const bool vec = hpgb_aligned16(a) && hpgb_aligned16(b);
    if (NEST || g.order >= 0) {
        OpInterp<NEST, true, LONLAT, DEGREES, false> op{a, b, pix, wgt, nullptr, base, g, hpgb_make_rk(g), hpgb_make_a2p_fast<false, false>(g)};
        return hpgb_launch_map<decltype(op), HPGB_INTERP_BLOCK>(op, n, vec, st, s, decltype(op)::DYN_SMEM);
    }
    OpInterp<false, false, LONLAT, DEGREES, false> op{a, b, pix, wgt, nullptr, base, g, hpgb_rk(), hpgb_a2p_fast()};
    return hpgb_launch_map(op, n, vec, st, s);
}

}  // namespace

int launch_interp(const double *a, const double *b, int64_t *pix, double *wgt, int64_t n, int64_t nside,
                  const int64_t *nside_v, int nest, int lonlat, int degrees, int64_t base, hpgb_status_t *st,
                  cudaStream_t s) {
    if (n < 0 || (n > 0 && (!a || !b || !pix || !wgt || !st))) return HPGB_E_BAD_ARGUMENT;
    if (n == 0) return HPGB_OK;
    if (!hpgb_aligned16(pix) || !hpgb_aligned16(wgt)) return HPGB_E_BAD_ARGUMENT;
    if (!nside_v) {
        int c = hpgb_check_nside(nside, nest);
        if (c) return c;
    }
#define GO(N, L, D) return launch_interp_t<N, L, D>(a, b, pix, wgt, n, nside, nside_v, base, st, s)
    if (nest) {
        if (lonlat) {
            if (degrees) GO(true, true, true);
            GO(true, true, false);
        }
        GO(true, false, false);
    }
    if (lonlat) {
        if (degrees) GO(false, true, true);
        GO(false, true, false);
    }
    GO(false, false, false);
#undef GO
}

namespace {


}  // namespace
