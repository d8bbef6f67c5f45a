// k_bounds.cu -- boundaries and max_pixel_radius kernels (the callers next to
// the conversion path; same element-loop shape as neighbors).
//   boundaries       : hpgeom/healpix_geom.c:807-888 (xyf2loc, locToPtg, boundaries) behind
//                      hpgeom/hpgeom.c:1956-2278.  Output (n, 4*step) + (n, 4*step), row-major.
//   max_pixel_radius : hpgeom/healpix_geom.c:480-502 behind hpgeom/hpgeom.c:3202-3440.
// Every +,-,*,/,sqrt is the reference's, in its order; acos / atan2 are the table-driven correctly
// rounded versions of hpgb_math.h (a boundary point has |z| <= 0.99 -> acos, or |z| > 0.99 -> atan2(sth, z)
// with sth <= 0.1425 |z|: the same two domains as pixel_to_angle).
#include <cuda_runtime.h>
#include <stdint.h>

#include "hpgb_core.h"
#include "hpgb_fp.h"
#include "hpgb_host.h"
#include "hpgb_internal.h"
#include "hpgb_launch.cuh"

namespace {

static __device__ const unsigned long long g_hpgb_acos_tab[HPGB_ACOS_ROWS * HPGB_ACOS_ROW_WORDS] = HPGB_ACOS_TABLE_INIT;
static __device__ const double g_hpgb_atanq_tab[HPGB_ATANQ_ROWS * HPGB_ATANQ_ROW_DOUBLES] = HPGB_ATANQ_TABLE_INIT;  // first 80 rows staged
#define HPGB_BND_TAB_WORDS (HPGB_ACOS_ROWS * HPGB_ACOS_ROW_WORDS + HPGB_ATANQ_ROWS_POLAR * HPGB_ATANQ_ROW_DOUBLES)
#ifndef HPGB_BND_BLOCK
#define HPGB_BND_BLOCK 1024
#endif

// hpgeom/healpix_geom.c:807-841 + :857-865, then hpgeom/hpgeom_utils.c:145-165
template <bool LONLAT, bool DEGREES>
__device__ __forceinline__ void boundary_point(const uint64_t *TA, const double *TQ, double x, double y, uint32_t face,
                                               double &a, double &b) {
    const double jr = (double)(int)hpgb_jrll(face) - x - y;
    // the three zones of xyf2loc with ONE division: the numerator is selected, not the branch (a warp of
    // random pixels would otherwise run all three); same operations per zone as the reference
    const bool north = jr < 1, south = jr > 3;
    const double nr = north ? jr : (south ? 4 - jr : 1.0);
    const double q = hpgb_div_tame((north | south) ? nr * nr : (2 - jr) * 2., 3.);
    const double z = north ? 1 - q : (south ? q - 1 : q);
    double sth = 0.0;
    const bool have_sth = (north && z > 0.99) || (south && z < -0.99);
    if (have_sth) sth = sqrt(q * (2.0 - q));
    double tmp = (double)(int)hpgb_jpll(face) * nr + x - y;
    if (tmp < 0) tmp += 8;
    if (tmp >= 8) tmp -= 8;
    // nr >= 1e-15 here: tame operands, the branch-free division gives the IEEE quotient
    const double phi = (nr < 1e-15) ? 0 : hpgb_div_tame(0.5 * HPGB_HALFPI * tmp, nr < 1e-15 ? 1.0 : nr);
    const double theta = have_sth ? hpgb_atan2_tab(TQ, sth, z) : hpgb_acos_tab(TA, z);
    hpgb_thetaphi_out<LONLAT, DEGREES>(theta, phi, a, b);
}

// one thread per (pixel, i): the four edge points number i of that pixel
template <bool NEST, bool LONLAT, bool DEGREES>
__global__ void __launch_bounds__(HPGB_BND_BLOCK) k_boundaries(const int64_t *__restrict__ pix, double *__restrict__ oa,
                                                               double *__restrict__ ob, const int64_t n, const int64_t step,
                                                               const hpgb_geom g, const int64_t *__restrict__ nside_v,
                                                               const int64_t base, hpgb_status_t *st) {
    extern __shared__ __align__(16) unsigned long long tabs[];
    for (int i = threadIdx.x; i < HPGB_ACOS_ROWS * HPGB_ACOS_ROW_WORDS; i += blockDim.x) tabs[i] = g_hpgb_acos_tab[i];
    for (int i = threadIdx.x; i < HPGB_ATANQ_ROWS_POLAR * HPGB_ATANQ_ROW_DOUBLES; i += blockDim.x)
        tabs[HPGB_ACOS_ROWS * HPGB_ACOS_ROW_WORDS + i] = (unsigned long long)__double_as_longlong(g_hpgb_atanq_tab[i]);
    __syncthreads();
    const uint64_t *TA = reinterpret_cast<const uint64_t *>(tabs);
    const double *TQ = reinterpret_cast<const double *>(tabs + HPGB_ACOS_ROWS * HPGB_ACOS_ROW_WORDS);
    const int64_t total = n * step;
    const int64_t stride = (int64_t)gridDim.x * HPGB_BND_BLOCK;
    hpgb_bad_t bad;
    for (int64_t w = (int64_t)blockIdx.x * HPGB_BND_BLOCK + threadIdx.x; w < total; w += stride) {
        const int64_t e = step == 1 ? w : w / step, i = w - e * step;
        hpgb_geom gg = g;
        const int64_t p = __ldg(pix + e);
        double *ra = oa + e * 4 * step, *rb = ob + e * 4 * step;
        bool ok = true;
        if (nside_v) {
            const int64_t ns = __ldg(nside_v + e);
            ok = hpgb_check_nside(ns, NEST) == 0;
            if (ok) gg = hpgb_make_geom(ns, NEST);
        }
        if (!ok || !hpgb_pixel_ok(gg, p)) {
            bad.flag(base + e);
            for (int k = 0; k < 4; k++) ra[i + k * step] = rb[i + k * step] = 0.0;
            continue;
        }
        int32_t ix, iy;
        uint32_t face;
        if (NEST) {
            uint32_t ux, uy;
            hpgb_nest2xyf(gg, p, ux, uy, face);
            ix = (int32_t)ux;
            iy = (int32_t)uy;
        } else if (gg.order >= 0) {
            hpgb_ring2xyf<true>(gg, p, ix, iy, face);
        } else {
            hpgb_ring2xyf<false>(gg, p, ix, iy, face);
        }
        // hpgeom/healpix_geom.c:867-888
        const double dns = (double)gg.nside;
        // nside and step * nside are integers in [1, 2^40]: tame operands, IEEE quotients without the slow path
        const double dc = hpgb_div_tame(0.5, dns);
        const double xc = hpgb_div_tame((double)ix + 0.5, dns);
        const double yc = hpgb_div_tame((double)iy + 0.5, dns);
        const double d = hpgb_div_tame(1.0, (double)(step * gg.nside));
        const double id = (double)i * d;
        boundary_point<LONLAT, DEGREES>(TA, TQ, xc + dc - id, yc + dc, face, ra[i], rb[i]);
        boundary_point<LONLAT, DEGREES>(TA, TQ, xc - dc, yc + dc - id, face, ra[i + step], rb[i + step]);
        boundary_point<LONLAT, DEGREES>(TA, TQ, xc - dc + id, yc - dc, face, ra[i + 2 * step], rb[i + 2 * step]);
        boundary_point<LONLAT, DEGREES>(TA, TQ, xc + dc, yc - dc + id, face, ra[i + 3 * step], rb[i + 3 * step]);
    }
    bad.commit(st);
}

// hpgeom/healpix_geom.c:480-502; one thread per nside
__global__ void __launch_bounds__(HPGB_BLOCK) k_max_pixrad(const int64_t *__restrict__ nside, double *__restrict__ out,
                                                           const int64_t n, const int degrees, const int64_t base,
                                                           hpgb_status_t *st) {
    const double *T = hpgb_stage_table();
    hpgb_bad_t bad;
    for (int64_t i = (int64_t)blockIdx.x * HPGB_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * HPGB_BLOCK) {
        const int64_t ns = nside[i];
        if (hpgb_check_nside(ns, 0)) {
            bad.flag(base + i);
            out[i] = 0.0;
            continue;
        }
        const double dns = (double)ns;
        const double vaz = 2. / 3.;
        const double phi_a = HPGB_PI / (4. * dns);
        double sintheta = sqrt((1.0 - vaz) * (1.0 + vaz));
        double sa, ca;
        hpgb_sincos_cr(T, phi_a, sa, ca);
        const double vax = sintheta * ca, vay = sintheta * sa;
        double t1 = 1. - 1. / dns;
        t1 *= t1;
        const double vbz = 1. - t1 / 3.;
        sintheta = sqrt((1.0 - vbz) * (1.0 + vbz));
        const double vbx = sintheta, vby = 0.0;
        const double cx = vay * vbz - vaz * vby, cy = vaz * vbx - vax * vbz, cz = vax * vby - vay * vbx;
        const double vdot = vax * vbx + vay * vby + vaz * vbz;
        double r = hpgb_atan2_any(T, sqrt(cx * cx + cy * cy + cz * cz), vdot);
        if (degrees) r *= 180.0 / HPGB_PI;
        out[i] = r;
    }
    bad.commit(st);
}

}  // namespace

int launch_boundaries(const int64_t *pix, double *a, double *b, int64_t n, int64_t nside, const int64_t *nside_v,
                      int64_t step, int nest, int lonlat, int degrees, int64_t base, hpgb_status_t *st, cudaStream_t s) {
    if (n < 0 || step < 1 || (n > 0 && (!pix || !a || !b || !st))) return HPGB_E_BAD_ARGUMENT;
    if (n == 0) return HPGB_OK;
    hpgb_geom g = hpgb_geom();
    if (!nside_v) {
        int c = hpgb_check_nside(nside, nest);
        if (c) return c;
        g = hpgb_make_geom(nside, nest);
    }
    const int smem = HPGB_BND_TAB_WORDS * 8;
    const int64_t total = n * step;
    int64_t grid = hpgb_sm_count();
    const int64_t want = (total + HPGB_BND_BLOCK - 1) / HPGB_BND_BLOCK;
    if (want < grid) grid = want;
#define GO(N, L, D)                                                                                              \
    do {                                                                                                         \
        static hpgb_occ_cache cache; /* per device: the dynamic-smem opt-in is a per-device attribute */        \
        (void)hpgb_resident_ctas_cached(cache, k_boundaries<N, L, D>, smem, HPGB_BND_BLOCK);                     \
        k_boundaries<N, L, D><<<(unsigned)grid, HPGB_BND_BLOCK, smem, s>>>(pix, a, b, n, step, g, nside_v, base, st); \
    } while (0)
    if (nest) {
        if (lonlat) {
            if (degrees) GO(true, true, true); else GO(true, true, false);
        } else GO(true, false, false);
    } else {
        if (lonlat) {
            if (degrees) GO(false, true, true); else GO(false, true, false);
        } else GO(false, false, false);
    }
#undef GO
    g_hpgb_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? HPGB_OK : HPGB_E_CUDA + (int)e;
}

int launch_max_pixrad(const int64_t *nside, double *out, int64_t n, int degrees, int64_t base, hpgb_status_t *st,
                      cudaStream_t s) {
    if (n < 0 || (n > 0 && (!nside || !out || !st))) return HPGB_E_BAD_ARGUMENT;
    if (n == 0) return HPGB_OK;
    int64_t grid = (n + HPGB_BLOCK - 1) / HPGB_BLOCK;
    if (grid > hpgb_sm_count() * 4) grid = hpgb_sm_count() * 4;
    k_max_pixrad<<<(unsigned)grid, HPGB_BLOCK, 0, s>>>(nside, out, n, degrees, base, st);
    g_hpgb_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? HPGB_OK : HPGB_E_CUDA + (int)e;
}
