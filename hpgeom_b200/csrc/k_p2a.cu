// k_p2a.cu -- pixel_to_angle / pixel_to_vector kernels
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
// pixel_to_angle   (reference loop: hpgeom/hpgeom.c:426-457)   8 B in + 16 B out per pixel
// -------------------------------------------------------------------------------------------
// acos / atan Taylor tables of hpgb_math.h (72 + 6.4 KB), staged into dynamic shared memory per CTA
static __device__ const unsigned long long g_hpgb_acos_tab[HPGB_ACOS_ROWS * HPGB_ACOS_ROW_WORDS] = HPGB_ACOS_TABLE_INIT;
static __device__ const double g_hpgb_atanq_tab[HPGB_ATANQ_ROWS * HPGB_ATANQ_ROW_DOUBLES] = HPGB_ATANQ_TABLE_INIT;
// one big CTA per SM: the 78 KB of tables are staged once per SM and 24 warps share them
#ifndef HPGB_P2A_BLOCK
#define HPGB_P2A_BLOCK 768
#endif
// vector_to_pixel on the streaming skeleton: ring depth / sub-blocks per tile (three float64 inputs:
// 24 KiB per sub-block and stage)
#ifndef HPGB_P2V_BLOCK
#define HPGB_P2V_BLOCK 640
#endif
#ifndef HPGB_P2V_MINB
#define HPGB_P2V_MINB 2
#endif
#ifndef HPGB_P2V_ST
#define HPGB_P2V_ST 3
#endif
#ifndef HPGB_P2V_H
#define HPGB_P2V_H 2
#endif
#ifndef HPGB_V2P_MINB
#define HPGB_V2P_MINB 2
#endif
#ifndef HPGB_V2P_ST
#define HPGB_V2P_ST 2
#endif
#ifndef HPGB_V2P_H
#define HPGB_V2P_H 2
#endif
#define HPGB_P2A_TAB_DOUBLES (HPGB_ACOS_ROWS * HPGB_ACOS_ROW_WORDS + HPGB_ATANQ_ROWS_POLAR * HPGB_ATANQ_ROW_DOUBLES)
// streaming form: ONE big CTA per SM (the tables are staged once per SM), tiles of 4 * BLOCK pixels
#ifndef HPGB_P2A_ST
#define HPGB_P2A_ST 3
#endif
#ifndef HPGB_P2A_H
#define HPGB_P2A_H 1
#endif
#define HPGB_P2A_QCAP 64  // per-warp queue of deferred pixels: 31 left over + 32 new, rounded up

template <bool NEST, bool POW2, bool LONLAT, bool DEGREES, bool VAR>
struct OpPix2Ang {
    const int64_t *pix;
    double *a, *b;
    const int64_t *nside_v;
    int64_t base;
    hpgb_geom g;
    hpgb_rk c;  // POW2 && !VAR: constants of the branch-free pix2loc
    typedef longlong2 In2;
    typedef const uint64_t *Ctx;  // acos table (64-bit words); the atan table (doubles) follows it
    static constexpr int DYN_SMEM = HPGB_P2A_TAB_DOUBLES * 8;
    static constexpr bool FAST = POW2 && !VAR;
    static __device__ __forceinline__ const double *atanq(Ctx T) {
        return reinterpret_cast<const double *>(T + HPGB_ACOS_ROWS * HPGB_ACOS_ROW_WORDS);
    }
    // streaming skeleton: one pixel array in, the tables + the per-warp queues behind the stage ring
    static constexpr int NIN = 1;
    static constexpr int MINB = 1;
    static constexpr bool TILE_HOOK = false;
    static constexpr bool SMEM_PROLOGUE = true;
    static constexpr bool HAS_EPILOGUE = FAST;
    static constexpr int EXTRA_SMEM = DYN_SMEM + (HPGB_P2A_BLOCK / 32) * HPGB_P2A_QCAP * 8;
    __device__ __forceinline__ const void *src(int) const { return pix; }
    __device__ __forceinline__ In2 load2s(const unsigned char *stage, int, int e) const {
        return *reinterpret_cast<const longlong2 *>(stage + 8 * e);
    }
    static __device__ __forceinline__ Ctx stage_tables(unsigned long long *dst) {
        for (int i = threadIdx.x; i < HPGB_ACOS_ROWS * HPGB_ACOS_ROW_WORDS; i += blockDim.x) dst[i] = g_hpgb_acos_tab[i];
        for (int i = threadIdx.x; i < HPGB_ATANQ_ROWS_POLAR * HPGB_ATANQ_ROW_DOUBLES; i += blockDim.x)
            dst[HPGB_ACOS_ROWS * HPGB_ACOS_ROW_WORDS + i] = (unsigned long long)__double_as_longlong(g_hpgb_atanq_tab[i]);
        __syncthreads();
        return reinterpret_cast<const uint64_t *>(dst);
    }
    __device__ __forceinline__ Ctx prologue() const {
        extern __shared__ __align__(16) unsigned long long hpgb_p2a_tabs[];
        return stage_tables(hpgb_p2a_tabs);
    }
    __device__ __forceinline__ Ctx prologue_extra(unsigned char *extra) const {
        return stage_tables(reinterpret_cast<unsigned long long *>(extra));
    }
    // ---- deferred rare pixels (streaming kernel only).  1 % of the pixels have |z| > 0.99 and take theta from
    // atan2(sin, z); patched where they arise they cost a ~100-instruction pass on ONE lane in 72 % of the warp
    // iterations.  Instead their indices go to a per-warp queue behind the tables (the output keeps the hot
    // path's value for now) and are redone 32 at a time -- dense lanes -- by the reference-shaped routine, which
    // also covers the two rarer cases (isqrt quirk, pixel out of range).
    static __device__ __forceinline__ int64_t *queue(Ctx T) {
        return reinterpret_cast<int64_t *>(const_cast<uint64_t *>(T) + HPGB_P2A_TAB_DOUBLES) + (threadIdx.x >> 5) * HPGB_P2A_QCAP;
    }
    __device__ __forceinline__ void drain(int64_t *q, int first, int count, Ctx T, hpgb_bad_t &bad) const {
        const int lane = threadIdx.x & 31;
        if (lane < count) {
            const int64_t i = q[first + lane];
            double va, vb;
            conv(g, hpgb_ld1(pix + i), i, T, bad, va, vb);
            hpgb_st1(a + i, va);
            hpgb_st1(b + i, vb);
        }
        __syncwarp();
    }
    __device__ __forceinline__ void run4u(int64_t i0, const In2 &v0, int64_t i1, const In2 &v1, const Ctx &T,
                                          hpgb_bad_t &bad) const {
        Hot h0 = hot(v0.x, T), h1 = hot(v0.y, T), h2 = hot(v1.x, T), h3 = hot(v1.y, T);
        hpgb_st2(a + i0, h0.va, h1.va);
        hpgb_st2(b + i0, h0.vb, h1.vb);
        hpgb_st2(a + i1, h2.va, h3.va);
        hpgb_st2(b + i1, h2.vb, h3.vb);
        unsigned pend = (h0.code ? 1u : 0u) | (h1.code ? 2u : 0u) | (h2.code ? 4u : 0u) | (h3.code ? 8u : 0u);
        if (__any_sync(0xFFFFFFFFu, pend != 0u)) {  // warp-uniform
            __syncwarp();  // the stores above precede the queued pixels' final stores (other lanes, same addresses)
            int64_t *q = queue(T);
            int cnt = bad.aux;
            const unsigned lt = (1u << (threadIdx.x & 31)) - 1u;
            unsigned m;
            while ((m = __ballot_sync(0xFFFFFFFFu, pend != 0u)) != 0u) {  // one round per flagged pixel of the busiest lane
                if (pend) {
                    const unsigned j = (unsigned)__ffs((int)pend) - 1u;
                    pend &= pend - 1u;
                    q[cnt + __popc(m & lt)] = (j < 2u ? i0 : i1) + (j & 1u);
                }
                cnt += __popc(m);
                __syncwarp();
                if (cnt >= 32) {
                    cnt -= 32;
                    drain(q, cnt, 32, T, bad);
                }
            }
            bad.aux = cnt;
        }
    }
    __device__ __forceinline__ void epilogue(const Ctx &T, hpgb_bad_t &bad) const {
        __syncwarp();
        if (bad.aux > 0) drain(queue(T), 0, bad.aux, T, bad);
        bad.aux = 0;
    }
    __device__ __forceinline__ In2 load2(int64_t i) const { return hpgb_ld2(pix + i); }
    // general path (any nside, per-element nside): the reference's branches as they are
    __device__ __forceinline__ void conv(const hpgb_geom &gg, int64_t p, int64_t i, Ctx T, hpgb_bad_t &bad, double &va,
                                         double &vb) const {
        if (!hpgb_pixel_ok(gg, p)) {
            bad.flag(base + i);
            va = vb = 0.0;
            return;
        }
        hpgb_pix2ang<NEST, POW2, LONLAT, DEGREES>(gg, T, atanq(T), p, va, vb);
    }
    // power-of-two nside: straight-line hot part for every pixel, then the rare cases (out of range,
    // |z| > 0.99, isqrt quirk) patched per element
    struct Hot {
        double va, vb, z, phi, tcap;
        int code;
    };
    __device__ __forceinline__ Hot hot(int64_t p, Ctx T) const {
        Hot h;
        h.code = hpgb_pix2ang_k<NEST, LONLAT, DEGREES>(c, g, T, p, h.va, h.vb, h.z, h.phi, h.tcap);
        if (!hpgb_pixel_ok(g, p)) h.code = 3;
        return h;
    }
    __device__ __forceinline__ void fix(Hot &h, int64_t p, int64_t i, Ctx T, hpgb_bad_t &bad) const {
        if (h.code == 1) {
            hpgb_pix2ang_polar<LONLAT, DEGREES>(atanq(T), h.z, h.phi, h.tcap, h.va, h.vb);
        } else if (h.code == 2) {
            hpgb_pix2ang<NEST, POW2, LONLAT, DEGREES>(g, T, atanq(T), p, h.va, h.vb);
        } else if (h.code == 3) {
            bad.flag(base + i);
            h.va = h.vb = 0.0;
        }
    }
    __device__ __forceinline__ void run2(int64_t i, const In2 &v, const Ctx &T, hpgb_bad_t &bad) const {
        if (FAST) {
            Hot h0 = hot(v.x, T), h1 = hot(v.y, T);
            if (h0.code | h1.code) {
                fix(h0, v.x, i, T, bad);
                fix(h1, v.y, i + 1, T, bad);
            }
            hpgb_st2(a + i, h0.va, h1.va);
            hpgb_st2(b + i, h0.vb, h1.vb);
        } else {
            double a0, b0, a1, b1;
            conv(g, v.x, i, T, bad, a0, b0);
            conv(g, v.y, i + 1, T, bad, a1, b1);
            hpgb_st2(a + i, a0, a1);
            hpgb_st2(b + i, b0, b1);
        }
    }
    __device__ __forceinline__ void run4(int64_t i0, const In2 &v0, int64_t i1, const In2 &v1, const Ctx &T,
                                         hpgb_bad_t &bad) const {
        if (FAST) {
            Hot h0 = hot(v0.x, T), h1 = hot(v0.y, T), h2 = hot(v1.x, T), h3 = hot(v1.y, T);
            // rare cases (1 % of the pixels are polar): one pixel per lane per pass, so a warp pays the
            // long patch once per pass instead of once per pixel slot
            unsigned pend = (h0.code ? 1u : 0u) | (h1.code ? 2u : 0u) | (h2.code ? 4u : 0u) | (h3.code ? 8u : 0u);
            while (pend) {
                const unsigned j = (unsigned)__ffs((int)pend) - 1u;
                pend &= pend - 1u;
                Hot h = j == 0u ? h0 : (j == 1u ? h1 : (j == 2u ? h2 : h3));
                const int64_t p = j == 0u ? v0.x : (j == 1u ? v0.y : (j == 2u ? v1.x : v1.y));
                fix(h, p, (j < 2u ? i0 : i1) + (j & 1u), T, bad);
                if (j == 0u) h0.va = h.va, h0.vb = h.vb;
                if (j == 1u) h1.va = h.va, h1.vb = h.vb;
                if (j == 2u) h2.va = h.va, h2.vb = h.vb;
                if (j == 3u) h3.va = h.va, h3.vb = h.vb;
            }
            hpgb_st2(a + i0, h0.va, h1.va);
            hpgb_st2(b + i0, h0.vb, h1.vb);
            hpgb_st2(a + i1, h2.va, h3.va);
            hpgb_st2(b + i1, h2.vb, h3.vb);
        } else {
            run2(i0, v0, T, bad);
            run2(i1, v1, T, bad);
        }
    }
    __device__ __forceinline__ void run1(int64_t i, const Ctx &T, hpgb_bad_t &bad) const {
        double va = 0.0, vb = 0.0;
        if (VAR) {
            hpgb_geom gg;
            if (!var_geom(nside_v, i, NEST, gg))
                bad.flag(base + i);
            else
                conv(gg, hpgb_ld1(pix + i), i, T, bad, va, vb);
        } else {
            conv(g, hpgb_ld1(pix + i), i, T, bad, va, vb);
        }
        hpgb_st1(a + i, va);
        hpgb_st1(b + i, vb);
    }
};

// ---- ring-table form (hpgb_fp.h "Ring tables"): nside = 2^k <= 2^HPGB_RINGTAB_MAX_ORDER and a batch of at least
// HPGB_RINGTAB_MIN_PER_RING pixels per ring.  k_ringtab_angle fills tab[ring] (one thread per ring, the per-pixel
// routine on the ring's first pixel), the streaming kernel decodes (ring, phi) and gathers tab[ring] from L2.
#ifndef HPGB_RINGTAB_MAX_ORDER
#define HPGB_RINGTAB_MAX_ORDER 20
#endif
#ifndef HPGB_RINGTAB_P2V_MAX_ORDER
#define HPGB_RINGTAB_P2V_MAX_ORDER 12
#endif
#ifndef HPGB_RINGTAB_MIN_PER_RING
#define HPGB_RINGTAB_MIN_PER_RING 8
#endif
#ifndef HPGB_P2AT_MINB
#define HPGB_P2AT_MINB 3
#endif
#ifndef HPGB_P2AT_ST
#define HPGB_P2AT_ST 3
#endif
#ifndef HPGB_P2AT_H
#define HPGB_P2AT_H 1  // small stage ring: the gathers want the L1 (measured: 228 vs 188 Gpix/s at nside 4096, 150 vs 138 at 2^20)
#endif
// gather from a ring table (A/B switch HPGB_TAB_LD: 0 = read-only path, 1 = L2 only, 2 = read-only, no L1 allocation)
#ifndef HPGB_TAB_LD
#define HPGB_TAB_LD 0
#endif
__device__ __forceinline__ double hpgb_tab_ld(const double *p) {
#if HPGB_TAB_LD == 1
    return __ldcg(p);
#elif HPGB_TAB_LD == 2
    double v;
    asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ double2 hpgb_tab_ld(const double2 *p) {
#if HPGB_TAB_LD == 1
    return __ldcg(p);
#elif HPGB_TAB_LD == 2
    double2 v;
    asm("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}
template <bool LONLAT, bool DEGREES>
__global__ void __launch_bounds__(HPGB_P2A_BLOCK) k_ringtab_angle(const hpgb_geom gr, double *tab, const int64_t nrings) {
    extern __shared__ __align__(16) unsigned long long hpgb_p2a_tabs[];
    typedef OpPix2Ang<false, true, LONLAT, DEGREES, false> Gen;
    const uint64_t *T = Gen::stage_tables(hpgb_p2a_tabs);
    for (int64_t ring = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ring < nrings; ring += (int64_t)gridDim.x * blockDim.x)
        tab[ring] = ring == 0 ? 0.0 : hpgb_ringtab_angle<LONLAT, DEGREES>(gr, T, Gen::atanq(T), ring);
}

template <bool NEST, bool LONLAT, bool DEGREES>
struct OpPix2AngTab {
    const int64_t *pix;
    double *a, *b;
    int64_t base;
    hpgb_geom g;
    hpgb_rk c;
    const double *tab;  // 4 nside entries, [ring]
    typedef longlong2 In2;
    typedef hpgb_no_ctx Ctx;
    HPGB_STREAM_IN_I64(pix, HPGB_P2AT_MINB)
    __device__ __forceinline__ Ctx prologue() const { return Ctx(); }
    __device__ __forceinline__ In2 load2(int64_t i) const { return hpgb_ld2(pix + i); }
    __device__ __forceinline__ void one(int64_t p, int64_t i, hpgb_bad_t &bad, double &va, double &vb) const {
        const bool okp = hpgb_pixel_ok(g, p);
        uint32_t ring;
        double ph;
        hpgb_pix2ang_tab<NEST, LONLAT, DEGREES>(c, g, okp ? p : 0, ring, ph);
        const double th = hpgb_tab_ld(tab + ring);
        va = LONLAT ? ph : th;
        vb = LONLAT ? th : ph;
        if (!okp) {
            bad.flag(base + i);
            va = vb = 0.0;
        }
    }
    __device__ __forceinline__ void run2(int64_t i, const In2 &v, const Ctx &, hpgb_bad_t &bad) const {
        double a0, b0, a1, b1;
        one(v.x, i, bad, a0, b0);
        one(v.y, i + 1, bad, a1, b1);
        hpgb_st2(a + i, a0, a1);
        hpgb_st2(b + i, b0, b1);
    }
    HPGB_DEFAULT_RUN4
    __device__ __forceinline__ void run1(int64_t i, const Ctx &, hpgb_bad_t &bad) const {
        double va, vb;
        one(hpgb_ld1(pix + i), i, bad, va, vb);
        hpgb_st1(a + i, va);
        hpgb_st1(b + i, vb);
    }
};

// does a batch of n pixels at this geometry take the ring-table form?
// (max_order: pixel_to_angle gains up to nside 2^20 -- 228 / 164 / 157 Gpix/s at nside 4096 / 2^16 / 2^20 against 140
// per pixel: beyond the L1 the gather costs one L2 sector per pixel --, pixel_to_vector only while the table stays in
// the L1, its sin / cos(phi) dominate either way)
static inline bool hpgb_use_ringtab(const hpgb_geom &g, int64_t n, int max_order) {
#if defined(HPGB_NO_RINGTAB)
    return false;
#else
    return g.order >= 0 && g.order <= max_order && n >= (int64_t)HPGB_RINGTAB_MIN_PER_RING * 4 * g.nside && n >= (1 << 20);
#endif
}

template <bool NEST, bool LONLAT, bool DEGREES>
int launch_p2a_ringtab(const int64_t *pix, double *a, double *b, int64_t n, const hpgb_geom &g, int64_t base,
                       hpgb_status_t *st, cudaStream_t s) {
    const int64_t nrings = 4 * g.nside;
    double *tab = nullptr;
    cudaError_t e = hpgb_dev_alloc_on((void **)&tab, (size_t)nrings * 8, s);
    if (e != cudaSuccess) return HPGB_E_CUDA + (int)e;
    typedef OpPix2Ang<false, true, LONLAT, DEGREES, false> Gen;
    static hpgb_occ_cache cache;
    hpgb_resident_ctas_cached(cache, k_ringtab_angle<LONLAT, DEGREES>, Gen::DYN_SMEM, HPGB_P2A_BLOCK);
    int64_t grid = (nrings + HPGB_P2A_BLOCK - 1) / HPGB_P2A_BLOCK;
    if (grid > hpgb_sm_count()) grid = hpgb_sm_count();
    hpgb_geom gr = g;
    gr.nest = 0;
    k_ringtab_angle<LONLAT, DEGREES><<<(unsigned)grid, HPGB_P2A_BLOCK, Gen::DYN_SMEM, s>>>(gr, tab, nrings);
    g_hpgb_launches.fetch_add(1, std::memory_order_relaxed);
    e = cudaGetLastError();
    int rc = e == cudaSuccess ? HPGB_OK : HPGB_E_CUDA + (int)e;
    if (rc == HPGB_OK) {
        OpPix2AngTab<NEST, LONLAT, DEGREES> op{pix, a, b, base, g, hpgb_make_rk(g), tab};
#if defined(HPGB_P2AT_MAP)
        rc = hpgb_launch_map(op, n, true, st, s);
#else
        rc = hpgb_launch_stream<decltype(op), HPGB_P2AT_ST, HPGB_P2AT_H>(op, n, st, s);
#endif
    }
    hpgb_dev_free_on(tab, s);
    return rc;
}

template <bool NEST, bool LONLAT, bool DEGREES>
int launch_p2a_t(const int64_t *pix, double *a, double *b, int64_t n, int64_t nside, const int64_t *nside_v,
                 int64_t base, hpgb_status_t *st, cudaStream_t s) {
    if (nside_v) {  // per-element nside: the general-nside code path handles powers of two as well
        OpPix2Ang<NEST, NEST, LONLAT, DEGREES, true> op{pix, a, b, nside_v, base, hpgb_geom(), hpgb_rk()};
        return hpgb_launch_map1(op, n, st, s, decltype(op)::DYN_SMEM);
    }
    const hpgb_geom g = hpgb_make_geom(nside, NEST);
    const bool vec = hpgb_aligned16(pix) && hpgb_aligned16(a) && hpgb_aligned16(b);
    if (NEST || g.order >= 0) {
        if (vec && hpgb_use_ringtab(g, n, HPGB_RINGTAB_MAX_ORDER)) return launch_p2a_ringtab<NEST, LONLAT, DEGREES>(pix, a, b, n, g, base, st, s);
        OpPix2Ang<NEST, true, LONLAT, DEGREES, false> op{pix, a, b, nullptr, base, g, hpgb_make_rk(g)};
#if !defined(HPGB_P2A_MAP)
        // one 768-thread CTA per SM on the TMA ring (3 stages x 3072 pixels), tables + queues behind the ring
        if (vec && n >= 4 * 4 * HPGB_P2A_BLOCK * HPGB_P2A_H)
            return hpgb_launch_stream<decltype(op), HPGB_P2A_ST, HPGB_P2A_H, HPGB_P2A_BLOCK, 1>(op, n, st, s);
#endif
        return hpgb_launch_map<decltype(op), HPGB_P2A_BLOCK>(op, n, vec, st, s, decltype(op)::DYN_SMEM);
    }
    OpPix2Ang<false, false, LONLAT, DEGREES, false> op{pix, a, b, nullptr, base, g, hpgb_rk()};
    return hpgb_launch_map(op, n, vec, st, s, decltype(op)::DYN_SMEM);
}

}  // namespace

int launch_p2a(const int64_t *pix, double *a, double *b, int64_t n, int64_t nside, const int64_t *nside_v, int nest,
               int lonlat, int degrees, int64_t base, hpgb_status_t *st, cudaStream_t s) {
    if (n < 0 || (n > 0 && (!pix || !a || !b || !st))) return HPGB_E_BAD_ARGUMENT;
    if (n == 0) return HPGB_OK;
    if (!nside_v) {
        int c = hpgb_check_nside(nside, nest);
        if (c) return c;
    }
#define GO(N, L, D) return launch_p2a_t<N, L, D>(pix, a, b, n, nside, nside_v, base, st, s)
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

// -------------------------------------------------------------------------------------------
// vector_to_pixel / pixel_to_vector   (hpgeom/hpgeom.c:2323-2345, :2617-2643)   32 B per element
// -------------------------------------------------------------------------------------------
// Fast path (hpgb_vector_to_pixel_fast_g) with the reference-order path as the rare, non-inlined
// fallback, exactly like angle_to_pixel.
template <bool NEST, bool VAR, bool HI>
struct OpVec2Pix {
    const double *x, *y, *z;
    int64_t *out;
    const int64_t *nside_v;
    int64_t base;
    hpgb_geom g;
    hpgb_a2p_fast f;
    hpgb_rk rk;  // RING only
    struct In2 {
        double2 x, y, z;
    };
    // per-CTA context: the sin / cos + atan rows, and (NEST, scalar nside) the 32-entry table of the folded face
    // (hpgb_sd_to_pix, hpgb_reorder.h "folded coordinates")
    struct Ctx {
        const double *T;
        const hpgb_u32x2 *flut;
    };
    static constexpr bool FOLD = NEST && !VAR;
    HPGB_STREAM_IN_F64x3(x, y, z, HPGB_V2P_MINB)
    __device__ __forceinline__ Ctx prologue() const {
        __shared__ hpgb_u32x2 s_flut[32];
        if (FOLD && threadIdx.x < 32) {
            hpgb_u32x2 e;
            hpgb_a2p_fold_entry(rk, threadIdx.x, e.x, e.y);
            s_flut[threadIdx.x] = e;
        }
        const double *T = hpgb_stage_table();  // ends with __syncthreads(): the table above is visible too
#if defined(HPGB_A2P_NO_FOLD)
        return Ctx{T, nullptr};
#else
        return Ctx{T, FOLD ? s_flut : nullptr};
#endif
    }
    __device__ __forceinline__ In2 load2(int64_t i) const { return In2{hpgb_ld2(x + i), hpgb_ld2(y + i), hpgb_ld2(z + i)}; }
    static __device__ __noinline__ int64_t slow(const hpgb_geom &gg, Ctx T, double vx, double vy, double vz) {
        return hpgb_vec2pix<NEST>(gg, T.T, vx, vy, vz);
    }
    __device__ __forceinline__ void run2(int64_t i, const In2 &v, const Ctx &T, hpgb_bad_t &) const {
        int64_t p0 = 0, p1 = 0;
        const uint32_t g0 = hpgb_vector_to_pixel_fast_g<NEST, HI>(g, f, rk, T.T, v.x.x, v.y.x, v.z.x, p0, T.flut);
        const uint32_t g1 = hpgb_vector_to_pixel_fast_g<NEST, HI>(g, f, rk, T.T, v.x.y, v.y.y, v.z.y, p1, T.flut);
        if ((g0 | g1) != 0u) {
            if (g0) p0 = slow(g, T, v.x.x, v.y.x, v.z.x);
            if (g1) p1 = slow(g, T, v.x.y, v.y.y, v.z.y);
        }
        hpgb_st2(out + i, p0, p1);
    }
    __device__ __forceinline__ void run4(int64_t i0, const In2 &v0, int64_t i1, const In2 &v1, const Ctx &T,
                                         hpgb_bad_t &) const {
        int64_t p0 = 0, p1 = 0, p2 = 0, p3 = 0;
        const uint32_t g0 = hpgb_vector_to_pixel_fast_g<NEST, HI>(g, f, rk, T.T, v0.x.x, v0.y.x, v0.z.x, p0, T.flut);
        const uint32_t g1 = hpgb_vector_to_pixel_fast_g<NEST, HI>(g, f, rk, T.T, v0.x.y, v0.y.y, v0.z.y, p1, T.flut);
        const uint32_t g2 = hpgb_vector_to_pixel_fast_g<NEST, HI>(g, f, rk, T.T, v1.x.x, v1.y.x, v1.z.x, p2, T.flut);
        const uint32_t g3 = hpgb_vector_to_pixel_fast_g<NEST, HI>(g, f, rk, T.T, v1.x.y, v1.y.y, v1.z.y, p3, T.flut);
        if ((g0 | g1 | g2 | g3) != 0u) {
            if (g0) p0 = slow(g, T, v0.x.x, v0.y.x, v0.z.x);
            if (g1) p1 = slow(g, T, v0.x.y, v0.y.y, v0.z.y);
            if (g2) p2 = slow(g, T, v1.x.x, v1.y.x, v1.z.x);
            if (g3) p3 = slow(g, T, v1.x.y, v1.y.y, v1.z.y);
        }
        hpgb_st2(out + i0, p0, p1);
        hpgb_st2(out + i1, p2, p3);
    }
    __device__ __forceinline__ void run1(int64_t i, const Ctx &T, hpgb_bad_t &bad) const {
        hpgb_geom gg = g;
        if (VAR && !var_geom(nside_v, i, NEST, gg)) {
            bad.flag(base + i);
            hpgb_st1(out + i, -1);
            return;
        }
        const double vx = hpgb_ld1(x + i), vy = hpgb_ld1(y + i), vz = hpgb_ld1(z + i);
        int64_t pix = 0;
        if (VAR || hpgb_vector_to_pixel_fast_g<NEST, HI>(gg, f, rk, T.T, vx, vy, vz, pix, T.flut) != 0u) pix = slow(gg, T, vx, vy, vz);
        hpgb_st1(out + i, pix);
    }
};

template <bool NEST>
int launch_v2p_t(const double *x, const double *y, const double *z, int64_t *out, int64_t n, int64_t nside,
                 const int64_t *nside_v, int64_t base, hpgb_status_t *st, cudaStream_t s) {
    if (nside_v) {
        OpVec2Pix<NEST, true, false> op{x, y, z, out, nside_v, base, hpgb_geom(), hpgb_a2p_fast(), hpgb_rk()};
        return hpgb_launch_map1(op, n, st, s);
    }
    const hpgb_geom g = hpgb_make_geom(nside, NEST);
    const bool vec = hpgb_aligned16(x) && hpgb_aligned16(y) && hpgb_aligned16(z) && hpgb_aligned16(out);
    if (NEST && g.order >= 16) {
        OpVec2Pix<NEST, false, NEST> op{x, y, z, out, nullptr, base, g, hpgb_make_a2p_fast<true, true>(g), hpgb_make_rk(g)};
        return hpgb_launch_auto<decltype(op), HPGB_V2P_ST, HPGB_V2P_H>(op, n, vec, st, s);
    }
    OpVec2Pix<NEST, false, false> op{x, y, z, out, nullptr, base, g, hpgb_make_a2p_fast<true, true>(g), hpgb_make_rk(g)};
    return hpgb_launch_auto<decltype(op), HPGB_V2P_ST, HPGB_V2P_H>(op, n, vec, st, s);
}

}  // namespace

int launch_v2p(const double *x, const double *y, const double *z, int64_t *out, int64_t n, int64_t nside,
               const int64_t *nside_v, int nest, int64_t base, hpgb_status_t *st, cudaStream_t s) {
    if (n < 0 || (n > 0 && (!x || !y || !z || !out || !st))) return HPGB_E_BAD_ARGUMENT;
    if (n == 0) return HPGB_OK;
    if (!nside_v) {
        int c = hpgb_check_nside(nside, nest);
        if (c) return c;
    }
    return nest ? launch_v2p_t<true>(x, y, z, out, n, nside, nside_v, base, st, s)
                : launch_v2p_t<false>(x, y, z, out, n, nside, nside_v, base, st, s);
}

namespace {

template <bool NEST, bool POW2, bool VAR>
struct OpPix2Vec {
    const int64_t *pix;
    double *x, *y, *z;
    const int64_t *nside_v;
    int64_t base;
    hpgb_geom g;
    hpgb_rk c;  // POW2 && !VAR
    typedef longlong2 In2;
    typedef const double *Ctx;
    static constexpr bool FAST = POW2 && !VAR;
    HPGB_STREAM_IN_I64(pix, HPGB_P2V_MINB)
    __device__ __forceinline__ Ctx prologue() const { return hpgb_stage_table(); }
    __device__ __forceinline__ In2 load2(int64_t i) const { return hpgb_ld2(pix + i); }
    // general path: the reference's branches as they are
    __device__ __forceinline__ void conv(const hpgb_geom &gg, int64_t p, int64_t i, Ctx T, hpgb_bad_t &bad, double &vx,
                                         double &vy, double &vz) const {
        if (!hpgb_pixel_ok(gg, p)) {
            bad.flag(base + i);
            vx = vy = vz = 0.0;
            return;
        }
        hpgb_pix2vec<NEST, POW2>(gg, T, p, vx, vy, vz);
    }
    // power-of-two nside: straight-line hot part, rare cases (out of range, isqrt quirk) patched afterwards
    struct Hot {
        double vx, vy, vz;
        bool rare;
    };
    __device__ __forceinline__ Hot hot(int64_t p, Ctx T) const {
        Hot h;
        h.rare = !hpgb_pix2vec_k<NEST>(c, g, T, p, h.vx, h.vy, h.vz);
        h.rare |= !hpgb_pixel_ok(g, p);
        return h;
    }
    __device__ __forceinline__ void run2(int64_t i, const In2 &v, const Ctx &T, hpgb_bad_t &bad) const {
        double x0, y0, z0, x1, y1, z1;
        if (FAST) {
            Hot h0 = hot(v.x, T), h1 = hot(v.y, T);
            if (h0.rare | h1.rare) {
                if (h0.rare) conv(g, v.x, i, T, bad, h0.vx, h0.vy, h0.vz);
                if (h1.rare) conv(g, v.y, i + 1, T, bad, h1.vx, h1.vy, h1.vz);
            }
            x0 = h0.vx, y0 = h0.vy, z0 = h0.vz, x1 = h1.vx, y1 = h1.vy, z1 = h1.vz;
        } else {
            conv(g, v.x, i, T, bad, x0, y0, z0);
            conv(g, v.y, i + 1, T, bad, x1, y1, z1);
        }
        hpgb_st2(x + i, x0, x1);
        hpgb_st2(y + i, y0, y1);
        hpgb_st2(z + i, z0, z1);
    }
    __device__ __forceinline__ void run4(int64_t i0, const In2 &v0, int64_t i1, const In2 &v1, const Ctx &T,
                                         hpgb_bad_t &bad) const {
        if (FAST) {
            Hot h0 = hot(v0.x, T), h1 = hot(v0.y, T), h2 = hot(v1.x, T), h3 = hot(v1.y, T);
            if (h0.rare | h1.rare | h2.rare | h3.rare) {
                if (h0.rare) conv(g, v0.x, i0, T, bad, h0.vx, h0.vy, h0.vz);
                if (h1.rare) conv(g, v0.y, i0 + 1, T, bad, h1.vx, h1.vy, h1.vz);
                if (h2.rare) conv(g, v1.x, i1, T, bad, h2.vx, h2.vy, h2.vz);
                if (h3.rare) conv(g, v1.y, i1 + 1, T, bad, h3.vx, h3.vy, h3.vz);
            }
            hpgb_st2(x + i0, h0.vx, h1.vx);
            hpgb_st2(y + i0, h0.vy, h1.vy);
            hpgb_st2(z + i0, h0.vz, h1.vz);
            hpgb_st2(x + i1, h2.vx, h3.vx);
            hpgb_st2(y + i1, h2.vy, h3.vy);
            hpgb_st2(z + i1, h2.vz, h3.vz);
        } else {
            run2(i0, v0, T, bad);
            run2(i1, v1, T, bad);
        }
    }
    __device__ __forceinline__ void run1(int64_t i, const Ctx &T, hpgb_bad_t &bad) const {
        double vx = 0.0, vy = 0.0, vz = 0.0;
        hpgb_geom gg = g;
        if (VAR && !var_geom(nside_v, i, NEST, gg))
            bad.flag(base + i);
        else
            conv(gg, hpgb_ld1(pix + i), i, T, bad, vx, vy, vz);
        hpgb_st1(x + i, vx);
        hpgb_st1(y + i, vy);
        hpgb_st1(z + i, vz);
    }
};

// ring-table form of pixel_to_vector (hpgb_fp.h "Ring tables"): tab[m] = {z, sin(theta)} of ring m <= 2 nside, the
// southern rings mirror them exactly; cos / sin(phi) per pixel as before
__global__ void __launch_bounds__(256) k_ringtab_zsth(const hpgb_geom gr, double2 *tab, const int64_t nrings) {
    for (int64_t ring = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ring < nrings; ring += (int64_t)gridDim.x * blockDim.x) {
        double z = 0.0, sth = 0.0;
        if (ring > 0) hpgb_ringtab_zsth(gr, ring, z, sth);
        tab[ring] = make_double2(z, sth);
    }
}
#ifndef HPGB_P2VT_MINB
#define HPGB_P2VT_MINB 2
#endif
#ifndef HPGB_P2VT_ST
#define HPGB_P2VT_ST 3
#endif
#ifndef HPGB_P2VT_H
#define HPGB_P2VT_H 2
#endif
template <bool NEST>
struct OpPix2VecTab {
    const int64_t *pix;
    double *x, *y, *z;
    int64_t base;
    hpgb_geom g;
    hpgb_rk c;
    const double2 *tab;  // 2 nside + 1 entries
    typedef longlong2 In2;
    typedef const double *Ctx;
    HPGB_STREAM_IN_I64(pix, HPGB_P2VT_MINB)
    __device__ __forceinline__ Ctx prologue() const { return hpgb_stage_table(); }
    __device__ __forceinline__ In2 load2(int64_t i) const { return hpgb_ld2(pix + i); }
    __device__ __forceinline__ void one(int64_t p, int64_t i, Ctx T, hpgb_bad_t &bad, double &vx, double &vy, double &vz) const {
        const bool okp = hpgb_pixel_ok(g, p);
        uint32_t ring;
        bool north;
        double phi, s, co;
        hpgb_pix2ringphi_k<NEST>(c, g, okp ? p : 0, ring, north, phi);
        const bool upper = ring <= c.ns2;
        const double2 zs = hpgb_tab_ld(tab + (upper ? ring : c.ns4 - ring));
#if defined(HPGB_P2V_SINCOS_SD)  // A/B switch, see hpgb_pix2vec_k
        hpgb_sincos_sd(T, phi, s, co);
#else
        hpgb_sincos_p(T, phi, s, co);
#endif
        vx = zs.y * co;
        vy = zs.y * s;
        vz = upper ? zs.x : -zs.x;
        if (!okp) {
            bad.flag(base + i);
            vx = vy = vz = 0.0;
        }
    }
    __device__ __forceinline__ void run2(int64_t i, const In2 &v, const Ctx &T, hpgb_bad_t &bad) const {
        double x0, y0, z0, x1, y1, z1;
        one(v.x, i, T, bad, x0, y0, z0);
        one(v.y, i + 1, T, bad, x1, y1, z1);
        hpgb_st2(x + i, x0, x1);
        hpgb_st2(y + i, y0, y1);
        hpgb_st2(z + i, z0, z1);
    }
    HPGB_DEFAULT_RUN4
    __device__ __forceinline__ void run1(int64_t i, const Ctx &T, hpgb_bad_t &bad) const {
        double vx, vy, vz;
        one(hpgb_ld1(pix + i), i, T, bad, vx, vy, vz);
        hpgb_st1(x + i, vx);
        hpgb_st1(y + i, vy);
        hpgb_st1(z + i, vz);
    }
};

template <bool NEST>
int launch_p2v_ringtab(const int64_t *pix, double *x, double *y, double *z, int64_t n, const hpgb_geom &g, int64_t base,
                       hpgb_status_t *st, cudaStream_t s) {
    const int64_t nrings = 2 * g.nside + 1;
    double2 *tab = nullptr;
    cudaError_t e = hpgb_dev_alloc_on((void **)&tab, (size_t)nrings * 16, s);
    if (e != cudaSuccess) return HPGB_E_CUDA + (int)e;
    int64_t grid = (nrings + 255) / 256;
    if (grid > 8 * (int64_t)hpgb_sm_count()) grid = 8 * (int64_t)hpgb_sm_count();
    hpgb_geom gr = g;
    gr.nest = 0;
    k_ringtab_zsth<<<(unsigned)grid, 256, 0, s>>>(gr, tab, nrings);
    g_hpgb_launches.fetch_add(1, std::memory_order_relaxed);
    e = cudaGetLastError();
    int rc = e == cudaSuccess ? HPGB_OK : HPGB_E_CUDA + (int)e;
    if (rc == HPGB_OK) {
        OpPix2VecTab<NEST> op{pix, x, y, z, base, g, hpgb_make_rk(g), tab};
#if defined(HPGB_P2AT_MAP)
        rc = hpgb_launch_map(op, n, true, st, s);
#else
        rc = hpgb_launch_stream<decltype(op), HPGB_P2VT_ST, HPGB_P2VT_H>(op, n, st, s);
#endif
    }
    hpgb_dev_free_on(tab, s);
    return rc;
}

}  // namespace

int launch_p2v(const int64_t *pix, double *x, double *y, double *z, int64_t n, int64_t nside, const int64_t *nside_v,
               int nest, int64_t base, hpgb_status_t *st, cudaStream_t s) {
    if (n < 0 || (n > 0 && (!pix || !x || !y || !z || !st))) return HPGB_E_BAD_ARGUMENT;
    if (n == 0) return HPGB_OK;
    if (nside_v) {
        if (nest) {
            OpPix2Vec<true, true, true> op{pix, x, y, z, nside_v, base, hpgb_geom(), hpgb_rk()};
            return hpgb_launch_map1(op, n, st, s);
        }
        OpPix2Vec<false, false, true> op{pix, x, y, z, nside_v, base, hpgb_geom(), hpgb_rk()};
        return hpgb_launch_map1(op, n, st, s);
    }
    int c = hpgb_check_nside(nside, nest);
    if (c) return c;
    const hpgb_geom g = hpgb_make_geom(nside, nest);
    const bool vec = hpgb_aligned16(pix) && hpgb_aligned16(x) && hpgb_aligned16(y) && hpgb_aligned16(z);
    // power-of-two nside: the TMA-fed streaming skeleton (the pixel tiles arrive through the shared-memory ring, so
    // load latency is hidden without occupancy and the ~95 registers of the fp64 body are affordable)
    if (vec && hpgb_use_ringtab(g, n, HPGB_RINGTAB_P2V_MAX_ORDER))
        return nest ? launch_p2v_ringtab<true>(pix, x, y, z, n, g, base, st, s)
                    : launch_p2v_ringtab<false>(pix, x, y, z, n, g, base, st, s);
    if (nest) {
        OpPix2Vec<true, true, false> op{pix, x, y, z, nullptr, base, g, hpgb_make_rk(g)};
#if defined(HPGB_P2V_MAP)
        return hpgb_launch_map<decltype(op), HPGB_P2V_BLOCK>(op, n, vec, st, s);
#else
        return hpgb_launch_auto<decltype(op), HPGB_P2V_ST, HPGB_P2V_H>(op, n, vec, st, s);
#endif
    }
    if (g.order >= 0) {
        OpPix2Vec<false, true, false> op{pix, x, y, z, nullptr, base, g, hpgb_make_rk(g)};
#if defined(HPGB_P2V_MAP)
        return hpgb_launch_map<decltype(op), HPGB_P2V_BLOCK>(op, n, vec, st, s);
#else
        return hpgb_launch_auto<decltype(op), HPGB_P2V_ST, HPGB_P2V_H>(op, n, vec, st, s);
#endif
    }
    OpPix2Vec<false, false, false> op{pix, x, y, z, nullptr, base, g, hpgb_rk()};
    return hpgb_launch_map(op, n, vec, st, s);
}

