// k_a2p.cu -- angle_to_pixel kernels (the headline path)
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

// ring depth / sub-blocks per tile / min resident CTAs of the streaming kernel (tunable per build)
#ifndef HPGB_A2P_ST
#define HPGB_A2P_ST 3
#endif
#ifndef HPGB_A2P_H
#define HPGB_A2P_H 2
#endif
#ifndef HPGB_A2P_MINB
#define HPGB_A2P_MINB 2
#endif
#ifndef HPGB_A2P_QBLOCK
#define HPGB_A2P_QBLOCK HPGB_BLOCK  // threads per CTA of the kernels that use the deferred-exact-path queue
#endif

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
// angle_to_pixel   (reference loop: hpgeom/hpgeom.c:126-157)   16 B in + 8 B out per point.
// Fast path (hpgb_fp.h) with the exact path as a rare, non-inlined fallback; RING with a
// non-power-of-two nside has no fast path (f.enabled = 0) and runs the exact path throughout.
// -------------------------------------------------------------------------------------------
template <bool NEST, bool LONLAT, bool DEGREES, bool VAR, bool HI>
struct OpAng2Pix {
    const double *a, *b;
    int64_t *out;
    const int64_t *nside_v;
    int64_t base;
    hpgb_geom g;
    hpgb_a2p_fast f;
    hpgb_rk rk;  // RING only
    struct In2 {
        double2 a, b;
    };
    // per-CTA context: the sin / cos table of the exact path, and (NEST, scalar nside) the 32-entry table that folds the
    // face into the edge-line coordinates (hpgb_sd_to_pix, hpgb_reorder.h "folded coordinates")
    struct Ctx {
        const double *T;
        const hpgb_u32x2 *flut;
    };
    static constexpr bool FOLD = NEST && !VAR;
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
    __device__ __forceinline__ In2 load2(int64_t i) const { return In2{hpgb_ld2(a + i), hpgb_ld2(b + i)}; }
    static constexpr int NIN = 2;
    static constexpr bool TILE_HOOK = false;
    static constexpr int EXTRA_SMEM = 0;
    static constexpr int MINB = HPGB_A2P_MINB;  // 2 CTAs/SM (<= 128 registers): measured best with 2048-point tiles
    __device__ __forceinline__ const void *src(int k) const { return k == 0 ? (const void *)a : (const void *)b; }
    __device__ __forceinline__ In2 load2s(const unsigned char *stage, int tile, int e) const {
        return In2{*reinterpret_cast<const double2 *>(stage + 8 * e),
                   *reinterpret_cast<const double2 *>(stage + 8 * tile + 8 * e)};
    }
    // exact path for one element (rare): non-inlined call, flags the element when it is invalid
    __device__ __forceinline__ int64_t slow(const hpgb_geom &gg, double va, double vb, int64_t i, Ctx T,
                                            hpgb_bad_t &bad) const {
        int64_t pix;
        if (!hpgb_angle_to_pixel_exact<NEST, LONLAT, DEGREES>(gg, T.T, va, vb, pix)) bad.flag(base + i);
        return pix;
    }
    __device__ __forceinline__ void run2(int64_t i, const In2 &v, const Ctx &T, hpgb_bad_t &bad) const {
        int64_t p0 = 0, p1 = 0;  // both fast paths first (independent -> the fp64 chains interleave), exact path after
        const uint32_t g0 = hpgb_angle_to_pixel_fast_g<NEST, LONLAT, DEGREES, HI>(g, f, rk, v.a.x, v.b.x, p0, T.flut);
        const uint32_t g1 = hpgb_angle_to_pixel_fast_g<NEST, LONLAT, DEGREES, HI>(g, f, rk, v.a.y, v.b.y, p1, T.flut);
        if (g0) p0 = slow(g, v.a.x, v.b.x, i, T, bad);
        if (g1) p1 = slow(g, v.a.y, v.b.y, i + 1, T, bad);
        hpgb_st2(out + i, p0, p1);
    }
#if defined(HPGB_A2P_QUEUE)
    // Deferred exact path.  ~2e-4 of the points need it; taken where it arises it runs on ONE lane while the other 31
    // wait (2.5 % of the warp iterations, ~2x their cost).  Instead a flagged point is parked in a per-warp queue in
    // shared memory -- its slot in `out` keeps the untrusted fast result for now -- and the exact path runs on 32
    // parked points at once (dense lanes); the remainder is drained when the CTA runs out of tiles (epilogue).
    // Ordering of the two stores to one address by different lanes of the warp: __syncwarp() in between.
    static constexpr bool HAS_EPILOGUE = true;
    static constexpr int QCAP = 64;  // 31 left over + 32 new, rounded up
    typedef int64_t QEntry;          // the element's index: its angles are read again from HBM when it is drained
    static __device__ __forceinline__ QEntry *queue() {
        __shared__ QEntry q[HPGB_A2P_QBLOCK / 32][QCAP];  // 512 B per warp
        return q[threadIdx.x >> 5];
    }
    __device__ __forceinline__ void drain(QEntry *q, int first, int count, Ctx T, hpgb_bad_t &bad) const {
        const int lane = threadIdx.x & 31;
        if (lane < count) {
            const int64_t i = q[first + lane];
            hpgb_st1(out + i, slow(g, hpgb_ld1(a + i), hpgb_ld1(b + i), i, T, bad));
        }
        __syncwarp();
    }
    __device__ __forceinline__ void park(uint32_t gj, int64_t i, int &cnt, Ctx T, hpgb_bad_t &bad) const {
        const unsigned m = __ballot_sync(0xFFFFFFFFu, gj != 0u);
        if (m == 0u) return;  // warp-uniform
        QEntry *q = queue();
        if (gj) q[cnt + __popc(m & ((1u << (threadIdx.x & 31)) - 1u))] = i;
        cnt += __popc(m);
        __syncwarp();
        if (cnt >= 32) {
            cnt -= 32;
            drain(q, cnt, 32, T, bad);
        }
    }
    __device__ __forceinline__ void run4u(int64_t i0, const In2 &v0, int64_t i1, const In2 &v1, const Ctx &T,
                                         hpgb_bad_t &bad) const {
        int64_t p0 = 0, p1 = 0, p2 = 0, p3 = 0;
        const uint32_t g0 = hpgb_angle_to_pixel_fast_g<NEST, LONLAT, DEGREES, HI>(g, f, rk, v0.a.x, v0.b.x, p0, T.flut);
        const uint32_t g1 = hpgb_angle_to_pixel_fast_g<NEST, LONLAT, DEGREES, HI>(g, f, rk, v0.a.y, v0.b.y, p1, T.flut);
        const uint32_t g2 = hpgb_angle_to_pixel_fast_g<NEST, LONLAT, DEGREES, HI>(g, f, rk, v1.a.x, v1.b.x, p2, T.flut);
        const uint32_t g3 = hpgb_angle_to_pixel_fast_g<NEST, LONLAT, DEGREES, HI>(g, f, rk, v1.a.y, v1.b.y, p3, T.flut);
        hpgb_st2(out + i0, p0, p1);
        hpgb_st2(out + i1, p2, p3);
        if (__any_sync(0xFFFFFFFFu, (g0 | g1 | g2 | g3) != 0u)) {  // warp-uniform, 2.5 % of the iterations
            __syncwarp();
            int cnt = bad.aux;  // warp-uniform queue length
            park(g0, i0, cnt, T, bad);
            park(g1, i0 + 1, cnt, T, bad);
            park(g2, i1, cnt, T, bad);
            park(g3, i1 + 1, cnt, T, bad);
            bad.aux = cnt;
        }
    }
    __device__ __forceinline__ void epilogue(const Ctx &T, hpgb_bad_t &bad) const {
        __syncwarp();
        if (bad.aux > 0) drain(queue(), 0, bad.aux, T, bad);
        bad.aux = 0;
    }
#endif
    __device__ __forceinline__ void run4(int64_t i0, const In2 &v0, int64_t i1, const In2 &v1, const Ctx &T,
                                         hpgb_bad_t &bad) const {
        int64_t p0 = 0, p1 = 0, p2 = 0, p3 = 0;
        // guard words: 0 = the fast result can be trusted
        const uint32_t g0 = hpgb_angle_to_pixel_fast_g<NEST, LONLAT, DEGREES, HI>(g, f, rk, v0.a.x, v0.b.x, p0, T.flut);
        const uint32_t g1 = hpgb_angle_to_pixel_fast_g<NEST, LONLAT, DEGREES, HI>(g, f, rk, v0.a.y, v0.b.y, p1, T.flut);
        const uint32_t g2 = hpgb_angle_to_pixel_fast_g<NEST, LONLAT, DEGREES, HI>(g, f, rk, v1.a.x, v1.b.x, p2, T.flut);
        const uint32_t g3 = hpgb_angle_to_pixel_fast_g<NEST, LONLAT, DEGREES, HI>(g, f, rk, v1.a.y, v1.b.y, p3, T.flut);
        if ((g0 | g1 | g2 | g3) != 0u) {
            if (g0) p0 = slow(g, v0.a.x, v0.b.x, i0, T, bad);
            if (g1) p1 = slow(g, v0.a.y, v0.b.y, i0 + 1, T, bad);
            if (g2) p2 = slow(g, v1.a.x, v1.b.x, i1, T, bad);
            if (g3) p3 = slow(g, v1.a.y, v1.b.y, i1 + 1, T, bad);
        }
        hpgb_st2(out + i0, p0, p1);
        hpgb_st2(out + i1, p2, p3);
    }
    __device__ __forceinline__ void run1(int64_t i, const Ctx &T, hpgb_bad_t &bad) const {
        hpgb_geom gg = g;
        hpgb_a2p_fast ff = f;
        if (VAR) {
            if (!var_geom(nside_v, i, NEST, gg)) {
                bad.flag(base + i);
                hpgb_st1(out + i, -1);
                return;
            }
            ff = hpgb_make_a2p_fast<LONLAT, DEGREES>(gg);
        }
        const double va = hpgb_ld1(a + i), vb = hpgb_ld1(b + i);
        int64_t pix;
        if (VAR || hpgb_angle_to_pixel_fast_g<NEST, LONLAT, DEGREES, HI>(gg, ff, rk, va, vb, pix, T.flut) != 0u) pix = slow(gg, va, vb, i, T, bad);
        hpgb_st1(out + i, pix);
    }
};

template <bool NEST, bool LONLAT, bool DEGREES>
int launch_a2p_t(const double *a, const double *b, int64_t *out, int64_t n, int64_t nside, const int64_t *nside_v,
                 int64_t base, hpgb_status_t *st, cudaStream_t s) {
    if (nside_v) {
        OpAng2Pix<NEST, LONLAT, DEGREES, true, false> op{a, b, out, nside_v, base, hpgb_geom(), hpgb_a2p_fast(), hpgb_rk()};
        return hpgb_launch_map1(op, n, st, s);
    }
    const hpgb_geom g = hpgb_make_geom(nside, NEST);
    const bool vec = hpgb_aligned16(a) && hpgb_aligned16(b) && hpgb_aligned16(out);
    if (NEST && g.order >= 16) {  // face bits only in the high word of the pixel: cheaper assembly
        OpAng2Pix<NEST, LONLAT, DEGREES, false, NEST> op{a, b, out, nullptr, base, g, hpgb_make_a2p_fast<LONLAT, DEGREES>(g), hpgb_make_rk(g)};
        if (vec) return hpgb_launch_stream<decltype(op), HPGB_A2P_ST, HPGB_A2P_H>(op, n, st, s);
        return hpgb_launch_map(op, n, vec, st, s);
    }
    OpAng2Pix<NEST, LONLAT, DEGREES, false, false> op{a, b, out, nullptr, base, g, hpgb_make_a2p_fast<LONLAT, DEGREES>(g), hpgb_make_rk(g)};
    // RING with a non-power-of-two nside has no fast path: every point takes the exact path, which wants the
    // occupancy of the plain grid-stride kernel, not the streaming ring
    if (vec && g.order >= 0) return hpgb_launch_stream<decltype(op), HPGB_A2P_ST, HPGB_A2P_H>(op, n, st, s);
    return hpgb_launch_map(op, n, vec, st, s);
}

}  // namespace

int launch_a2p(const double *a, const double *b, int64_t *out, int64_t n, int64_t nside, const int64_t *nside_v,
               int nest, int lonlat, int degrees, int64_t base, hpgb_status_t *st, cudaStream_t s) {
    if (n < 0 || (n > 0 && (!a || !b || !out || !st))) return HPGB_E_BAD_ARGUMENT;
    if (n == 0) return HPGB_OK;
    if (!nside_v) {
        int c = hpgb_check_nside(nside, nest);
        if (c) return c;
    }
#define GO(N, L, D) return launch_a2p_t<N, L, D>(a, b, out, n, nside, nside_v, base, st, s)
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

