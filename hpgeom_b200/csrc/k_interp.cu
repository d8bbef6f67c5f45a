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
// 640-thread CTA per SM shares them.  The sin/cos table of the reference-shaped fallback is static.
static __device__ const unsigned long long g_hpgb_acos_tab[HPGB_ACOS_ROWS * HPGB_ACOS_ROW_WORDS] = HPGB_ACOS_TABLE_INIT;
static __device__ const double g_hpgb_atanq_tab[HPGB_ATANQ_ROWS * HPGB_ATANQ_ROW_DOUBLES] = HPGB_ATANQ_TABLE_INIT;
#define HPGB_INTERP_TAB_WORDS (HPGB_ACOS_ROWS * HPGB_ACOS_ROW_WORDS + HPGB_ATANQ_ROWS * HPGB_ATANQ_ROW_DOUBLES)
#ifndef HPGB_INTERP_BLOCK
#define HPGB_INTERP_BLOCK 640
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
    static constexpr bool kLonLat = LONLAT, kDegrees = DEGREES;
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
    // the ring-table kernel keeps the acos / atan2 tables in global memory (only its prologue reads them)
    __device__ __forceinline__ Ctx prologue_global_tables() const {
        __shared__ double sT[4 * HPGB_SINCOS_ROWS];
        for (int i = threadIdx.x; i < 4 * HPGB_SINCOS_ROWS; i += blockDim.x) sT[i] = g_hpgb_sincos[i];
        __syncthreads();
        return Ctx{sT, reinterpret_cast<const uint64_t *>(g_hpgb_acos_tab), g_hpgb_atanq_tab};
    }
    __device__ __forceinline__ In2 load2(int64_t i) const { return In2{hpgb_ld2(a + i), hpgb_ld2(b + i)}; }
    static __device__ __noinline__ void slow(const hpgb_geom &gg, const double *T, double theta, double phi, int64_t *p,
                                             double *w) {
        hpgb_interpol<NEST, POW2>(gg, T, theta, phi, p, w);
    }
    // the four pixels and weights of one point (the reference's loop body, hpgeom/hpgeom.c:3491-3523)
    __device__ __forceinline__ void values(const hpgb_geom &gg, double va, double vb, int64_t i, const Ctx &c, hpgb_bad_t &bad,
                                           int64_t *p, double *w) const {
        p[0] = p[1] = p[2] = p[3] = -1;
        w[0] = w[1] = w[2] = w[3] = 0.0;
        double theta, phi;
        if (!hpgb_angles_exact<LONLAT, DEGREES>(va, vb, theta, phi)) {
            bad.flag(base + i);
        } else {
            int64_t ir1;
            if (FAST && hpgb_ring_above_fast(gg, f, theta, ir1)) {
                const hpgb_mp_tab mp{c.TA, c.TQ};
                hpgb_interpol_rings<NEST, true>(gg, rk, mp, theta, phi, ir1, p, w);
            } else {
                int64_t tp[4];  // the out-of-line call gets its own row, so p / w stay in registers
                double tw[4];
                slow(gg, c.T, theta, phi, tp, tw);
#pragma unroll
                for (int m = 0; m < 4; m++) p[m] = tp[m], w[m] = tw[m];
            }
        }
    }
    __device__ __forceinline__ void row(const hpgb_geom &gg, double va, double vb, int64_t i, const Ctx &c, hpgb_bad_t &bad) const {
        int64_t p[4];
        double w[4];
        values(gg, va, vb, i, c, bad, p, w);
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

// Scalar power-of-two nside, aligned operands.  The body of one point is ~1000 instructions; the generic
// skeleton inlined it four times (k_map2: two pairs in flight), 19 000 SASS instructions = 300 KB of code that
// no instruction cache holds.  Here it exists ONCE: a warp takes 64 consecutive points, every lane converts its
// two points in a rolled loop and drops the two (4,) rows into XOR-swizzled staging buffers, then the warp
// writes 2 KiB of pixels and 2 KiB of weights with fully coalesced 128-bit streaming stores.  The next chunk's
// angles are loaded before the current one is converted.
template <class Op>
__global__ void __launch_bounds__(HPGB_INTERP_BLOCK) k_interp_rows(const Op op, const int64_t n, hpgb_status_t *st) {
    extern __shared__ __align__(16) unsigned long long hpgb_interp_tabs[];
    const typename Op::Ctx ctx = op.prologue();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    longlong2 *wp = reinterpret_cast<longlong2 *>(hpgb_interp_tabs + HPGB_INTERP_TAB_WORDS) + warp * 256;  // 128 units pixels,
    longlong2 *ww = wp + 128;                                                                                // 128 units weights
    const int64_t gw = (int64_t)blockIdx.x * (HPGB_INTERP_BLOCK / 32) + warp;
    const int64_t nw = (int64_t)gridDim.x * (HPGB_INTERP_BLOCK / 32);
    const int64_t nfull = n >> 6, nchunk = (n + 63) >> 6;  // the last chunk may be partial: same code, guarded
    hpgb_bad_t bad;
    const int sw = (lane >> 1) & 3;
    auto fetch = [&](int64_t c) {
        typename Op::In2 r;
        const int64_t i = c * 64 + 2 * lane;
        if (c < nfull) return op.load2(i);
        r.a.x = i < n ? hpgb_ld1(op.a + i) : 0.0;  // past the end: a valid dummy point, never stored
        r.b.x = i < n ? hpgb_ld1(op.b + i) : 0.0;
        r.a.y = i + 1 < n ? hpgb_ld1(op.a + i + 1) : 0.0;
        r.b.y = i + 1 < n ? hpgb_ld1(op.b + i + 1) : 0.0;
        return r;
    };
    int64_t ch = gw;
    typename Op::In2 v = typename Op::In2();
    if (ch < nchunk) v = fetch(ch);
    for (; ch < nchunk; ch += nw) {
        const typename Op::In2 cur = v;
        if (ch + nw < nchunk) v = fetch(ch + nw);
#pragma unroll 1
        for (int q = 0; q < 2; q++) {
            int64_t p[4];
            double w[4];
            op.values(op.g, q ? cur.a.y : cur.a.x, q ? cur.b.y : cur.b.x, ch * 64 + 2 * lane + q, ctx, bad, p, w);
#pragma unroll
            for (int j = 0; j < 2; j++) {
                wp[lane * 4 + ((q * 2 + j) ^ sw)] = make_longlong2(p[2 * j], p[2 * j + 1]);
                ww[lane * 4 + ((q * 2 + j) ^ sw)] = make_longlong2(__double_as_longlong(w[2 * j]), __double_as_longlong(w[2 * j + 1]));
            }
        }
        __syncwarp();
        longlong2 *dp = reinterpret_cast<longlong2 *>(op.pix + ch * 256);
        longlong2 *dw = reinterpret_cast<longlong2 *>(op.wgt + ch * 256);
        const int64_t units = ch < nfull ? 128 : 2 * (n - ch * 64);  // two 16-byte units per point
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int idx = k * 32 + lane, src = idx >> 2, j = idx & 3, ssw = (src >> 1) & 3;
            if (idx < units) {
                hpgb_row_st(dp + idx, wp[src * 4 + (j ^ ssw)]);
                hpgb_row_st(dw + idx, ww[src * 4 + (j ^ ssw)]);
            }
        }
        __syncwarp();
    }
    bad.commit(st);
}

// Scalar power-of-two nside, aligned operands, second generation: the points of a warp are SORTED BY ZONE before
// the expensive part.  With random points a third of every warp's lanes lie in the polar caps (ring colatitude from
// the atan2 table, ring length 4 * ring), two thirds in the belt (acos table, 4 nside pixels a ring, one shared
// phi / dphi): k_interp_rows executed both colatitude routines and both ring set-ups for all of them (22-25 of 32
// lanes active on average, 935 instructions a point).  Here a lane first does the cheap part of its two points
// (angle checks, ring pair from the polynomial sine), then parks each point in the warp's BELT or CAP queue in shared
// memory; whenever a queue holds 32 points the warp runs that zone's straight-line body (hpgb_interp_belt /
// hpgb_interp_cap, hpgb_fp.h) on them with every lane busy.  Left-overs stay queued for the next chunk and are
// drained at the end.  The rows go straight to HBM from the lane that computed them: a (4,) int64 row and a (4,)
// float64 row are one 32-byte sector each, so nothing is gained by transposing them through shared memory first.
#define HPGB_INTERP_QCAP 64
#define HPGB_INTERP_QBYTES (HPGB_INTERP_QCAP * (8 + 8 + 8 + 4))
// LUT (nside <= HPGB_INTERP_LUT_MAX_NSIDE, enough points to pay for it): what depends only on the ring -- its colatitude,
// a cap ring's pixel spacing and the reciprocal of that -- is computed ONCE per CTA into shared memory by the prologue
// (2 nside ring entries spread over the CTA's threads, the acos / atan2 tables read from global memory) and the zone
// bodies are the hpgb_interp_*_lut ones of hpgb_fp.h: no inverse trigonometry in the point loop at all.
//   dynamic shared memory: th[2 nside + 2] | cap[nside] (16 B) | 16 fold entries | the warps' queues
static __host__ __device__ inline int hpgb_interp_lut_bytes(int64_t nside) { return (int)((2 * nside + 2) * 8 + nside * 16 + 128); }
template <class Op, bool NEST, bool LUT>
__global__ void __launch_bounds__(HPGB_INTERP_BLOCK) k_interp_sorted(const Op op, const int64_t n, const double dphi_belt,
                                                                    hpgb_status_t *st) {
    extern __shared__ __align__(16) unsigned long long hpgb_interp_tabs[];
    const typename Op::Ctx ctx = LUT ? op.prologue_global_tables() : op.prologue();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t ns_l = (uint32_t)op.g.nside;
    double *th_tab = reinterpret_cast<double *>(hpgb_interp_tabs);
    hpgb_cdiv *cap_tab = reinterpret_cast<hpgb_cdiv *>(th_tab + 2 * ns_l + 2);
    hpgb_u32x2 *flut = reinterpret_cast<hpgb_u32x2 *>(cap_tab + ns_l);
    hpgb_cdiv dphi_c;
    dphi_c.d = dphi_belt;
    dphi_c.inv = 1.0 / dphi_belt;
    if (LUT) {
        for (uint32_t nring = 1u + threadIdx.x; nring <= 2u * ns_l; nring += blockDim.x) {
            double th;
            hpgb_cdiv cd;
            hpgb_interp_ring_entry(op.g, ctx.TA, ctx.TQ, nring, th, cd);
            th_tab[nring] = th;
            if (nring < ns_l) cap_tab[nring] = cd;
        }
        if (threadIdx.x < 16) {
            hpgb_u32x2 e;
            hpgb_r2n_lut_entry(op.rk, threadIdx.x, e.x, e.y);
            flut[threadIdx.x] = e;
        }
        __syncthreads();
    }
    unsigned char *qmem = reinterpret_cast<unsigned char *>(hpgb_interp_tabs) +
                          (LUT ? hpgb_interp_lut_bytes(op.g.nside) : HPGB_INTERP_TAB_WORDS * 8) + warp * (2 * HPGB_INTERP_QBYTES);
    struct Queue {
        double *th, *ph;
        int64_t *idx;
        uint32_t *ir;
        __device__ __forceinline__ explicit Queue(unsigned char *m)
            : th(reinterpret_cast<double *>(m)), ph(th + HPGB_INTERP_QCAP), idx(reinterpret_cast<int64_t *>(ph + HPGB_INTERP_QCAP)),
              ir(reinterpret_cast<uint32_t *>(idx + HPGB_INTERP_QCAP)) {}
    };
    const Queue qb(qmem), qc(qmem + HPGB_INTERP_QBYTES);
    int nb = 0, nc = 0;  // queue lengths, warp-uniform
    const unsigned lt = (1u << lane) - 1u;
    const uint32_t ns = (uint32_t)op.g.nside;
    const double dns = (double)op.g.nside;
    (void)ns;
    const int64_t gw = (int64_t)blockIdx.x * (HPGB_INTERP_BLOCK / 32) + warp;
    const int64_t nw = (int64_t)gridDim.x * (HPGB_INTERP_BLOCK / 32);
    const int64_t nfull = n >> 6, nchunk = (n + 63) >> 6;
    hpgb_bad_t bad;
    // a (4,) row is one 32-byte sector: one 256-bit store when the arrays are 32-byte aligned (every allocation is;
    // a caller's odd 16-byte-aligned view takes the two-halves form).  The points of a queue are not in index order, so
    // a warp's store instruction touches 32 different sectors either way -- the 256-bit form halves the requests.
    const bool st256 = (((uintptr_t)op.pix | (uintptr_t)op.wgt) & 31u) == 0;
    auto emit = [&](int64_t i, const int64_t *p, const double *w) {
        if (st256) {
            hpgb_st4(op.pix + 4 * i, p[0], p[1], p[2], p[3]);
            hpgb_st4(op.wgt + 4 * i, w[0], w[1], w[2], w[3]);
        } else {
            hpgb_st2(op.pix + 4 * i, p[0], p[1]);
            hpgb_st2(op.pix + 4 * i + 2, p[2], p[3]);
            hpgb_st2(op.wgt + 4 * i, w[0], w[1]);
            hpgb_st2(op.wgt + 4 * i + 2, w[2], w[3]);
        }
    };
    auto run_belt = [&](int first, int count) {
        if (lane < count) {
            int64_t p[4];
            double w[4];
            if (LUT)
                hpgb_interp_belt_lut<NEST>(op.g, op.rk, flut, th_tab, dphi_c, qb.th[first + lane], qb.ph[first + lane], qb.ir[first + lane], p, w);
            else
                hpgb_interp_belt<NEST>(op.g, op.rk, ctx.TA, dphi_belt, qb.th[first + lane], qb.ph[first + lane], qb.ir[first + lane], p, w);
            emit(qb.idx[first + lane], p, w);
        }
        __syncwarp();
    };
    auto run_cap = [&](int first, int count) {
        if (lane < count) {
            int64_t p[4];
            double w[4];
            if (LUT)
                hpgb_interp_cap_lut<NEST>(op.g, op.rk, th_tab, cap_tab, qc.th[first + lane], qc.ph[first + lane], qc.ir[first + lane], p, w);
            else
                hpgb_interp_cap<NEST>(op.g, op.rk, ctx.TQ, qc.th[first + lane], qc.ph[first + lane], qc.ir[first + lane], p, w);
            emit(qc.idx[first + lane], p, w);
        }
        __syncwarp();
    };
    auto fetch = [&](int64_t c) {
        typename Op::In2 r;
        const int64_t i = c * 64 + 2 * lane;
        if (c < nfull) return op.load2(i);
        r.a.x = i < n ? hpgb_ld1(op.a + i) : 0.0;
        r.b.x = i < n ? hpgb_ld1(op.b + i) : 0.0;
        r.a.y = i + 1 < n ? hpgb_ld1(op.a + i + 1) : 0.0;
        r.b.y = i + 1 < n ? hpgb_ld1(op.b + i + 1) : 0.0;
        return r;
    };
    int64_t ch = gw;
    typename Op::In2 v = typename Op::In2();
    if (ch < nchunk) v = fetch(ch);
    for (; ch < nchunk; ch += nw) {
        const typename Op::In2 cur = v;
        if (ch + nw < nchunk) v = fetch(ch + nw);
        // the cheap part of BOTH points first, as one straight-line block (two independent dependency chains per lane:
        // this part is latency-bound at 20 warps per SM), then the queueing point by point
        double theta2[2], phi2[2];
        uint32_t ir2[2];
        int zone2[2];  // 0 belt, 1 cap, 2 the reference-shaped routine, 3 nothing to do
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int64_t i = ch * 64 + 2 * lane + q;
            const double va = q ? cur.a.y : cur.a.x, vb = q ? cur.b.y : cur.b.x;
            theta2[q] = phi2[q] = 0.0;
            ir2[q] = 0u;
            zone2[q] = 3;
            if (i < n) {
                if (!hpgb_angles_exact<Op::kLonLat, Op::kDegrees>(va, vb, theta2[q], phi2[q])) {
                    zone2[q] = 4;  // refused
                } else {
#if defined(HPGB_INTERP_ZONE_V1)
                    int64_t ir64 = 0;
                    zone2[q] = (hpgb_ring_above_fast(op.g, op.f, theta2[q], ir64) && hpgb_interp_phi_tame(phi2[q])) ? hpgb_interp_zone(ns, ir64) : 2;
                    ir2[q] = (uint32_t)ir64;
#else
                    const int z = hpgb_ring_zone_fast(op.rk, op.f, dns, theta2[q], ir2[q]);
                    zone2[q] = hpgb_interp_phi_tame(phi2[q]) ? z : 2;
#endif
                }
            }
        }
#pragma unroll 1
        for (int q = 0; q < 2; q++) {
            const int64_t i = ch * 64 + 2 * lane + q;
            const double theta = q ? theta2[1] : theta2[0], phi = q ? phi2[1] : phi2[0];
            const uint32_t ir1 = q ? ir2[1] : ir2[0];
            const int zone = q ? zone2[1] : zone2[0];
            if (zone == 4) {
                bad.flag(op.base + i);
                const int64_t p[4] = {-1, -1, -1, -1};
                const double w[4] = {0.0, 0.0, 0.0, 0.0};
                emit(i, p, w);
            }
            // one set of stores for both queues: the cap queue lies HPGB_INTERP_QBYTES behind the belt queue
            const unsigned mb = __ballot_sync(0xFFFFFFFFu, zone == 0);
            const unsigned mc = __ballot_sync(0xFFFFFFFFu, zone == 1);
            if ((unsigned)zone <= 1u) {
                const int pos = zone ? nc + __popc(mc & lt) : nb + __popc(mb & lt);
                unsigned char *qz = qmem + (zone ? HPGB_INTERP_QBYTES : 0);
                reinterpret_cast<double *>(qz)[pos] = theta;
                reinterpret_cast<double *>(qz + 8 * HPGB_INTERP_QCAP)[pos] = phi;
                reinterpret_cast<int64_t *>(qz + 16 * HPGB_INTERP_QCAP)[pos] = i;
                reinterpret_cast<uint32_t *>(qz + 24 * HPGB_INTERP_QCAP)[pos] = ir1;
            }
            nb += __popc(mb);
            nc += __popc(mc);
            if (zone == 2) {  // pole rows, pairs straddling a cap edge, points on a ring: rare
                int64_t p[4];
                double w[4];
                Op::slow(op.g, ctx.T, theta, phi, p, w);
                emit(i, p, w);
            }
            __syncwarp();
            if (nb >= 32) {
                nb -= 32;
                run_belt(nb, 32);
            }
            if (nc >= 32) {
                nc -= 32;
                run_cap(nc, 32);
            }
        }
    }
    if (nb > 0) run_belt(0, nb);
    if (nc > 0) run_cap(0, nc);
    bad.commit(st);
}

// points below which the ring table does not pay for itself (its prologue costs ~13 ring entries per thread)
#ifndef HPGB_INTERP_LUT_MIN_N
#define HPGB_INTERP_LUT_MIN_N (1 << 20)
#endif
template <class Op, bool NEST, bool LUT>
int launch_interp_sorted_t(const Op &op, int64_t n, hpgb_status_t *st, cudaStream_t s) {
    const int smem = (LUT ? hpgb_interp_lut_bytes(HPGB_INTERP_LUT_MAX_NSIDE) : Op::DYN_SMEM) + (HPGB_INTERP_BLOCK / 32) * 2 * HPGB_INTERP_QBYTES;
    static hpgb_occ_cache cache;  // per device: the dynamic-smem opt-in is a per-device attribute
    (void)hpgb_resident_ctas_cached(cache, k_interp_sorted<Op, NEST, LUT>, smem, HPGB_INTERP_BLOCK);
    int64_t grid = hpgb_sm_count();
    const int64_t want = (((n + 63) >> 6) + HPGB_INTERP_BLOCK / 32 - 1) / (HPGB_INTERP_BLOCK / 32);
    if (want < grid) grid = want > 0 ? want : 1;
    // 2 pi / (4 nside): the belt rings' pixel spacing, the same IEEE division the reference does per ring
    const double dphi_belt = HPGB_TWOPI / (double)(4 * op.g.nside);
    const int smem_launch = (LUT ? hpgb_interp_lut_bytes(op.g.nside) : Op::DYN_SMEM) + (HPGB_INTERP_BLOCK / 32) * 2 * HPGB_INTERP_QBYTES;
    k_interp_sorted<Op, NEST, LUT><<<(unsigned)grid, HPGB_INTERP_BLOCK, smem_launch, s>>>(op, n, dphi_belt, st);
    g_hpgb_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? HPGB_OK : HPGB_E_CUDA + (int)e;
}
template <class Op, bool NEST>
int launch_interp_sorted(const Op &op, int64_t n, hpgb_status_t *st, cudaStream_t s) {
#if !defined(HPGB_INTERP_NO_LUT)
    if (op.g.nside <= HPGB_INTERP_LUT_MAX_NSIDE && n >= HPGB_INTERP_LUT_MIN_N) return launch_interp_sorted_t<Op, NEST, true>(op, n, st, s);
#endif
    return launch_interp_sorted_t<Op, NEST, false>(op, n, st, s);
}

template <class Op>
int launch_interp_rows(const Op &op, int64_t n, hpgb_status_t *st, cudaStream_t s) {
    const int smem = Op::DYN_SMEM + (HPGB_INTERP_BLOCK / 32) * 256 * 16;
    static hpgb_occ_cache cache;  // per device: the dynamic-smem opt-in is a per-device attribute
    (void)hpgb_resident_ctas_cached(cache, k_interp_rows<Op>, smem, HPGB_INTERP_BLOCK);
    int64_t grid = hpgb_sm_count();
    const int64_t want = (((n + 63) >> 6) + HPGB_INTERP_BLOCK / 32 - 1) / (HPGB_INTERP_BLOCK / 32);
    if (want < grid) grid = want > 0 ? want : 1;
    k_interp_rows<Op><<<(unsigned)grid, HPGB_INTERP_BLOCK, smem, s>>>(op, n, st);
    g_hpgb_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? HPGB_OK : HPGB_E_CUDA + (int)e;
}

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
#if defined(HPGB_INTERP_ROWS)
        if (vec) return launch_interp_rows(op, n, st, s);
#else
        if (vec) return launch_interp_sorted<decltype(op), NEST>(op, n, st, s);
#endif
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

