// k_reorder.cu -- integer kernels: nest_to_ring / ring_to_nest, neighbors, upgrade_pixels, lonlat_to_thetaphi
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
#include "hpgb_reorder.h"

// ring depth / sub-blocks per tile of the streaming kernel (tunable per build for A/B timing)
#ifndef HPGB_RE_ST
#define HPGB_RE_ST 3
#endif
#ifndef HPGB_RE_H
#define HPGB_RE_H 2
#endif
// ring -> nest streaming kernel: min resident CTAs/SM (register cap), two cap pixels per lane in pass B
#ifndef HPGB_R2N_MINB
#define HPGB_R2N_MINB 4
#endif
#ifndef HPGB_R2N_UNROLL2
#define HPGB_R2N_UNROLL2 0
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
// nest_to_ring / ring_to_nest   (reference loops: hpgeom/hpgeom.c:1429-1450, :1699-1720)
// 8 B in + 8 B out per pixel, integer only.
// -------------------------------------------------------------------------------------------
template <bool TO_RING, bool VAR, bool HI>
struct OpReorder {
    const int64_t *in;
    int64_t *out;
    const int64_t *nside_v;
    int64_t base;  // global index of element 0 (chunked host pipeline)
    hpgb_geom g;
    hpgb_rk c;     // per-nside constants of the minimum-instruction cores (hpgb_reorder.h)
    typedef longlong2 In2;
    typedef hpgb_no_ctx Ctx;

    __device__ __forceinline__ Ctx prologue() const { return Ctx(); }
    // scalar-nside hot path.  HI (nside >= 2^16): npix is a multiple of 2^32, so the range check is
    // one 32-bit compare.  The conversion itself is evaluated unconditionally (pure arithmetic, any
    // input is safe); invalid elements are patched afterwards, off the hot path.
    __device__ __forceinline__ bool ok(int64_t p) const {
        return HI ? ((uint32_t)((uint64_t)p >> 32) < c.npix_hi) : hpgb_pixel_ok(g, p);
    }
    __device__ __forceinline__ int64_t core(int64_t p) const {
        return TO_RING ? hpgb_nest2ring_k<HI>(c, (uint32_t)(uint64_t)p, (uint32_t)((uint64_t)p >> 32))
                       : hpgb_ring2nest_k<HI>(c, g, p);
    }
    __device__ __forceinline__ int64_t conv(int64_t p, int64_t i, hpgb_bad_t &bad) const {
        if (!ok(p)) {
            bad.flag(base + i);
            return -1;
        }
        return core(p);
    }
    __device__ __forceinline__ In2 load2(int64_t i) const { return hpgb_ld2(in + i); }
    static constexpr int NIN = 1;
    static constexpr int MINB = TO_RING ? 1 : HPGB_R2N_MINB;
    // ring -> nest runs the warp-cooperative tile routine below; nest -> ring the per-thread run4
    static constexpr bool TILE_HOOK = !TO_RING && !VAR;
    static constexpr int EXTRA_SMEM = TILE_HOOK ? 128 + (HPGB_BLOCK / 32) * 256 * 2 : 0;
    __device__ __forceinline__ const void *src(int) const { return in; }

    // ---------------------------------------------------------------------------------------
    // ring -> nest, one tile, in three passes over the shared-memory stage (hpgb_reorder.h, "folded coordinates":
    // a slot's 8 bytes go  RING pixel -> (X, Y) -> NEST pixel).  Cap pixels need the isqrt path (~60 instructions),
    // belt pixels a dozen; with random pixels a third of every warp's lanes would drag the other two thirds through
    // the cap path.  So the warp sorts them instead:
    //   pass A  every thread turns its 8 pixels into (X, Y) as if they were belt pixels -- 9 integer-pipe
    //           instructions, one table load -- and stores the pair over the input unless the pixel is a cap pixel;
    //           cap pixels set a bit in a per-thread mask; one warp prefix sum of the bit counts gives every thread
    //           its place in the warp's queue of cap SLOT NUMBERS (2 bytes per entry);
    //   pass B  the warp drains the queue 32 cap pixels at a time (dense lanes): slot -> pixel -> (X, Y) -> slot;
    //   pass C  every thread Morton-interleaves its 8 (X, Y) pairs -- the one place the 26-instruction bit shuffle
    //           exists, executed once per pixel -- and streams the results to HBM with 128-bit stores.
    // A warp only ever touches its own 256 slots, so __syncwarp() is the only synchronisation.
    // ---------------------------------------------------------------------------------------
    static constexpr bool SMEM_PROLOGUE = TILE_HOOK;
    static constexpr int LUT_BYTES = 128;
    __device__ __forceinline__ Ctx prologue_extra(unsigned char *extra) const {
        if (threadIdx.x < 16) {
            hpgb_u32x2 e;
            hpgb_r2n_lut_entry(c, threadIdx.x, e.x, e.y);
            reinterpret_cast<hpgb_u32x2 *>(extra)[threadIdx.x] = e;
        }
        return Ctx();  // visible after the __syncthreads() that follows the mbarrier set-up in k_stream
    }
    template <int TILE>
    __device__ __forceinline__ void tile(unsigned char *stage, unsigned char *extra, int64_t e_base, int tid, const Ctx &,
                                         hpgb_bad_t &bad, uint64_t *empty_bar) const {
        static_assert(TILE == 2048, "queue sizes assume 8 elements per thread");
        const int lane = tid & 31, warp = tid >> 5;
        const hpgb_u32x2 *lut = reinterpret_cast<const hpgb_u32x2 *>(extra);
        unsigned short *qslot = reinterpret_cast<unsigned short *>(extra + LUT_BYTES) + warp * 256;
        longlong2 *st2 = reinterpret_cast<longlong2 *>(stage);
        unsigned long long *st1 = reinterpret_cast<unsigned long long *>(stage);
        hpgb_u32x2 *stx = reinterpret_cast<hpgb_u32x2 *>(stage);
        // ---- pass A ----
        // "cap" here = not provably a belt pixel: ring - nside > 2 nside, or the pixel is so far out of range that the
        // 32-bit ring number aliases (bits k+2 .. 31 of the high word of pix - ncap set; negative pixels included).
        // Every refused pixel is therefore a "cap" pixel and is dealt with in pass B, on dense lanes.
        uint32_t capm = 0u;
        longlong2 v[TILE / 512];
#pragma unroll
        for (int j = 0; j < TILE / 512; j++) v[j] = st2[j * 256 + tid];
        const uint32_t lut_s = hpgb_smem_u32(lut) + 8u, slot_s = hpgb_smem_u32(stage) + 16u * (uint32_t)tid;
        const uint32_t hi_lim = 1u << (c.k + 2u);
#pragma unroll
        for (int j = 0; j < TILE / 512; j++) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const uint64_t ipb = (uint64_t)(e ? v[j].y : v[j].x) - (((uint64_t)c.ncap_hi << 32) | c.ncap_lo);
                const uint32_t tmp = (uint32_t)((int64_t)ipb >> (c.k + 2u));
                const uint32_t jp = ((uint32_t)ipb & c.ns4m1) + (tmp >> 1);
                const uint32_t jm = jp + c.ns - tmp;
                const uint32_t la = lut_s + 16u * (jp >> c.k) + 8u * (jm >> c.k);  // entry 2 ifp + ifm + 1
                // X = jm + lx, Y = ly - jp, stored over the input -- all under "not cap" (the table index of a cap
                // pixel is garbage); cap pixels set their bit in capm
                asm volatile(
                    "{\n\t"
                    ".reg .pred q;\n\t"
                    ".reg .u32 lx, ly;\n\t"
                    "setp.gt.u32 q, %1, %2;\n\t"
                    "setp.ge.or.u32 q, %3, %4, q;\n\t"
                    "@!q ld.shared.v2.u32 {lx, ly}, [%5];\n\t"
                    "@!q add.u32 lx, %6, lx;\n\t"
                    "@!q sub.u32 ly, ly, %7;\n\t"
                    "@!q st.shared.v2.u32 [%8], {lx, ly};\n\t"
                    "@q add.u32 %0, %0, %9;\n\t"
                    "}"
                    : "+r"(capm)
                    : "r"(tmp), "r"(c.ns2), "r"((uint32_t)(ipb >> 32)), "r"(hi_lim), "r"(la), "r"(jm), "r"(jp),
                      "r"(slot_s + (uint32_t)(j * 4096 + e * 8)), "r"(1u << (2 * j + e))
                    : "memory");
            }
        }
        // exclusive prefix sum of the per-thread cap counts over the warp
        const uint32_t mine = (uint32_t)__popc(capm);
        uint32_t incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)  // shfl.up's predicate = "the source lane exists"
            asm volatile("{\n\t.reg .pred q;\n\t.reg .u32 u;\n\tshfl.sync.up.b32 u|q, %0, %1, 0, 0xffffffff;\n\t@q add.u32 %0, %0, u;\n\t}"
                         : "+r"(incl)
                         : "r"(d));
        const int ncap = (int)__shfl_sync(0xFFFFFFFFu, incl, 31);
        unsigned short *qw = qslot + (incl - mine);
#pragma unroll
        for (int b = 0; b < 8; b++) {
            if (capm & (1u << b)) *qw++ = (unsigned short)((b >> 1) * 512 + 2 * tid + (b & 1));
        }
        __syncwarp();
        // ---- pass B ----  (two cap pixels per lane and step while the queue holds more than 32: the isqrt chain is
        // ~250 cycles of dependent instructions, a second independent chain hides half of it)
        auto cap1 = [&](int slot) {
            const uint64_t p = st1[slot];
            hpgb_u32x2 r;
            if (ok((int64_t)p)) {
                hpgb_r2n_cap_fold<HI>(c, g, p, r.x, r.y);
            } else {  // refused: result -1 = interleave(~0, ~0)
                bad.flag(base + e_base + slot);
                r.x = r.y = ~0u;
            }
            stx[slot] = r;
        };
#if HPGB_R2N_UNROLL2
        int j0 = 0;
        for (; j0 + 32 < ncap; j0 += 64) {  // warp-uniform condition: the second half-step has at least one entry
            const int ja = j0 + lane, jb = ja + 32;
            const bool vb = jb < ncap;
            const int sa = qslot[ja], sb = qslot[vb ? jb : ja];
            const uint64_t pa = st1[sa], pb = st1[sb];
            hpgb_u32x2 ra, rb;
            const bool oka = ok((int64_t)pa), okb = ok((int64_t)pb);
            hpgb_r2n_cap_fold<HI>(c, g, oka ? pa : 0ull, ra.x, ra.y);
            hpgb_r2n_cap_fold<HI>(c, g, okb ? pb : 0ull, rb.x, rb.y);
            if (!(oka && okb)) {
                if (!oka) bad.flag(base + e_base + sa), ra.x = ra.y = ~0u;
                if (!okb) bad.flag(base + e_base + sb), rb.x = rb.y = ~0u;
            }
            stx[sa] = ra;
            if (vb) stx[sb] = rb;
        }
        if (j0 + lane < ncap) cap1(qslot[j0 + lane]);
#else
        for (int j = lane; j < ncap; j += 32) cap1(qslot[j]);
#endif
        __syncwarp();
        // ---- pass C ----
#pragma unroll
        for (int j = 0; j < TILE / 512; j++) {
            const int slot = j * 512 + 2 * tid;
            const longlong2 v = st2[slot >> 1];
            const uint64_t a = hpgb_interleave((uint32_t)(uint64_t)v.x, (uint32_t)((uint64_t)v.x >> 32));
            const uint64_t b = hpgb_interleave((uint32_t)(uint64_t)v.y, (uint32_t)((uint64_t)v.y >> 32));
            hpgb_st2(out + e_base + slot, (int64_t)a, (int64_t)b);
        }
        // generic-proxy writes to the stage must be ordered before the TMA refill (async proxy)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) hpgb_mbar_arrive(empty_bar);
    }
    __device__ __forceinline__ In2 load2s(const unsigned char *stage, int, int e) const {
        return *reinterpret_cast<const longlong2 *>(stage + 8 * e);
    }
    __device__ __forceinline__ void run2(int64_t i, const In2 &v, const Ctx &, hpgb_bad_t &bad) const {
        hpgb_st2(out + i, conv(v.x, i, bad), conv(v.y, i + 1, bad));
    }
    __device__ __forceinline__ void run4(int64_t i0, const In2 &v0, int64_t i1, const In2 &v1, const Ctx &,
                                         hpgb_bad_t &bad) const {
        int64_t r0 = core(v0.x), r1 = core(v0.y), r2 = core(v1.x), r3 = core(v1.y);
        const bool k0 = ok(v0.x), k1 = ok(v0.y), k2 = ok(v1.x), k3 = ok(v1.y);
        if (!(k0 & k1 & k2 & k3)) {  // rare: at least one pixel out of range
            if (!k0) r0 = -1, bad.flag(base + i0);
            if (!k1) r1 = -1, bad.flag(base + i0 + 1);
            if (!k2) r2 = -1, bad.flag(base + i1);
            if (!k3) r3 = -1, bad.flag(base + i1 + 1);
        }
        hpgb_st2(out + i0, r0, r1);
        hpgb_st2(out + i1, r2, r3);
    }
    __device__ __forceinline__ void run1(int64_t i, const Ctx &, hpgb_bad_t &bad) const {
        if (VAR) {  // per-element nside: general cores
            hpgb_geom gg;
            const int64_t p = hpgb_ld1(in + i);
            if (!var_geom(nside_v, i, 1, gg) || !hpgb_pixel_ok(gg, p)) {
                bad.flag(base + i);
                hpgb_st1(out + i, -1);
                return;
            }
            hpgb_st1(out + i, TO_RING ? hpgb_nest2ring(gg, p) : hpgb_ring2nest(gg, p));
        } else {
            hpgb_st1(out + i, conv(hpgb_ld1(in + i), i, bad));
        }
    }
};

template <bool TO_RING>
int launch_reorder(const int64_t *in, int64_t *out, int64_t n, int64_t nside, const int64_t *nside_v, int64_t base,
                   hpgb_status_t *st, cudaStream_t s) {
    if (n < 0 || (n > 0 && (!in || !out || !st))) return HPGB_E_BAD_ARGUMENT;
    if (n == 0) return HPGB_OK;  // the reference checks nside inside the loop: no elements, no error
    if (nside_v) {
        OpReorder<TO_RING, true, false> op{in, out, nside_v, base, hpgb_geom(), hpgb_rk()};
        return hpgb_launch_map1(op, n, st, s);
    }
    int c = hpgb_check_nside(nside, 1);
    if (c) return c;
    const hpgb_geom g = hpgb_make_geom(nside, 1);
#if defined(HPGB_NO_STREAM)
    const bool vec = false;
    if (g.order >= 16) {
        OpReorder<TO_RING, false, true> op{in, out, nullptr, base, g, hpgb_make_rk(g)};
        return hpgb_launch_map(op, n, true, st, s);
    }
#else
    const bool vec = hpgb_aligned16(in) && hpgb_aligned16(out);
#endif
    if (g.order >= 16) {
        OpReorder<TO_RING, false, true> op{in, out, nullptr, base, g, hpgb_make_rk(g)};
        if (vec) return hpgb_launch_stream<decltype(op), HPGB_RE_ST, HPGB_RE_H>(op, n, st, s);
        return hpgb_launch_map(op, n, false, st, s);
    }
    OpReorder<TO_RING, false, false> op{in, out, nullptr, base, g, hpgb_make_rk(g)};
    if (vec) return hpgb_launch_stream<decltype(op), HPGB_RE_ST, HPGB_RE_H>(op, n, st, s);
    return hpgb_launch_map(op, n, false, st, s);
}

// -------------------------------------------------------------------------------------------
// neighbors   (hpgeom/hpgeom.c:2928-2959)   8 B in + 64 B out per pixel, rows of 8 (row-major)
// One thread per pixel; the row leaves as four 128-bit streaming stores.
// -------------------------------------------------------------------------------------------
template <bool NEST, bool POW2, bool VAR>
struct OpNeighbors {
    const int64_t *pix;
    int64_t *out;
    const int64_t *nside_v;
    int64_t base;
    hpgb_geom g;
    hpgb_rk c;  // constants of the minimum-instruction RING core (scalar power-of-two nside)
    typedef longlong2 In2;
    typedef hpgb_no_ctx Ctx;
    HPGB_DEFAULT_RUN4
    HPGB_STREAM_IN_I64(pix, 2)
    __device__ __forceinline__ Ctx prologue() const { return Ctx(); }
    __device__ __forceinline__ In2 load2(int64_t i) const { return hpgb_ld2(pix + i); }
    __device__ __forceinline__ void compute(const hpgb_geom &gg, int64_t p, int64_t i, hpgb_bad_t &bad, int64_t *r) const {
        if (!hpgb_pixel_ok(gg, p)) {
            bad.flag(base + i);
#pragma unroll
            for (int m = 0; m < 8; m++) r[m] = -1;
        } else if (!NEST && POW2 && !VAR) {
            hpgb_neighbors_ring_k(c, gg, p, r);
        } else {
            hpgb_neighbors<NEST, POW2>(gg, p, r);
        }
    }
    __device__ __forceinline__ void row(const hpgb_geom &gg, int64_t p, int64_t i, hpgb_bad_t &bad) const {
        int64_t r[8];
        compute(gg, p, i, bad, r);
        int64_t *o = out + 8 * i;  // 64-byte rows: always 16-byte aligned when `out` is
#pragma unroll
        for (int m = 0; m < 8; m += 2) hpgb_st2(o + m, r[m], r[m + 1]);
    }
    // two consecutive pixels per lane = 128 contiguous output bytes per lane, 4 KiB per warp:
    // transposed through shared memory so every store instruction is fully coalesced
    __device__ __forceinline__ void run2(int64_t i, const In2 &v, const Ctx &, hpgb_bad_t &bad) const {
        __shared__ longlong2 sbuf[HPGB_BLOCK / 32][32 * 8];
        int64_t r0[8], r1[8];
        compute(g, v.x, i, bad, r0);
        compute(g, v.y, i + 1, bad, r1);
        if (__activemask() == 0xFFFFFFFFu) {
            longlong2 u[8];
#pragma unroll
            for (int m = 0; m < 4; m++) {
                u[m] = make_longlong2(r0[2 * m], r0[2 * m + 1]);
                u[4 + m] = make_longlong2(r1[2 * m], r1[2 * m + 1]);
            }
            const int lane = threadIdx.x & 31;
            hpgb_warp_store_units<8>(reinterpret_cast<longlong2 *>(out + 8 * (i - 2 * lane)), u, sbuf[threadIdx.x >> 5]);
        } else {
#pragma unroll
            for (int m = 0; m < 8; m += 2) {
                hpgb_st2(out + 8 * i + m, r0[m], r0[m + 1]);
                hpgb_st2(out + 8 * i + 8 + m, r1[m], r1[m + 1]);
            }
        }
    }
    __device__ __forceinline__ void run1(int64_t i, const Ctx &, hpgb_bad_t &bad) const {
        hpgb_geom gg = g;
        if (VAR && !var_geom(nside_v, i, NEST, gg)) {
            bad.flag(base + i);
            for (int m = 0; m < 8; m++) out[8 * i + m] = -1;
            return;
        }
        row(gg, hpgb_ld1(pix + i), i, bad);
    }
};

// neighbours, scalar nside = 2^k, 16-byte aligned operands.  The per-pixel body (RING: ring -> xyf, five
// ring set-ups, eight offsets, hpgb_neighbors_ring_k, ~300 instructions; NEST: Morton-plane arithmetic)
// was inlined four times by the generic skeleton and overflowed the instruction cache (RING: 54 %, NEST:
// 39 % of the stall samples were instruction fetch).  Here it exists ONCE: a warp takes 64 consecutive pixels, every lane converts its two pixels in
// a rolled loop and drops the rows into the XOR-swizzled staging buffer, then the warp writes the 4 KiB
// of rows with fully coalesced 128-bit streaming stores.  The next chunk's pixels are loaded before the
// current one is converted.
template <bool NEST>
__global__ void __launch_bounds__(HPGB_BLOCK) k_neighbors_rows(const int64_t *__restrict__ pix, int64_t *__restrict__ out,
                                                               const int64_t n, const hpgb_geom g, const hpgb_rk c,
                                                               const int64_t base, hpgb_status_t *st) {
    __shared__ longlong2 sbuf[HPGB_BLOCK / 32][32 * 8];
    const int lane = threadIdx.x & 31;
    longlong2 *ws = sbuf[threadIdx.x >> 5];
    const int64_t gw = (int64_t)blockIdx.x * (HPGB_BLOCK / 32) + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * (HPGB_BLOCK / 32);
    const int64_t nchunk = n >> 6;
    hpgb_bad_t bad;
    auto convert = [&](int64_t p, int64_t i, int64_t *r) {
        if (!hpgb_pixel_ok(g, p)) {
            bad.flag(base + i);
#pragma unroll
            for (int m = 0; m < 8; m++) r[m] = -1;
        } else if (NEST) {
            hpgb_neighbors_nest_k(g, p, r);
        } else {
            hpgb_neighbors_ring_k(c, g, p, r);
        }
    };
    // the pixels of the next chunk(s) are in flight while one is converted (measured: two ahead is best for
    // NEST, +3 %; the RING body needs the registers more)
#ifndef HPGB_NB_AHEAD
#define HPGB_NB_AHEAD (NEST ? 2 : 1)
#endif
    int64_t ch = gw;
    longlong2 v[HPGB_NB_AHEAD];
#pragma unroll
    for (int a = 0; a < HPGB_NB_AHEAD; a++) {
        v[a] = make_longlong2(0, 0);
        if (ch + a * nw < nchunk) v[a] = hpgb_ld2(pix + (ch + a * nw) * 64 + 2 * lane);
    }
    for (; ch < nchunk; ch += nw) {
        const longlong2 cur = v[0];
#pragma unroll
        for (int a = 0; a + 1 < HPGB_NB_AHEAD; a++) v[a] = v[a + 1];
        if (ch + HPGB_NB_AHEAD * nw < nchunk) v[HPGB_NB_AHEAD - 1] = hpgb_ld2(pix + (ch + HPGB_NB_AHEAD * nw) * 64 + 2 * lane);
#if defined(HPGB_NB_DIRECT256)
        // A/B: rows straight from the computing lane as two 256-bit stores each (no staging)
#pragma unroll 1
        for (int q = 0; q < 2; q++) {
            int64_t r[8];
            const int64_t i = ch * 64 + 2 * lane + q;
            convert(q ? cur.y : cur.x, i, r);
            hpgb_st4(out + 8 * i, r[0], r[1], r[2], r[3]);
            hpgb_st4(out + 8 * i + 4, r[4], r[5], r[6], r[7]);
        }
#else
#pragma unroll 1
        for (int q = 0; q < 2; q++) {
            int64_t r[8];
            convert(q ? cur.y : cur.x, ch * 64 + 2 * lane + q, r);
#pragma unroll
            for (int j = 0; j < 4; j++) ws[lane * 8 + ((q * 4 + j) ^ (lane & 7))] = make_longlong2(r[2 * j], r[2 * j + 1]);
        }
        __syncwarp();
        longlong2 *dst = reinterpret_cast<longlong2 *>(out + ch * 512);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int idx = k * 32 + lane, src = idx >> 3, j = idx & 7;
            hpgb_row_st(dst + idx, ws[src * 8 + (j ^ (src & 7))]);
        }
        __syncwarp();
#endif
    }
    // tail (< 64 pixels): one pixel per thread, direct stores
    const int64_t t = (nchunk << 6) + (int64_t)blockIdx.x * HPGB_BLOCK + threadIdx.x;
    if (t < n) {
        int64_t r[8];
        convert(hpgb_ld1(pix + t), t, r);
#pragma unroll
        for (int m = 0; m < 8; m += 2) hpgb_st2(out + 8 * t + m, r[m], r[m + 1]);
    }
    bad.commit(st);
}

}  // namespace

int launch_neighbors(const int64_t *pix, int64_t *out, int64_t n, int64_t nside, const int64_t *nside_v, int nest,
                     int64_t base, hpgb_status_t *st, cudaStream_t s) {
    if (n < 0 || (n > 0 && (!pix || !out || !st))) return HPGB_E_BAD_ARGUMENT;
    if (n == 0) return HPGB_OK;
    if (!hpgb_aligned16(out)) return HPGB_E_BAD_ARGUMENT;  // rows are written with 128-bit stores
    if (nside_v) {
        if (nest) {
            OpNeighbors<true, true, true> op{pix, out, nside_v, base, hpgb_geom(), hpgb_rk()};
            return hpgb_launch_map1(op, n, st, s);
        }
        OpNeighbors<false, false, true> op{pix, out, nside_v, base, hpgb_geom(), hpgb_rk()};
        return hpgb_launch_map1(op, n, st, s);
    }
    int c = hpgb_check_nside(nside, nest);
    if (c) return c;
    const hpgb_geom g = hpgb_make_geom(nside, nest);
    const bool vec = hpgb_aligned16(pix);
    if (g.order >= 0 && vec) {
        static hpgb_occ_cache cache[2];
        const int occ = nest ? hpgb_resident_ctas_cached(cache[1], k_neighbors_rows<true>, 0, HPGB_BLOCK)
                             : hpgb_resident_ctas_cached(cache[0], k_neighbors_rows<false>, 0, HPGB_BLOCK);
        int64_t grid = (int64_t)hpgb_sm_count() * occ;
        const int64_t want = ((n >> 6) + HPGB_BLOCK / 32 - 1) / (HPGB_BLOCK / 32);
        if (want < grid) grid = want > 0 ? want : 1;
        if (nest)
            k_neighbors_rows<true><<<(unsigned)grid, HPGB_BLOCK, 0, s>>>(pix, out, n, g, hpgb_rk(), base, st);
        else
            k_neighbors_rows<false><<<(unsigned)grid, HPGB_BLOCK, 0, s>>>(pix, out, n, g, hpgb_make_rk(g), base, st);
        g_hpgb_launches.fetch_add(1, std::memory_order_relaxed);
        cudaError_t e = cudaGetLastError();
        return e == cudaSuccess ? HPGB_OK : HPGB_E_CUDA + (int)e;
    }
    if (nest) {
        OpNeighbors<true, true, false> op{pix, out, nullptr, base, g, hpgb_rk()};
        return hpgb_launch_map(op, n, vec, st, s);
    }
    if (g.order >= 0) {
        OpNeighbors<false, true, false> op{pix, out, nullptr, base, g, hpgb_make_rk(g)};
        return hpgb_launch_map(op, n, vec, st, s);
    }
    OpNeighbors<false, false, false> op{pix, out, nullptr, base, g, hpgb_rk()};
    return hpgb_launch_map(op, n, vec, st, s);
}

namespace {

// -------------------------------------------------------------------------------------------
// upgrade_pixels   (hpgeom/hpgeom.py:515-556 fused: [ring->nest] -> children -> [nest->ring])
// One thread per OUTPUT pixel (coalesced 8 B..16 B stores); the parent is re-read through L1/L2
// by its 4^k children.  Element error classes (priority via the high bits of the atomicMin key):
//   class 0: pixel out of range              (the reference's ring_to_nest check)
//   class 1: nest pixel == npix-1            (the reference range-checks p+1, hpgeom.py:505)
// -------------------------------------------------------------------------------------------
#define HPGB_UPGRADE_CLASS1 ((int64_t)1 << 62)

// One thread per group of C consecutive children (NEST: 2 = one 16-byte unit, every warp store is
// 512 contiguous bytes; RING: 4 = two 128-bit stores).  The parent is converted ring -> nest once per group,
// each child nest -> ring with the minimum-instruction cores of hpgb_reorder.h.  For
// nside_upgrade == nside (one child) the group is the single child.
template <bool NEST, int C, bool ST256 = false>
__global__ void __launch_bounds__(HPGB_BLOCK) k_upgrade(const int64_t *__restrict__ pix, int64_t *__restrict__ out,
                                                        const int64_t n_groups, const int gshift, const hpgb_geom g,
                                                        const hpgb_rk c, const hpgb_rk cu, const int64_t base,
                                                        hpgb_status_t *st) {
    const int64_t stride = (int64_t)gridDim.x * HPGB_BLOCK;
    hpgb_bad_t bad;
    auto emit = [&](int64_t w, int64_t p) {
        const int64_t i = w >> gshift;  // parent
        int64_t r[C];
        if (!hpgb_pixel_ok(g, p)) {
            bad.flag(base + i);
#pragma unroll
            for (int j = 0; j < C; j++) r[j] = -1;
        } else {
            if (NEST) {
                if (p == g.npix - 1) bad.flag(HPGB_UPGRADE_CLASS1 | (base + i));
                // first child of the group: parent << 2k, plus C * (group index within the parent)
                const int64_t first = (p << (gshift + (C == 4 ? 2 : (C == 2 ? 1 : 0)))) + ((w & (((int64_t)1 << gshift) - 1)) * C);
#pragma unroll
                for (int j = 0; j < C; j++) r[j] = first + j;
            } else {
                bool last;
                hpgb_upgrade_ring_group<C>(c, cu, g, p, (uint64_t)(w & (((int64_t)1 << gshift) - 1)) * C, r, last);
                if (last) bad.flag(HPGB_UPGRADE_CLASS1 | (base + i));
            }
        }
        if (C == 4) {
            if (ST256) {  // four children = one 32-byte sector: one 256-bit store
                hpgb_st4(out + 4 * w, r[0], r[1], r[2], r[3]);
            } else {
                hpgb_st2(out + 4 * w, r[0], r[1]);
                hpgb_st2(out + 4 * w + 2, r[2], r[3]);
            }
        } else if (C == 2) {
            hpgb_st2(out + 2 * w, r[0], r[C - 1]);
        } else {
            hpgb_st1(out + w, r[0]);
        }
    };
    // four parent loads in flight per thread: each warp-wide load only brings in 8..32 distinct
    // parents, so without this the kernel is bound by load latency, not bandwidth
#ifndef HPGB_UPG_RING_U
#define HPGB_UPG_RING_U 4
#endif
#ifndef HPGB_UPG_NEST_U
#define HPGB_UPG_NEST_U 16
#endif
    constexpr int U = NEST ? HPGB_UPG_NEST_U : HPGB_UPG_RING_U;
    int64_t w = (int64_t)blockIdx.x * HPGB_BLOCK + threadIdx.x;
    for (; w + (U - 1) * stride < n_groups; w += U * stride) {
        int64_t p[U];
#pragma unroll
        for (int u = 0; u < U; u++) p[u] = __ldg(pix + ((w + u * stride) >> gshift));
#pragma unroll
        for (int u = 0; u < U; u++) emit(w + u * stride, p[u]);
    }
    for (; w < n_groups; w += stride) emit(w, __ldg(pix + (w >> gshift)));
    bad.commit(st);
}

}  // namespace

int launch_upgrade(const int64_t *pix, int64_t *out, int64_t n, int64_t nside, int64_t nside_upgrade, int nest,
                   int64_t base, hpgb_status_t *st, cudaStream_t s) {
    if (n < 0 || (n > 0 && (!pix || !out || !st))) return HPGB_E_BAD_ARGUMENT;
    int c = hpgb_check_nside(nside, 1);
    if (c) return c;
    c = hpgb_check_nside(nside_upgrade, 1);
    if (c || nside_upgrade < nside) return HPGB_E_BAD_ARGUMENT;
    if (n == 0) return HPGB_OK;
    const hpgb_geom g = hpgb_make_geom(nside, 1), gu = hpgb_make_geom(nside_upgrade, 1);
    const hpgb_rk rk = hpgb_make_rk(g), rku = hpgb_make_rk(gu);
    const int shift = 2 * (gu.order - g.order);
    // children per thread: NEST 2 (one fully coalesced 16-byte unit per lane), RING 4 (the parent's
    // ring -> nest conversion is shared by the four), 1 when nside_upgrade == nside or `out` is misaligned
    // (out 32-byte aligned: four children per thread in both schemes, written by one 256-bit store)
#if defined(HPGB_UPG_NO256)
    const bool st256 = false;
#else
    const bool st256 = shift >= 2 && hpgb_aligned32(out);
#endif
    const bool four = shift >= 2 && hpgb_aligned16(out) && (!nest || st256), two = shift >= 2 && hpgb_aligned16(out) && nest && !four;
    const int gshift = four ? shift - 2 : (two ? shift - 1 : shift);  // log2(groups per parent)
    const int64_t n_groups = n << gshift;
    int64_t grid = (int64_t)hpgb_sm_count() * 8;
    const int64_t want = (n_groups + HPGB_BLOCK - 1) / HPGB_BLOCK;
    if (want < grid) grid = want;
#define GO(N, C, S) k_upgrade<N, C, S><<<(unsigned)grid, HPGB_BLOCK, 0, s>>>(pix, out, n_groups, gshift, g, rk, rku, base, st)
    if (nest) {
        if (four) GO(true, 4, true); else if (two) GO(true, 2, false); else GO(true, 1, false);
    } else {
        if (four && st256) GO(false, 4, true); else if (four) GO(false, 4, false); else GO(false, 1, false);
    }
#undef GO
    g_hpgb_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? HPGB_OK : HPGB_E_CUDA + (int)e;
}

namespace {

// -------------------------------------------------------------------------------------------
// lonlat_to_thetaphi   (hpgeom/hpgeom.py:60-88, numpy formulas)   16 B in + 16 B out
// -------------------------------------------------------------------------------------------
template <bool DEGREES>
struct OpLonLat {
    const double *lon, *lat;
    double *theta, *phi;
    int64_t base;
    struct In2 {
        double2 lon, lat;
    };
    typedef hpgb_no_ctx Ctx;
    HPGB_DEFAULT_RUN4
    HPGB_STREAM_IN_F64x2(lon, lat, 4)
    __device__ __forceinline__ Ctx prologue() const { return Ctx(); }
    __device__ __forceinline__ In2 load2(int64_t i) const { return In2{hpgb_ld2(lon + i), hpgb_ld2(lat + i)}; }
    __device__ __forceinline__ void run2(int64_t i, const In2 &v, const Ctx &, hpgb_bad_t &bad) const {
        double t0, p0, t1, p1;
        if (!hpgb_lonlat_to_thetaphi_np<DEGREES>(v.lon.x, v.lat.x, t0, p0)) bad.flag(base + i);
        if (!hpgb_lonlat_to_thetaphi_np<DEGREES>(v.lon.y, v.lat.y, t1, p1)) bad.flag(base + i + 1);
        hpgb_st2(theta + i, t0, t1);
        hpgb_st2(phi + i, p0, p1);
    }
    __device__ __forceinline__ void run1(int64_t i, const Ctx &, hpgb_bad_t &bad) const {
        double t, p;
        if (!hpgb_lonlat_to_thetaphi_np<DEGREES>(hpgb_ld1(lon + i), hpgb_ld1(lat + i), t, p)) bad.flag(base + i);
        hpgb_st1(theta + i, t);
        hpgb_st1(phi + i, p);
    }
};

}  // namespace

int launch_lonlat(const double *lon, const double *lat, double *theta, double *phi, int64_t n, int degrees,
                  int64_t base, hpgb_status_t *st, cudaStream_t s) {
    if (n < 0 || (n > 0 && (!lon || !lat || !theta || !phi || !st))) return HPGB_E_BAD_ARGUMENT;
    const bool vec = hpgb_aligned16(lon) && hpgb_aligned16(lat) && hpgb_aligned16(theta) && hpgb_aligned16(phi);
    if (degrees) {
        OpLonLat<true> op{lon, lat, theta, phi, base};
        return hpgb_launch_auto(op, n, vec, st, s);
    }
    OpLonLat<false> op{lon, lat, theta, phi, base};
    return hpgb_launch_auto(op, n, vec, st, s);
}

int launch_n2r(const int64_t *in, int64_t *out, int64_t n, int64_t nside, const int64_t *nside_v, int64_t base,
               hpgb_status_t *st, cudaStream_t s) {
    return launch_reorder<true>(in, out, n, nside, nside_v, base, st, s);
}
int launch_r2n(const int64_t *in, int64_t *out, int64_t n, int64_t nside, const int64_t *nside_v, int64_t base,
               hpgb_status_t *st, cudaStream_t s) {
    return launch_reorder<false>(in, out, n, nside, nside_v, base, st, s);
}
