// k_ranges.cu -- pixel_ranges_to_pixels on the device (reference: hpgeom/hpgeom.c:3795-3896): the
// variable-length expansion of (M,2) [lo, hi) ranges into the flat pixel list.  The reference
// walks the ranges twice with one thread (count + validate, :3843-3856; expand, :3868-3876); here
//
//   hpgb_ranges_offsets  validates hi >= lo (first bad row -> status) and writes the exclusive
//                        prefix sums of the range lengths, offsets[0..M] (offsets[M] = output size);
//                        three small launches, no scratch: the per-tile totals travel through the
//                        offsets array itself (slot (t+1)*TILE is tile t's total in phase 1, the
//                        running total after phase 2 -- which IS its final value);
//   hpgb_ranges_expand   writes any window [out_begin, out_begin + out_count) of the pixel list.
//                        Each CTA owns a contiguous run of 2048-pixel output tiles, finds its first
//                        range with ONE binary search and then walks forward: a window of 1024
//                        offsets + range starts is staged in shared memory and KEPT across tiles
//                        (it only moves on when a tile reaches past it).  A full tile inside the
//                        window runs without any barrier: a thread whose eight outputs fall into
//                        one range (long ranges: always, usually the range of its previous tile)
//                        just adds; otherwise it searches the window per 16-byte unit.  128-bit
//                        streaming stores, 8 B written per pixel: 5.95 TB/s on long ranges.
//
// The same pair expands the range tables of query_circle (k_query.cu), single and batched.
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "hpgb_core.h"
#include "hpgb_host.h"
#include "hpgb_internal.h"
#include "hpgb_launch.cuh"

#define RG_BLOCK 256
#define RG_TILE 2048   // outputs per expand tile / ranges per scan tile (8 per thread)
#define RG_CH 1024     // offsets staged per round of the expand kernel

namespace {

// inclusive scan of one int64 per thread over a CTA of NW warps; *total = sum over the CTA.
// sw = NW + 1 shared int64 slots; two barriers.
template <int NW>
__device__ __forceinline__ int64_t block_scan_incl(int64_t v, int64_t *sw, int64_t *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int64_t u = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) v += u;
    }
    __syncthreads();  // previous use of sw is over
    if (lane == 31) sw[warp] = v;
    __syncthreads();
    int64_t before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) {
        const int64_t s = sw[w];
        if (w < warp) before += s;
        all += s;
    }
    *total = all;
    return v + before;
}

__device__ __forceinline__ int64_t range_len(const longlong2 r, int inclusive, bool &bad) {
    bad = r.y < r.x;  // hpgeom/hpgeom.c:3847
    return bad ? 0 : (int64_t)((uint64_t)r.y - (uint64_t)r.x) + inclusive;
}

// phase 1: total length of every tile of RG_TILE ranges -> offs[min((t+1) * RG_TILE, m)]
__global__ void __launch_bounds__(RG_BLOCK) k_ranges_tilesum(const longlong2 *__restrict__ ranges, int64_t m, int inclusive,
                                                             int64_t *__restrict__ offs, hpgb_status_t *st) {
    __shared__ int64_t sw[RG_BLOCK / 32 + 1];
    const int64_t r0 = (int64_t)blockIdx.x * RG_TILE;
    hpgb_bad_t bad;
    int64_t sum = 0;
#pragma unroll
    for (int j = 0; j < RG_TILE / RG_BLOCK; j++) {
        const int64_t r = r0 + j * RG_BLOCK + threadIdx.x;
        if (r < m) {
            bool b;
            sum += range_len(ranges[r], inclusive, b);
            if (b) bad.flag(r);
        }
    }
    int64_t total;
    block_scan_incl<RG_BLOCK / 32>(sum, sw, &total);
    if (threadIdx.x == 0) {
        const int64_t end = r0 + RG_TILE;
        offs[end < m ? end : m] = total;
    }
    bad.commit(st);
}

// phase 2 (one CTA): running totals over the tile slots; offs[0] = 0
__global__ void __launch_bounds__(1024) k_ranges_scan_tiles(int64_t *__restrict__ offs, int64_t m, int64_t ntile) {
    __shared__ int64_t sw[33];
    int64_t carry = 0;
    for (int64_t b = 0; b < ntile; b += 1024) {
        const int64_t t = b + threadIdx.x;
        const int64_t end = (t + 1) * RG_TILE;
        const int64_t pos = end < m ? end : m;
        const int64_t v = t < ntile ? offs[pos] : 0;
        int64_t total;
        const int64_t incl = block_scan_incl<32>(v, sw, &total);
        if (t < ntile) offs[pos] = carry + incl;
        carry += total;
    }
    if (threadIdx.x == 0) offs[0] = 0;
}

// phase 3: exclusive offsets inside every tile (slots i >= 1; slot 0 of a tile is already final)
__global__ void __launch_bounds__(RG_BLOCK) k_ranges_tilescan(const longlong2 *__restrict__ ranges, int64_t m, int inclusive,
                                                              int64_t *__restrict__ offs) {
    __shared__ int64_t sw[RG_BLOCK / 32 + 1];
    const int64_t r0 = (int64_t)blockIdx.x * RG_TILE;
    int64_t carry = blockIdx.x ? offs[r0] : 0;
#pragma unroll 1
    for (int j = 0; j < RG_TILE / RG_BLOCK; j++) {
        const int64_t r = r0 + j * RG_BLOCK + threadIdx.x;
        int64_t len = 0;
        if (r < m) {
            bool b;
            len = range_len(ranges[r], inclusive, b);
        }
        int64_t total;
        const int64_t incl = block_scan_incl<RG_BLOCK / 32>(len, sw, &total);
        // exclusive offset of range r + 1 = carry + incl; slot r0 + RG_TILE and slot m belong to phase 2
        const int64_t slot = r + 1;
        if (slot < m && slot < r0 + RG_TILE) offs[slot] = carry + incl;
        carry += total;
    }
}

// largest idx in [0, RG_CH) with soff[idx] <= o   (soff[0] <= o is an invariant of the caller)
__device__ __forceinline__ int window_find(const int64_t *soff, int64_t o) {
    int idx = 0;
#pragma unroll
    for (int s = RG_CH / 2; s >= 1; s >>= 1)
        if (soff[idx + s] <= o) idx += s;
    return idx;
}

// store policy of the pixel list (A/B switch): 0 = streaming (evict-first), 1 = plain write-back
#ifndef RG_ST
#define RG_ST 0
#endif
__device__ __forceinline__ void rg_st2(int64_t *p, int64_t a, int64_t b) {
#if RG_ST == 1
    *reinterpret_cast<longlong2 *>(p) = make_longlong2(a, b);
#else
    __stcs(reinterpret_cast<longlong2 *>(p), make_longlong2(a, b));
#endif
}

template <bool VEC>
__global__ void __launch_bounds__(RG_BLOCK) k_ranges_expand(const longlong2 *__restrict__ ranges, const int64_t *__restrict__ offs,
                                                            int64_t m, int64_t out_begin, int64_t out_count,
                                                            int64_t *__restrict__ out, int64_t tiles_per_cta) {
    __shared__ int64_t soff[RG_CH + 1];  // offs[wbase .. wbase + RG_CH] (entries past m hold offs[m] = total)
    __shared__ int64_t slo[RG_CH];       // start of ranges wbase .. wbase + RG_CH - 1
    __shared__ int64_t s_cur;
    const int tid = threadIdx.x;
    const int64_t ntiles = (out_count + RG_TILE - 1) / RG_TILE;
    const int64_t t0 = (int64_t)blockIdx.x * tiles_per_cta;
    const int64_t t1 = t0 + tiles_per_cta < ntiles ? t0 + tiles_per_cta : ntiles;
    if (t0 >= t1) return;
    if (tid == 0) {  // the range holding this CTA's first output: largest r < m with offs[r] <= o
        const int64_t o = out_begin + t0 * RG_TILE;
        int64_t lo = 0, hi = m;  // answer in [lo, hi)
        while (hi - lo > 1) {
            const int64_t mid = lo + ((hi - lo) >> 1);
            if (offs[mid] <= o) lo = mid; else hi = mid;
        }
        s_cur = lo;
    }
    __syncthreads();
    int64_t wbase = s_cur;
    auto load_window = [&]() {
        for (int i = tid; i <= RG_CH; i += RG_BLOCK) {
            const int64_t r = wbase + i;
            soff[i] = offs[r < m ? r : m];
            if (i < RG_CH) slo[i] = r < m ? ranges[r].x : 0;
        }
        __syncthreads();
    };
    load_window();
    // The window is kept across tiles and only moves on when a tile reaches past it, so between moves
    // the warps run without any barrier (long ranges: the kernel is a plain fill).
    int idx_prev = 0;  // window slot of the range that held this thread's previous output
    for (int64_t t = t0; t < t1; t++) {
        const int64_t e0 = t * RG_TILE;                      // tile start, relative to out
        const int64_t o0 = out_begin + e0;                   // ... as a global output index
        const int64_t o_end = (e0 + RG_TILE < out_count ? e0 + RG_TILE : out_count) + out_begin;
        // Common case, decided uniformly: a full tile that the current window covers.  No pending mask,
        // no window moves; a thread whose eight outputs fall into one range (long ranges: always, and
        // usually the range of its previous tile) just adds.
        if (VEC && o_end - o0 == RG_TILE && (wbase + RG_CH >= m || soff[RG_CH] >= o_end)) {
            const int64_t of = o0 + 2 * tid, ol = of + 3 * (2 * RG_BLOCK) + 1;
            int idx = idx_prev;
            if (!(soff[idx] <= of && soff[idx + 1] > of)) idx = window_find(soff, of);
            int64_t v[RG_TILE / RG_BLOCK];
            if (soff[idx + 1] > ol) {
                const int64_t b = slo[idx] - soff[idx] + of;
#pragma unroll
                for (int k = 0; k < RG_TILE / RG_BLOCK; k++) v[k] = b + ((k >> 1) * (2 * RG_BLOCK) + (k & 1));
            } else {
#pragma unroll
                for (int k = 0; k < RG_TILE / RG_BLOCK; k++) {
                    const int64_t o = of + ((k >> 1) * (2 * RG_BLOCK) + (k & 1));
                    if (soff[idx + 1] <= o) {
                        if (k & 1) {
                            while (soff[idx + 1] <= o) idx++;  // next output: usually one step (more over empty rows)
                        } else {
                            idx = window_find(soff, o);
                        }
                    }
                    v[k] = slo[idx] + (o - soff[idx]);
                }
            }
            idx_prev = idx;
#pragma unroll
            for (int k = 0; k < RG_TILE / RG_BLOCK; k += 2) rg_st2(out + e0 + (k >> 1) * (2 * RG_BLOCK) + 2 * tid, v[k], v[k + 1]);
            continue;
        }
        int64_t val[RG_TILE / RG_BLOCK];
        unsigned pending = 0;
#pragma unroll
        for (int k = 0; k < RG_TILE / RG_BLOCK; k++)
            if (o0 + (k >> 1) * (2 * RG_BLOCK) + 2 * tid + (k & 1) < o_end) pending |= 1u << k;
        for (;;) {  // uniform over the CTA: wbase, cover, o_end are
            const bool last_window = wbase + RG_CH >= m;
            const int64_t cover = soff[RG_CH];  // outputs below it belong to ranges inside the window
            int idx = -1;
#pragma unroll
            for (int k = 0; k < RG_TILE / RG_BLOCK; k++) {
                const int64_t o = o0 + (k >> 1) * (2 * RG_BLOCK) + 2 * tid + (k & 1);
                if (((pending >> k) & 1u) && (last_window || o < cover)) {
                    if (idx < 0 || soff[idx + 1] <= o) {
                        if ((k & 1) && idx >= 0) {
                            while (idx + 1 < RG_CH && soff[idx + 1] <= o) idx++;  // next output: usually one step (more over empty rows)
                        } else {
                            idx = window_find(soff, o);
                        }
                    }
                    val[k] = slo[idx] + (o - soff[idx]);
                    pending &= ~(1u << k);
                }
            }
            if (last_window || cover >= o_end) break;
            __syncthreads();  // everybody is done with this window
            wbase += RG_CH;
            load_window();
            idx_prev = 0;
        }
#pragma unroll
        for (int k = 0; k < RG_TILE / RG_BLOCK; k += 2) {
            const int64_t e = e0 + (k >> 1) * (2 * RG_BLOCK) + 2 * tid;
            const bool a = e + out_begin < o_end, b = e + 1 + out_begin < o_end;
            if (VEC && b) {
                rg_st2(out + e, val[k], val[k + 1]);
            } else {
                if (a) hpgb_st1(out + e, val[k]);
                if (b) hpgb_st1(out + e + 1, val[k + 1]);
            }
        }
    }
}

}  // namespace

int launch_ranges_offsets(const int64_t *ranges, int64_t m, int inclusive, int64_t *offs, hpgb_status_t *st, cudaStream_t s) {
    if (m < 0 || !offs || (m > 0 && (!ranges || !st))) return HPGB_E_BAD_ARGUMENT;
    if (!hpgb_aligned16(ranges)) return HPGB_E_BAD_ARGUMENT;  // rows are read as 16-byte (lo, hi) pairs
    if (m == 0) {
        cudaError_t e = cudaMemsetAsync(offs, 0, 8, s);
        return e == cudaSuccess ? HPGB_OK : HPGB_E_CUDA + (int)e;
    }
    const int64_t ntile = (m + RG_TILE - 1) / RG_TILE;
    const longlong2 *r2 = reinterpret_cast<const longlong2 *>(ranges);
    k_ranges_tilesum<<<(unsigned)ntile, RG_BLOCK, 0, s>>>(r2, m, inclusive ? 1 : 0, offs, st);
    k_ranges_scan_tiles<<<1, 1024, 0, s>>>(offs, m, ntile);
    k_ranges_tilescan<<<(unsigned)ntile, RG_BLOCK, 0, s>>>(r2, m, inclusive ? 1 : 0, offs);
    g_hpgb_launches.fetch_add(3, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? HPGB_OK : HPGB_E_CUDA + (int)e;
}

int launch_ranges_expand(const int64_t *ranges, const int64_t *offs, int64_t m, int64_t out_begin, int64_t out_count,
                         int64_t *out, cudaStream_t s) {
    if (m < 0 || out_begin < 0 || out_count < 0) return HPGB_E_BAD_ARGUMENT;
    if (out_count == 0) return HPGB_OK;
    if (m == 0 || !ranges || !offs || !out || !hpgb_aligned16(ranges)) return HPGB_E_BAD_ARGUMENT;
    const int64_t ntiles = (out_count + RG_TILE - 1) / RG_TILE;
    static hpgb_occ_cache cache;
    const int occ = hpgb_resident_ctas_cached(cache, k_ranges_expand<true>, 0, RG_BLOCK);
    int64_t grid = (int64_t)hpgb_sm_count() * occ;
    if (grid > ntiles) grid = ntiles;
    const int64_t per = (ntiles + grid - 1) / grid;
    grid = (ntiles + per - 1) / per;
    const longlong2 *r2 = reinterpret_cast<const longlong2 *>(ranges);
    if (hpgb_aligned16(out))
        k_ranges_expand<true><<<(unsigned)grid, RG_BLOCK, 0, s>>>(r2, offs, m, out_begin, out_count, out, per);
    else
        k_ranges_expand<false><<<(unsigned)grid, RG_BLOCK, 0, s>>>(r2, offs, m, out_begin, out_count, out, per);
    g_hpgb_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? HPGB_OK : HPGB_E_CUDA + (int)e;
}

// ---------------------------------------------------------------------------------------------
// host-pointer entry points: the ranges (16 B each) go to the device once, the pixel list comes
// back through the chunked pipeline of hpgb_host.cu (expand kernel of chunk i+1 overlaps the D2H
// copy of chunk i).
// ---------------------------------------------------------------------------------------------
namespace {
struct DevRanges {
    int64_t *ranges = nullptr, *offs = nullptr;
    hpgb_status_t *st = nullptr;
    ~DevRanges() {
        if (ranges) hpgb_dev_free(ranges);
        if (offs) hpgb_dev_free(offs);
        if (st) hpgb_dev_free(st);
    }
};
#define CKR(call)                                            \
    do {                                                     \
        cudaError_t e_ = (call);                             \
        if (e_ != cudaSuccess) { cudaSetDevice(prev); return HPGB_E_CUDA + (int)e_; } \
    } while (0)

int stage_ranges(const int64_t *ranges, int64_t m, int inclusive, int device, DevRanges &d, int64_t *total, int64_t *first_bad) {
    if (first_bad) *first_bad = -1;
    *total = 0;
    if (m < 0 || (m > 0 && !ranges)) return HPGB_E_BAD_ARGUMENT;
    const int ndev = hpgb_device_count();
    if (ndev <= 0) return HPGB_E_NO_DEVICE;
    if (device < 0 || device >= ndev) return HPGB_E_BAD_ARGUMENT;
    if (m == 0) return HPGB_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    CKR(cudaSetDevice(device));
    CKR(hpgb_dev_alloc((void **)&d.ranges, (size_t)m * 16));
    CKR(hpgb_dev_alloc((void **)&d.offs, (size_t)(m + 1) * 8));
    CKR(hpgb_dev_alloc((void **)&d.st, sizeof(hpgb_status_t)));
    CKR(cudaMemcpyAsync(d.ranges, ranges, (size_t)m * 16, cudaMemcpyHostToDevice, 0));
    CKR(cudaMemsetAsync(d.st, 0xFF, sizeof(hpgb_status_t), 0));
    const int rc = launch_ranges_offsets(d.ranges, m, inclusive, d.offs, d.st, 0);
    if (rc != HPGB_OK) { cudaSetDevice(prev); return rc; }
    hpgb_status_t hs;
    CKR(cudaMemcpy(total, d.offs + m, 8, cudaMemcpyDeviceToHost));
    CKR(cudaMemcpy(&hs, d.st, sizeof hs, cudaMemcpyDeviceToHost));
    if (first_bad && hs.first_bad != HPGB_NO_BAD_ELEMENT) *first_bad = (int64_t)hs.first_bad;
    cudaSetDevice(prev);
    return HPGB_OK;
}
}  // namespace

extern "C" {

int hpgb_ranges_offsets(const int64_t *d_ranges_m2, int64_t m, int inclusive, int64_t *d_offsets_mp1,
                        hpgb_status_t *d_status, void *stream) {
    return launch_ranges_offsets(d_ranges_m2, m, inclusive, d_offsets_mp1, d_status, (cudaStream_t)stream);
}
int hpgb_ranges_expand(const int64_t *d_ranges_m2, const int64_t *d_offsets_mp1, int64_t m, int64_t out_begin,
                       int64_t out_count, int64_t *d_out, void *stream) {
    return launch_ranges_expand(d_ranges_m2, d_offsets_mp1, m, out_begin, out_count, d_out, (cudaStream_t)stream);
}
int hpgb_host_ranges_count(const int64_t *ranges_m2, int64_t m, int inclusive, int device, int64_t *total,
                           int64_t *first_bad) {
    device = hpgb_host_first_device(device);  // HPGB_DEVICE_ALL: this entry point's scratch lives on one GPU
    if (!total) return HPGB_E_BAD_ARGUMENT;
    DevRanges d;
    return stage_ranges(ranges_m2, m, inclusive, device, d, total, first_bad);
}
int hpgb_host_ranges_expand(const int64_t *ranges_m2, int64_t m, int inclusive, int64_t *out, int64_t total, int device) {
    device = hpgb_host_first_device(device);  // HPGB_DEVICE_ALL: this entry point's scratch lives on one GPU
    DevRanges d;
    int64_t tot = 0, bad = -1;
    int rc = stage_ranges(ranges_m2, m, inclusive, device, d, &tot, &bad);
    if (rc != HPGB_OK) return rc;
    if (bad >= 0 || tot != total || (total > 0 && !out)) return HPGB_E_BAD_ARGUMENT;
    if (total == 0) return HPGB_OK;
    hpgb_pipe_arr arrs[1] = {{nullptr, (char *)out, 8}};
    return hpgb_host_pipeline(device, total, arrs, 1,
                              [&](void **dp, int64_t base, int64_t cn, hpgb_status_t *, cudaStream_t s) {
                                  return launch_ranges_expand(d.ranges, d.offs, m, base, cn, (int64_t *)dp[0], s);
                              },
                              nullptr);
}

}  // extern "C"
