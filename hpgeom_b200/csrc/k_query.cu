// k_query.cu -- query_circle (reference: hpgeom/hpgeom.c:765-864 and query_disc, hpgeom/healpix_geom.c:541-800).
// The element functions (one ring, one pixel of the search tree) are in hpgb_query.h; every variant ends in
// an (M,2) table of [lo, hi) pixel ranges that the kernels of k_ranges.cu expand.  Four parts:
//   1. RING, one disc   : per-disc plan on the host, one GPU thread per ring, grid-wide ticket scan for the
//                         inclusive edge trimming;
//   2. NEST, one disc   : the hierarchical search walked breadth first, one launch per order (the small
//                         orders in one single-CTA launch), device radix sort of the ranges (cub: plumbing);
//   3. RING, many discs : one launch over all (disc, ring) pairs;
//   4. NEST, many discs : the search over (disc, pixel) frontiers, ranges ordered by (disc, lo).
// See include/hpgb.h for the C ABI.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <atomic>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "hpgb_core.h"
#include "hpgb_fp.h"
#include "hpgb_host.h"
#include "hpgb_internal.h"
#include "hpgb_launch.cuh"
#include "hpgb_query.h"

namespace {

// Inclusive mode: a ring whose edge scan (hpgb_qc_ring_trim) is not settled after QC_SEQ_STEPS checks
// per end becomes a work item; the whole grid then scans its candidates in parallel (k_qc_scan) and
// k_qc_finish writes its rows.  Such rings are few -- the ones that graze the enlarged disc, and the
// up-to-two guard rings that lie outside it altogether, where the reference walks the complete ring
// pixel by pixel (up to 4 nside pixels x 4 (fact - 1) edge samples, sequentially).
//   work[0] = number of items; item k = work[8 + 8k ..]: ring index j, ip_lo, ip_hi, unresolved-end
//   mask, found distance from the low / high end (INT64_MAX: none), ticket counters low / high.
#define QC_SEQ_STEPS 24
#define QC_ITEM(work, k) ((work) + 8 + 8 * (k))

__global__ void __launch_bounds__(HPGB_BLOCK) k_qc_ring_ranges(const hpgb_disc_t d, const hpgb_geom g, const hpgb_geom g2,
                                                               int64_t *__restrict__ ranges, int64_t *__restrict__ work) {
    const double *T = hpgb_stage_table();
    const int64_t nring = d.irmax - d.irmin + 1;
    const int64_t stride = (int64_t)gridDim.x * HPGB_BLOCK;
    for (int64_t j = (int64_t)blockIdx.x * HPGB_BLOCK + threadIdx.x; j < nring; j += stride) {
        int64_t row[4] = {0, 0, 0, 0};
        int64_t ip_lo, ip_hi, nr, ipix1;
        if (hpgb_qc_ring_raw(d, g, T, d.irmin + j, ip_lo, ip_hi, nr, ipix1)) {
            unsigned mask = 0u;
            if (d.fct > 1) mask = hpgb_qc_ring_trim(d, g, g2, T, nr, ipix1, ip_lo, ip_hi, QC_SEQ_STEPS);
            if (mask) {
                const unsigned long long k = atomicAdd(reinterpret_cast<unsigned long long *>(work), 1ull);
                int64_t *it = QC_ITEM(work, (int64_t)k);
                it[0] = j;
                it[1] = ip_lo;
                it[2] = ip_hi;
                it[3] = (int64_t)mask;
                it[4] = it[5] = INT64_MAX;
                it[6] = it[7] = 0;
            } else {
                hpgb_qc_ring_rows(ip_lo, ip_hi, nr, ipix1, row);
            }
        }
        longlong2 *o = reinterpret_cast<longlong2 *>(ranges) + 1 + 2 * j;
        o[0] = make_longlong2(row[0], row[1]);
        o[1] = make_longlong2(row[2], row[3]);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        longlong2 *o = reinterpret_cast<longlong2 *>(ranges);
        o[0] = make_longlong2(0, d.north_hi > 0 ? d.north_hi : 0);
        o[d.m - 1] = d.south_lo >= 0 ? make_longlong2(d.south_lo, g.npix) : make_longlong2(0, 0);
    }
}

// END = 0: nearest non-trimmable pixel from the low end of every item; END = 1 (second launch): from
// the high end, down to the pixel the low scan stopped at.  CTAs draw chunks of HPGB_BLOCK candidates
// in order of distance from the end (ticket counter) and stop once a chunk starts beyond the best
// distance found so far, so the minimum is exact and nothing past it is evaluated twice.
template <int END>
__global__ void __launch_bounds__(HPGB_BLOCK) k_qc_scan(const hpgb_disc_t d, const hpgb_geom g, const hpgb_geom g2,
                                                        int64_t *__restrict__ work) {
    const double *T = hpgb_stage_table();
    __shared__ int64_t s_ticket, s_found;
    const int64_t nitems = work[0];
    for (int64_t k = 0; k < nitems; k++) {
        int64_t *it = QC_ITEM(work, k);
        const unsigned mask = (unsigned)it[3];
        if (!(mask & (1u << END))) continue;
        int64_t lo = it[1];
        const int64_t hi = it[2];
        if (END == 1 && (mask & 1u)) {
            if (it[4] == INT64_MAX) continue;  // nothing overlaps: the ring is empty
            lo += it[4];
        }
        const int64_t len = hi - lo + 1;
        int64_t nr, ipix1;
        bool shifted;
        hpgb_ring_info_small(g, d.irmin + it[0], ipix1, nr, shifted);
        for (;;) {
            __syncthreads();
            if (threadIdx.x == 0) {
                s_ticket = (int64_t)atomicAdd(reinterpret_cast<unsigned long long *>(it + 6 + END), 1ull);
                s_found = *reinterpret_cast<volatile int64_t *>(it + 4 + END);
            }
            __syncthreads();
            const int64_t dist0 = s_ticket * HPGB_BLOCK;
            if (dist0 >= len || dist0 > s_found) break;
            const int64_t dist = dist0 + threadIdx.x;
            if (dist < len) {
                const int64_t ip = END ? hi - dist : lo + dist;
                if (!hpgb_check_pixel_ring(g, g2, T, ip, nr, ipix1, d.fct, d.z0, d.phi, d.cosrsmall, d.cpix))
                    atomicMin(reinterpret_cast<long long *>(it + 4 + END), (long long)dist);
            }
        }
    }
}

__global__ void __launch_bounds__(HPGB_BLOCK) k_qc_finish(const hpgb_disc_t d, const hpgb_geom g, const int64_t *__restrict__ work,
                                                          int64_t *__restrict__ ranges) {
    const int64_t nitems = work[0];
    const int64_t stride = (int64_t)gridDim.x * HPGB_BLOCK;
    for (int64_t k = (int64_t)blockIdx.x * HPGB_BLOCK + threadIdx.x; k < nitems; k += stride) {
        const int64_t *it = QC_ITEM(work, k);
        const unsigned mask = (unsigned)it[3];
        int64_t ip_lo = it[1], ip_hi = it[2];
        if (mask & 1u) ip_lo = it[4] == INT64_MAX ? ip_hi + 1 : ip_lo + it[4];
        if (ip_lo <= ip_hi) ip_hi = it[5] == INT64_MAX ? ip_lo : ip_hi - it[5];
        int64_t nr, ipix1, row[4];
        bool shifted;
        hpgb_ring_info_small(g, d.irmin + it[0], ipix1, nr, shifted);
        hpgb_qc_ring_rows(ip_lo, ip_hi, nr, ipix1, row);
        longlong2 *o = reinterpret_cast<longlong2 *>(ranges) + 1 + 2 * it[0];
        o[0] = make_longlong2(row[0], row[1]);
        o[1] = make_longlong2(row[2], row[3]);
    }
}

__global__ void k_qc_whole(int64_t *ranges, int64_t npix) {
    ranges[0] = 0;
    ranges[1] = npix;
}

}  // namespace

static int64_t qc_work_bytes(const hpgb_disc_t *plan) {
    if (!plan || plan->whole_sphere || plan->fct <= 1) return 0;
    const int64_t nring = plan->irmax >= plan->irmin ? plan->irmax - plan->irmin + 1 : 0;
    return 64 * (nring + 1);
}

int launch_qc_ring_ranges(const hpgb_disc_t *plan, int64_t *ranges, int64_t *work, cudaStream_t s) {
    if (!plan || !ranges || plan->m < 1 || !hpgb_aligned16(ranges)) return HPGB_E_BAD_ARGUMENT;
    const int c = hpgb_check_nside(plan->nside, 0);
    if (c) return c;
    const hpgb_geom g = hpgb_make_geom(plan->nside, 0);
    int launches = 1;
    if (plan->whole_sphere) {
        k_qc_whole<<<1, 1, 0, s>>>(ranges, g.npix);
    } else {
        const bool trim = qc_work_bytes(plan) > 0;
        if (trim && (!work || !hpgb_aligned16(work))) return HPGB_E_BAD_ARGUMENT;
        const hpgb_geom g2 = hpgb_make_geom(plan->nside * plan->fct, 0);
        const int64_t nring = plan->irmax - plan->irmin + 1;
        int64_t grid = (nring + HPGB_BLOCK - 1) / HPGB_BLOCK;
        const int64_t cap = (int64_t)hpgb_sm_count() * 8;
        if (grid > cap) grid = cap;
        if (grid < 1) grid = 1;
        if (trim) {
            cudaError_t e = cudaMemsetAsync(work, 0, 64, s);
            if (e != cudaSuccess) return HPGB_E_CUDA + (int)e;
        }
        k_qc_ring_ranges<<<(unsigned)grid, HPGB_BLOCK, 0, s>>>(*plan, g, g2, ranges, work);
        if (trim && nring > 0) {
            const unsigned wide = (unsigned)hpgb_sm_count() * 4u;
            k_qc_scan<0><<<wide, HPGB_BLOCK, 0, s>>>(*plan, g, g2, work);
            k_qc_scan<1><<<wide, HPGB_BLOCK, 0, s>>>(*plan, g, g2, work);
            k_qc_finish<<<(unsigned)(grid < 64 ? grid : 64), HPGB_BLOCK, 0, s>>>(*plan, g, work, ranges);
            launches += 3;
        }
    }
    g_hpgb_launches.fetch_add(launches, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? HPGB_OK : HPGB_E_CUDA + (int)e;
}

extern "C" {

int hpgb_query_circle_ring_plan(int64_t nside, double a, double b, double radius, int inclusive, int64_t fact,
                                int lonlat, int degrees, hpgb_disc_t *p) {
    // argument checks in the wrapper's order: hpgeom/hpgeom.c:793-838
    double theta, phi;
    bool ok;
    if (lonlat) {
        ok = degrees ? hpgb_angles_exact<true, true>(a, b, theta, phi) : hpgb_angles_exact<true, false>(a, b, theta, phi);
        if (ok && degrees) radius *= HPGB_PI / 180.0;
    } else {
        ok = hpgb_angles_exact<false, false>(a, b, theta, phi);
    }
    if (!ok) return HPGB_E_BAD_ANGLE;
    if (radius <= 0) return HPGB_E_BAD_RADIUS;
    const int c = hpgb_check_nside(nside, 0);
    if (c) return c;
    if (!inclusive) fact = 0;
    else if (fact <= 0 || fact * nside > ((int64_t)1 << 29)) return HPGB_E_BAD_FACT;
    return hpgb_qc_plan(nside, theta, phi, radius, fact, p);
}

int64_t hpgb_query_circle_ring_work_bytes(const hpgb_disc_t *plan) { return qc_work_bytes(plan); }

int hpgb_query_circle_ring_ranges(const hpgb_disc_t *plan, int64_t *d_ranges_m2, void *d_work, void *stream) {
    return launch_qc_ring_ranges(plan, d_ranges_m2, (int64_t *)d_work, (cudaStream_t)stream);
}

}  // extern "C"

// ---- host-pointer forms: table + offsets stay on the device, pixels come back chunked ----
namespace {
struct DevTable {
    int64_t *ranges = nullptr, *offs = nullptr, *work = nullptr;
    hpgb_status_t *st = nullptr;
    ~DevTable() {
        if (work) hpgb_dev_free(work);
        if (ranges) hpgb_dev_free(ranges);
        if (offs) hpgb_dev_free(offs);
        if (st) hpgb_dev_free(st);
    }
};
#define CKQ(call)                                                                      \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) { cudaSetDevice(prev); return HPGB_E_CUDA + (int)e_; } \
    } while (0)

int build_table(const hpgb_disc_t *plan, int device, DevTable &d, int64_t *total) {
    *total = 0;
    if (!plan || plan->m < 1) return HPGB_E_BAD_ARGUMENT;
    const int ndev = hpgb_device_count();
    if (ndev <= 0) return HPGB_E_NO_DEVICE;
    if (device < 0 || device >= ndev) return HPGB_E_BAD_ARGUMENT;
    int prev = 0;
    cudaGetDevice(&prev);
    CKQ(cudaSetDevice(device));
    CKQ(hpgb_dev_alloc((void **)&d.ranges, (size_t)plan->m * 16));
    CKQ(hpgb_dev_alloc((void **)&d.offs, (size_t)(plan->m + 1) * 8));
    CKQ(hpgb_dev_alloc((void **)&d.st, sizeof(hpgb_status_t)));
    CKQ(cudaMemsetAsync(d.st, 0xFF, sizeof(hpgb_status_t), 0));
    if (qc_work_bytes(plan) > 0) CKQ(hpgb_dev_alloc((void **)&d.work, (size_t)qc_work_bytes(plan)));
    int rc = launch_qc_ring_ranges(plan, d.ranges, d.work, 0);
    if (rc == HPGB_OK) rc = launch_ranges_offsets(d.ranges, plan->m, 0, d.offs, d.st, 0);
    if (rc != HPGB_OK) {
        cudaSetDevice(prev);
        return rc;
    }
    CKQ(cudaMemcpy(total, d.offs + plan->m, 8, cudaMemcpyDeviceToHost));
    cudaSetDevice(prev);
    return HPGB_OK;
}
}  // namespace

extern "C" {

int hpgb_host_query_circle_ring_count(const hpgb_disc_t *plan, int device, int64_t *total) {
    device = hpgb_host_first_device(device);  // HPGB_DEVICE_ALL: this entry point's scratch lives on one GPU
    if (!total) return HPGB_E_BAD_ARGUMENT;
    DevTable d;
    return build_table(plan, device, d, total);
}

int hpgb_host_query_circle_ring_fill(const hpgb_disc_t *plan, int64_t *out, int64_t total, int device) {
    device = hpgb_host_first_device(device);  // HPGB_DEVICE_ALL: this entry point's scratch lives on one GPU
    DevTable d;
    int64_t tot = 0;
    const int rc = build_table(plan, device, d, &tot);
    if (rc != HPGB_OK) return rc;
    if (tot != total || (total > 0 && !out)) return HPGB_E_BAD_ARGUMENT;
    if (total == 0) return HPGB_OK;
    const int64_t m = plan->m;
    hpgb_pipe_arr arrs[1] = {{nullptr, (char *)out, 8}};
    return hpgb_host_pipeline(device, total, arrs, 1,
                              [&](void **dp, int64_t base, int64_t cn, hpgb_status_t *, cudaStream_t s) {
                                  return launch_ranges_expand(d.ranges, d.offs, m, base, cn, (int64_t *)dp[0], s);
                              },
                              nullptr);
}

}  // extern "C"

// =============================================================================================
// NEST scheme: breadth-first hierarchical search (hpgb_qn_visit, hpgb_query.h), one launch per order.
// =============================================================================================
namespace {

// one order of the walk: verdict per frontier pixel, warp-aggregated appends to the next frontier
// (4 children per kept pixel) and to this order's range list.  cnt[0] = children written, cnt[1] = ranges.
__global__ void __launch_bounds__(HPGB_BLOCK) k_qn_level(const hpgb_qn_plan P, const hpgb_geom g, const int o,
                                                         const int64_t *__restrict__ frontier, const int64_t n,
                                                         int64_t *__restrict__ next, int64_t *__restrict__ rlo,
                                                         int64_t *__restrict__ rhi, unsigned long long *cnt) {
    const double *T = hpgb_stage_table();
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t stride = (int64_t)gridDim.x * HPGB_BLOCK;
    const int64_t rounds = (n + stride - 1) / stride;
    for (int64_t k = 0; k < rounds; k++) {
        const int64_t i = k * stride + (int64_t)blockIdx.x * HPGB_BLOCK + threadIdx.x;
        int v = 0;
        int64_t pix = 0, lo = 0, hi = 0;
        if (i < n) {
            pix = frontier[i];
            v = hpgb_qn_visit(P, g, T, o, pix, lo, hi);
        }
        const unsigned m1 = __ballot_sync(0xFFFFFFFFu, v == 1), m2 = __ballot_sync(0xFFFFFFFFu, v == 2);
        unsigned long long b1 = 0, b2 = 0;
        if (lane == 0) {
            if (m1) b1 = atomicAdd(cnt + 1, (unsigned long long)__popc(m1));
            if (m2) b2 = atomicAdd(cnt, 4ull * (unsigned long long)__popc(m2));
        }
        b1 = __shfl_sync(0xFFFFFFFFu, b1, 0);
        b2 = __shfl_sync(0xFFFFFFFFu, b2, 0);
        if (v == 1) {
            const unsigned long long r = b1 + (unsigned long long)__popc(m1 & lt);
            rlo[r] = lo;
            rhi[r] = hi;
        } else if (v == 2) {
            longlong2 *c = reinterpret_cast<longlong2 *>(next + b2 + 4ull * (unsigned long long)__popc(m2 & lt));
            c[0] = make_longlong2(4 * pix, 4 * pix + 1);
            c[1] = make_longlong2(4 * pix + 2, 4 * pix + 3);
        }
    }
}

// The first orders of the walk have a handful of pixels (12, then at most 4x per order): one launch + one
// host round trip per order costs more than the work.  This single-CTA kernel walks orders 0, 1, ... for as long
// as the next frontier is sure to fit QN_SMALL_CAP entries (and the ranges QN_SMALL_RCAP), ping-ponging between
// two global buffers; it leaves the frontier in bufA and {next order, frontier size, ranges} in state[0..2].
// A small query (a few thousand pixels) finishes here.
#define QN_SMALL_CAP 16384
#define QN_SMALL_RCAP 32768
__global__ void __launch_bounds__(HPGB_BLOCK) k_qn_small(const hpgb_qn_plan P, int64_t *__restrict__ bufA, int64_t *__restrict__ bufB,
                                                         int64_t *__restrict__ rlo, int64_t *__restrict__ rhi, int64_t *__restrict__ state) {
    const double *T = hpgb_stage_table();
    __shared__ int s_next, s_r;
    int64_t *src = bufA, *dst = bufB;
    if (threadIdx.x < 12) src[threadIdx.x] = threadIdx.x;
    if (threadIdx.x == 0) s_next = 0, s_r = 0;
    __syncthreads();
    int n = 12, o = 0, r = 0;
    while (o <= P.omax && n > 0 && 4 * n <= QN_SMALL_CAP && r + n <= QN_SMALL_RCAP) {
        const hpgb_geom g = hpgb_make_geom((int64_t)1 << o, 1);
        for (int i = threadIdx.x; i < n; i += HPGB_BLOCK) {
            const int64_t pix = src[i];
            int64_t lo, hi;
            const int v = hpgb_qn_visit(P, g, T, o, pix, lo, hi);
            if (v == 1) {
                const int k = atomicAdd(&s_r, 1);
                rlo[k] = lo;
                rhi[k] = hi;
            } else if (v == 2) {
                const int k = atomicAdd(&s_next, 4);
                dst[k] = 4 * pix, dst[k + 1] = 4 * pix + 1, dst[k + 2] = 4 * pix + 2, dst[k + 3] = 4 * pix + 3;
            }
        }
        __syncthreads();
        n = s_next;
        r = s_r;
        __syncthreads();
        if (threadIdx.x == 0) s_next = 0;
        int64_t *t = src;
        src = dst;
        dst = t;
        o++;
        __syncthreads();
    }
    if (src != bufA)
        for (int i = threadIdx.x; i < n; i += HPGB_BLOCK) bufA[i] = src[i];
    if (threadIdx.x == 0) state[0] = o, state[1] = n, state[2] = r;
}

// sorted (lo, hi) -> (m,2) table; a repeated single pixel (see hpgb_query.h) becomes an empty row
__global__ void __launch_bounds__(HPGB_BLOCK) k_qn_table(const int64_t *__restrict__ lo, const int64_t *__restrict__ hi, int64_t m,
                                                         longlong2 *__restrict__ table) {
    const int64_t stride = (int64_t)gridDim.x * HPGB_BLOCK;
    for (int64_t i = (int64_t)blockIdx.x * HPGB_BLOCK + threadIdx.x; i < m; i += stride) {
        const int64_t a = lo[i];
        const bool dup = i > 0 && lo[i - 1] == a;
        table[i] = make_longlong2(a, dup ? a : hi[i]);
    }
}

struct QnHandle {
    int device = 0;
    int64_t m = 0, total = 0;
    int64_t *ranges = nullptr, *offs = nullptr;
    int64_t n_discs = 0;            // batched search: number of discs and
    int64_t *disc_offs = nullptr;   // the n_discs + 1 boundaries of their pixel lists (device)
};

struct DevBuf {  // frees on scope exit
    std::vector<void *> ptrs;
    ~DevBuf() {
        for (void *p : ptrs) hpgb_dev_free(p);
    }
    template <class Tp>
    cudaError_t alloc(Tp **p, size_t bytes) {
        cudaError_t e = hpgb_dev_alloc((void **)p, bytes);
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
    void release(void *p) {
        for (auto &q : ptrs)
            if (q == p) q = nullptr;
    }
};

int qn_build(const hpgb_qn_plan &P, int64_t nside, QnHandle *h) {
    int prev = 0;
    cudaGetDevice(&prev);
    CKQ(cudaSetDevice(h->device));
    DevBuf pool;
    const int64_t npix = 12 * nside * nside;
    int64_t *keys = nullptr, *vals = nullptr;  // unsorted ranges
    int64_t R = 0;
    if (P.whole_sphere) {
        R = 1;
        CKQ(pool.alloc(&keys, 8));
        CKQ(pool.alloc(&vals, 8));
        const int64_t z = 0;
        CKQ(cudaMemcpy(keys, &z, 8, cudaMemcpyHostToDevice));
        CKQ(cudaMemcpy(vals, &npix, 8, cudaMemcpyHostToDevice));
    } else {
        unsigned long long *cnt = nullptr;
        CKQ(pool.alloc(&cnt, 16));
        // orders with small frontiers: one single-CTA launch (k_qn_small)
        int64_t *cur = nullptr, *bufB = nullptr, *rlo0 = nullptr, *rhi0 = nullptr, *state = nullptr;
        CKQ(pool.alloc(&cur, (size_t)QN_SMALL_CAP * 8));
        CKQ(pool.alloc(&bufB, (size_t)QN_SMALL_CAP * 8));
        CKQ(pool.alloc(&rlo0, (size_t)QN_SMALL_RCAP * 8));
        CKQ(pool.alloc(&rhi0, (size_t)QN_SMALL_RCAP * 8));
        CKQ(pool.alloc(&state, 32));
        k_qn_small<<<1, HPGB_BLOCK, 0, 0>>>(P, cur, bufB, rlo0, rhi0, state);
        g_hpgb_launches.fetch_add(1, std::memory_order_relaxed);
        int64_t hs[3];
        CKQ(cudaMemcpy(hs, state, 24, cudaMemcpyDeviceToHost));
        int64_t n = hs[1];
        std::vector<int64_t *> lvl_lo, lvl_hi;
        std::vector<int64_t> lvl_n;
        lvl_lo.push_back(rlo0);
        lvl_hi.push_back(rhi0);
        lvl_n.push_back(hs[2]);
        R += hs[2];
        for (int o = (int)hs[0]; o <= P.omax && n > 0; o++) {
            int64_t *next = nullptr, *rlo = nullptr, *rhi = nullptr;
            CKQ(pool.alloc(&next, (size_t)n * 32));
            CKQ(pool.alloc(&rlo, (size_t)n * 8));
            CKQ(pool.alloc(&rhi, (size_t)n * 8));
            CKQ(cudaMemsetAsync(cnt, 0, 16, 0));
            const hpgb_geom g = hpgb_make_geom((int64_t)1 << o, 1);
            int64_t grid = (n + HPGB_BLOCK - 1) / HPGB_BLOCK;
            const int64_t cap = (int64_t)hpgb_sm_count() * 8;
            if (grid > cap) grid = cap;
            k_qn_level<<<(unsigned)grid, HPGB_BLOCK, 0, 0>>>(P, g, o, cur, n, next, rlo, rhi, cnt);
            g_hpgb_launches.fetch_add(1, std::memory_order_relaxed);
            unsigned long long hc[2];
            CKQ(cudaMemcpy(hc, cnt, 16, cudaMemcpyDeviceToHost));
            lvl_lo.push_back(rlo);
            lvl_hi.push_back(rhi);
            lvl_n.push_back((int64_t)hc[1]);
            R += (int64_t)hc[1];
            hpgb_dev_free(cur);
            pool.release(cur);
            cur = next;
            n = (int64_t)hc[0];
        }
        CKQ(pool.alloc(&keys, (size_t)R * 8));
        CKQ(pool.alloc(&vals, (size_t)R * 8));
        int64_t at = 0;
        for (size_t l = 0; l < lvl_n.size(); l++) {
            if (lvl_n[l] > 0) {
                CKQ(cudaMemcpyAsync(keys + at, lvl_lo[l], (size_t)lvl_n[l] * 8, cudaMemcpyDeviceToDevice, 0));
                CKQ(cudaMemcpyAsync(vals + at, lvl_hi[l], (size_t)lvl_n[l] * 8, cudaMemcpyDeviceToDevice, 0));
            }
            at += lvl_n[l];
        }
    }
    h->m = R;
    h->total = 0;
    if (R > 0) {
        int64_t *skeys = nullptr, *svals = nullptr;
        void *tmp = nullptr;
        size_t tmp_bytes = 0;
        CKQ(pool.alloc(&skeys, (size_t)R * 8));
        CKQ(pool.alloc(&svals, (size_t)R * 8));
        CKQ(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, skeys, vals, svals, R, 0, 62, (cudaStream_t)0));
        CKQ(pool.alloc(&tmp, tmp_bytes));
        CKQ(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, skeys, vals, svals, R, 0, 62, (cudaStream_t)0));
        CKQ(hpgb_dev_alloc((void **)&h->ranges, (size_t)R * 16));
        CKQ(hpgb_dev_alloc((void **)&h->offs, (size_t)(R + 1) * 8));
        int64_t grid = (R + HPGB_BLOCK - 1) / HPGB_BLOCK;
        const int64_t cap = (int64_t)hpgb_sm_count() * 8;
        if (grid > cap) grid = cap;
        k_qn_table<<<(unsigned)grid, HPGB_BLOCK, 0, 0>>>(skeys, svals, R, reinterpret_cast<longlong2 *>(h->ranges));
        g_hpgb_launches.fetch_add(2, std::memory_order_relaxed);
        hpgb_status_t *st = nullptr;
        CKQ(pool.alloc(&st, sizeof(hpgb_status_t)));
        CKQ(cudaMemsetAsync(st, 0xFF, sizeof(hpgb_status_t), 0));
        const int rc = launch_ranges_offsets(h->ranges, R, 0, h->offs, st, 0);
        if (rc != HPGB_OK) {
            cudaSetDevice(prev);
            return rc;
        }
        CKQ(cudaMemcpy(&h->total, h->offs + R, 8, cudaMemcpyDeviceToHost));
    }
    cudaError_t e = cudaDeviceSynchronize();
    cudaSetDevice(prev);
    return e == cudaSuccess ? HPGB_OK : HPGB_E_CUDA + (int)e;
}

}  // namespace

extern "C" {

int hpgb_query_circle_nest_open(int64_t nside, double a, double b, double radius, int inclusive, int64_t fact, int lonlat,
                                int degrees, int device, void **handle, int64_t *m, int64_t *total) {
    device = hpgb_host_first_device(device);  // HPGB_DEVICE_ALL: this entry point's scratch lives on one GPU
    if (!handle || !m || !total) return HPGB_E_BAD_ARGUMENT;
    *handle = nullptr;
    // argument checks in the wrapper's order: hpgeom/hpgeom.c:793-838
    double theta, phi;
    bool ok;
    if (lonlat) {
        ok = degrees ? hpgb_angles_exact<true, true>(a, b, theta, phi) : hpgb_angles_exact<true, false>(a, b, theta, phi);
        if (ok && degrees) radius *= HPGB_PI / 180.0;
    } else {
        ok = hpgb_angles_exact<false, false>(a, b, theta, phi);
    }
    if (!ok) return HPGB_E_BAD_ANGLE;
    if (radius <= 0) return HPGB_E_BAD_RADIUS;
    const int c = hpgb_check_nside(nside, 1);
    if (c) return c;
    if (!inclusive) fact = 0;
    else if (fact <= 0 || fact * nside > ((int64_t)1 << 29) || (fact & (fact - 1))) return HPGB_E_BAD_FACT;
    hpgb_qn_plan P;
    int rc = hpgb_qn_make_plan(nside, theta, phi, radius, fact, &P);
    if (rc != HPGB_OK) return rc;
    const int ndev = hpgb_device_count();
    if (ndev <= 0) return HPGB_E_NO_DEVICE;
    if (device < 0 || device >= ndev) return HPGB_E_BAD_ARGUMENT;
    QnHandle *h = new QnHandle();
    h->device = device;
    rc = qn_build(P, nside, h);
    if (rc != HPGB_OK) {
        if (h->ranges) hpgb_dev_free(h->ranges);
        if (h->offs) hpgb_dev_free(h->offs);
        delete h;
        return rc;
    }
    *handle = h;
    *m = h->m;
    *total = h->total;
    return HPGB_OK;
}

int hpgb_query_nest_close(void *handle) {
    QnHandle *h = (QnHandle *)handle;
    if (!h) return HPGB_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(h->device);
    if (h->ranges) hpgb_dev_free(h->ranges);
    if (h->offs) hpgb_dev_free(h->offs);
    if (h->disc_offs) hpgb_dev_free(h->disc_offs);
    cudaSetDevice(prev);
    delete h;
    return HPGB_OK;
}

// merged [lo, hi) ranges as the reference's range set holds them (touching ranges fused,
// hpgeom/hpgeom_stack.c:231-250): out_m2 has room for m rows, *n_merged rows are written
int hpgb_query_nest_ranges(void *handle, int64_t *out_m2, int64_t *n_merged) {
    QnHandle *h = (QnHandle *)handle;
    if (!h || !n_merged || (h->m > 0 && !out_m2)) return HPGB_E_BAD_ARGUMENT;
    *n_merged = 0;
    if (h->m == 0) return HPGB_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    CKQ(cudaSetDevice(h->device));
    CKQ(cudaMemcpy(out_m2, h->ranges, (size_t)h->m * 16, cudaMemcpyDeviceToHost));
    cudaSetDevice(prev);
    int64_t k = 0;  // in place: rows are sorted and disjoint; drop empty rows, fuse touching ones
    for (int64_t i = 0; i < h->m; i++) {
        const int64_t lo = out_m2[2 * i], hi = out_m2[2 * i + 1];
        if (hi <= lo) continue;
        if (k > 0 && lo <= out_m2[2 * k - 1]) {
            if (hi > out_m2[2 * k - 1]) out_m2[2 * k - 1] = hi;
        } else {
            out_m2[2 * k] = lo;
            out_m2[2 * k + 1] = hi;
            k++;
        }
    }
    *n_merged = k;
    return HPGB_OK;
}

int hpgb_query_nest_pixels_host(void *handle, int64_t *out, int64_t total) {
    QnHandle *h = (QnHandle *)handle;
    if (!h || total != h->total || (total > 0 && !out)) return HPGB_E_BAD_ARGUMENT;
    if (total == 0) return HPGB_OK;
    hpgb_pipe_arr arrs[1] = {{nullptr, (char *)out, 8}};
    return hpgb_host_pipeline(h->device, total, arrs, 1,
                              [&](void **dp, int64_t base, int64_t cn, hpgb_status_t *, cudaStream_t s) {
                                  return launch_ranges_expand(h->ranges, h->offs, h->m, base, cn, (int64_t *)dp[0], s);
                              },
                              nullptr);
}

int hpgb_query_nest_pixels_device(void *handle, int64_t *d_out, int64_t total, void *stream) {
    QnHandle *h = (QnHandle *)handle;
    if (!h || total != h->total || (total > 0 && !d_out)) return HPGB_E_BAD_ARGUMENT;
    return launch_ranges_expand(h->ranges, h->offs, h->m, 0, total, d_out, (cudaStream_t)stream);
}

}  // extern "C"

// =============================================================================================
// Many discs at once, RING scheme ("a batched many-discs kernel").  One small disc
// is launch-latency bound on a GPU; a batch is not: the D per-disc plans (same nside and fact) go to the
// device, ONE launch runs every (disc, ring) pair -- a thread finds its disc by binary search over the
// prefix sums of the ring counts -- and one offsets + expand pass writes all pixel lists back to back.
// Row layout: disc d owns rows [2 ring_start[d] + 2 d, 2 ring_start[d+1] + 2 (d+1)): north polar row, two
// rows per ring, south polar row -- i.e. D single-disc tables concatenated.
// =============================================================================================
namespace {

__device__ __forceinline__ int64_t qb_find_disc(const int64_t *__restrict__ ring_start, int64_t D, int64_t j) {
    int64_t lo = 0, hi = D;  // largest d < D with ring_start[d] <= j
    while (hi - lo > 1) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if (ring_start[mid] <= j) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(HPGB_BLOCK) k_qb_ring_ranges(const hpgb_disc_t *__restrict__ plans,
                                                               const int64_t *__restrict__ ring_start, const int64_t D,
                                                               const hpgb_geom g, const hpgb_geom g2,
                                                               int64_t *__restrict__ ranges, int64_t *__restrict__ work) {
    const double *T = hpgb_stage_table();
    const int64_t nring = ring_start[D];
    const int64_t stride = (int64_t)gridDim.x * HPGB_BLOCK;
    const int64_t t0 = (int64_t)blockIdx.x * HPGB_BLOCK + threadIdx.x;
    for (int64_t j = t0; j < nring; j += stride) {
        const int64_t dsc = qb_find_disc(ring_start, D, j);
        const hpgb_disc_t d = plans[dsc];
        int64_t row[4] = {0, 0, 0, 0};
        int64_t ip_lo, ip_hi, nr, ipix1;
        if (hpgb_qc_ring_raw(d, g, T, d.irmin + (j - ring_start[dsc]), ip_lo, ip_hi, nr, ipix1)) {
            unsigned mask = 0u;
            if (d.fct > 1) mask = hpgb_qc_ring_trim(d, g, g2, T, nr, ipix1, ip_lo, ip_hi, QC_SEQ_STEPS);
            if (mask) {
                const unsigned long long k = atomicAdd(reinterpret_cast<unsigned long long *>(work), 1ull);
                int64_t *it = QC_ITEM(work, (int64_t)k);
                it[0] = j;
                it[1] = ip_lo;
                it[2] = ip_hi;
                it[3] = (int64_t)mask;
                it[4] = it[5] = INT64_MAX;
                it[6] = it[7] = 0;
            } else {
                hpgb_qc_ring_rows(ip_lo, ip_hi, nr, ipix1, row);
            }
        }
        longlong2 *o = reinterpret_cast<longlong2 *>(ranges) + 2 * j + 2 * dsc + 1;
        o[0] = make_longlong2(row[0], row[1]);
        o[1] = make_longlong2(row[2], row[3]);
    }
    for (int64_t dsc = t0; dsc < D; dsc += stride) {
        const hpgb_disc_t d = plans[dsc];
        longlong2 *o = reinterpret_cast<longlong2 *>(ranges);
        o[2 * ring_start[dsc] + 2 * dsc] = make_longlong2(0, d.north_hi > 0 ? d.north_hi : 0);
        o[2 * ring_start[dsc + 1] + 2 * dsc + 1] = d.south_lo >= 0 ? make_longlong2(d.south_lo, g.npix) : make_longlong2(0, 0);
    }
}

template <int END>
__global__ void __launch_bounds__(HPGB_BLOCK) k_qb_scan(const hpgb_disc_t *__restrict__ plans, const int64_t *__restrict__ ring_start,
                                                        const int64_t D, const hpgb_geom g, const hpgb_geom g2,
                                                        int64_t *__restrict__ work) {
    const double *T = hpgb_stage_table();
    __shared__ int64_t s_ticket, s_found;
    const int64_t nitems = work[0];
    for (int64_t k = 0; k < nitems; k++) {
        int64_t *it = QC_ITEM(work, k);
        const unsigned mask = (unsigned)it[3];
        if (!(mask & (1u << END))) continue;
        int64_t lo = it[1];
        const int64_t hi = it[2];
        if (END == 1 && (mask & 1u)) {
            if (it[4] == INT64_MAX) continue;
            lo += it[4];
        }
        const int64_t len = hi - lo + 1;
        const int64_t dsc = qb_find_disc(ring_start, D, it[0]);
        const hpgb_disc_t d = plans[dsc];
        int64_t nr, ipix1;
        bool shifted;
        hpgb_ring_info_small(g, d.irmin + (it[0] - ring_start[dsc]), ipix1, nr, shifted);
        for (;;) {
            __syncthreads();
            if (threadIdx.x == 0) {
                s_ticket = (int64_t)atomicAdd(reinterpret_cast<unsigned long long *>(it + 6 + END), 1ull);
                s_found = *reinterpret_cast<volatile int64_t *>(it + 4 + END);
            }
            __syncthreads();
            const int64_t dist0 = s_ticket * HPGB_BLOCK;
            if (dist0 >= len || dist0 > s_found) break;
            const int64_t dist = dist0 + threadIdx.x;
            if (dist < len) {
                const int64_t ip = END ? hi - dist : lo + dist;
                if (!hpgb_check_pixel_ring(g, g2, T, ip, nr, ipix1, d.fct, d.z0, d.phi, d.cosrsmall, d.cpix))
                    atomicMin(reinterpret_cast<long long *>(it + 4 + END), (long long)dist);
            }
        }
    }
}

__global__ void __launch_bounds__(HPGB_BLOCK) k_qb_finish(const hpgb_disc_t *__restrict__ plans, const int64_t *__restrict__ ring_start,
                                                          const int64_t D, const hpgb_geom g, const int64_t *__restrict__ work,
                                                          int64_t *__restrict__ ranges) {
    const int64_t nitems = work[0];
    const int64_t stride = (int64_t)gridDim.x * HPGB_BLOCK;
    for (int64_t k = (int64_t)blockIdx.x * HPGB_BLOCK + threadIdx.x; k < nitems; k += stride) {
        const int64_t *it = QC_ITEM(work, k);
        const unsigned mask = (unsigned)it[3];
        int64_t ip_lo = it[1], ip_hi = it[2];
        if (mask & 1u) ip_lo = it[4] == INT64_MAX ? ip_hi + 1 : ip_lo + it[4];
        if (ip_lo <= ip_hi) ip_hi = it[5] == INT64_MAX ? ip_lo : ip_hi - it[5];
        const int64_t dsc = qb_find_disc(ring_start, D, it[0]);
        int64_t nr, ipix1, row[4];
        bool shifted;
        hpgb_ring_info_small(g, plans[dsc].irmin + (it[0] - ring_start[dsc]), ipix1, nr, shifted);
        hpgb_qc_ring_rows(ip_lo, ip_hi, nr, ipix1, row);
        longlong2 *o = reinterpret_cast<longlong2 *>(ranges) + 2 * it[0] + 2 * dsc + 1;
        o[0] = make_longlong2(row[0], row[1]);
        o[1] = make_longlong2(row[2], row[3]);
    }
}

// pixels of disc d = offs[row_start[d+1]] - offs[row_start[d]]: the D + 1 list boundaries
__global__ void __launch_bounds__(HPGB_BLOCK) k_qb_disc_offsets(const int64_t *__restrict__ ring_start, const int64_t D,
                                                                const int64_t *__restrict__ offs, int64_t *__restrict__ disc_offs) {
    const int64_t stride = (int64_t)gridDim.x * HPGB_BLOCK;
    for (int64_t d = (int64_t)blockIdx.x * HPGB_BLOCK + threadIdx.x; d <= D; d += stride) disc_offs[d] = offs[2 * ring_start[d] + 2 * d];
}

}  // namespace

extern "C" {

// d_plans: D plans of hpgb_query_circle_ring_plan (same nside and fct) copied to the device; d_ring_start:
// D + 1 prefix sums of max(irmax - irmin + 1, 0) (whole-sphere plans count 0 rings and carry north_hi = npix:
// hpgb_query_circle_ring_batch_prepare does both).  d_ranges: 2 * ring_start[D] + 2 * D rows.  d_work:
// 64 * (ring_start[D] + 1) bytes when fct > 1, else NULL.
int hpgb_query_circle_ring_batch_ranges(const hpgb_disc_t *d_plans, const int64_t *d_ring_start, int64_t D, int64_t total_rings,
                                        int64_t nside, int64_t fct, int64_t *d_ranges_m2, void *d_work, void *stream) {
    if (D < 0 || total_rings < 0 || (D > 0 && (!d_plans || !d_ring_start || !d_ranges_m2))) return HPGB_E_BAD_ARGUMENT;
    if (D == 0) return HPGB_OK;
    const int c = hpgb_check_nside(nside, 0);
    if (c) return c;
    if (fct < 1 || fct * nside > ((int64_t)1 << 29) || !hpgb_aligned16(d_ranges_m2)) return HPGB_E_BAD_ARGUMENT;
    const bool trim = fct > 1;
    if (trim && (!d_work || !hpgb_aligned16(d_work))) return HPGB_E_BAD_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream;
    const hpgb_geom g = hpgb_make_geom(nside, 0), g2 = hpgb_make_geom(nside * fct, 0);
    const int64_t nthreads = total_rings > D ? total_rings : D;
    int64_t grid = (nthreads + HPGB_BLOCK - 1) / HPGB_BLOCK;
    const int64_t cap = (int64_t)hpgb_sm_count() * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    if (trim) {
        cudaError_t e = cudaMemsetAsync(d_work, 0, 64, s);
        if (e != cudaSuccess) return HPGB_E_CUDA + (int)e;
    }
    k_qb_ring_ranges<<<(unsigned)grid, HPGB_BLOCK, 0, s>>>(d_plans, d_ring_start, D, g, g2, d_ranges_m2, (int64_t *)d_work);
    int launches = 1;
    if (trim && total_rings > 0) {
        const unsigned wide = (unsigned)hpgb_sm_count() * 4u;
        k_qb_scan<0><<<wide, HPGB_BLOCK, 0, s>>>(d_plans, d_ring_start, D, g, g2, (int64_t *)d_work);
        k_qb_scan<1><<<wide, HPGB_BLOCK, 0, s>>>(d_plans, d_ring_start, D, g, g2, (int64_t *)d_work);
        k_qb_finish<<<(unsigned)(grid < 64 ? grid : 64), HPGB_BLOCK, 0, s>>>(d_plans, d_ring_start, D, g, (const int64_t *)d_work, d_ranges_m2);
        launches += 3;
    }
    g_hpgb_launches.fetch_add(launches, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? HPGB_OK : HPGB_E_CUDA + (int)e;
}

// plans for D discs in one call (host); stops at the first disc whose arguments the reference would reject:
// returns its code and index
int hpgb_query_circle_ring_plan_many(int64_t nside, const double *a, const double *b, const double *radius, int64_t D,
                                     int inclusive, int64_t fact, int lonlat, int degrees, hpgb_disc_t *plans, int64_t *first_bad) {
    if (first_bad) *first_bad = -1;
    if (D < 0 || (D > 0 && (!a || !b || !radius || !plans))) return HPGB_E_BAD_ARGUMENT;
    for (int64_t d = 0; d < D; d++) {
        const int rc = hpgb_query_circle_ring_plan(nside, a[d], b[d], radius[d], inclusive, fact, lonlat, degrees, plans + d);
        if (rc != HPGB_OK) {
            if (first_bad) *first_bad = d;
            return rc;
        }
    }
    return HPGB_OK;
}

// host side of a batch: normalises the plans (whole-sphere discs become "no rings + north polar range
// (0, npix)") and fills ring_start[0..D]; returns the total number of rings
int64_t hpgb_query_circle_ring_batch_prepare(hpgb_disc_t *plans, int64_t D, int64_t *ring_start) {
    int64_t acc = 0;
    for (int64_t d = 0; d < D; d++) {
        hpgb_disc_t &p = plans[d];
        if (p.whole_sphere) {
            p.irmin = 1;
            p.irmax = 0;
            p.north_hi = 12 * p.nside * p.nside;
            p.south_lo = -1;
        }
        ring_start[d] = acc;
        acc += p.irmax >= p.irmin ? p.irmax - p.irmin + 1 : 0;
    }
    ring_start[D] = acc;
    return acc;
}

// Host form: every pixel list back to back in `out` (cap pixels), list d = out[disc_offsets[d] .. disc_offsets[d+1]).
// Call with out = NULL to learn the sizes (disc_offsets is filled, *total set), then with the buffer.
int hpgb_host_query_circle_ring_batch(hpgb_disc_t *plans, int64_t D, int device, int64_t *disc_offsets, int64_t *total,
                                      int64_t *out, int64_t cap) {
    device = hpgb_host_first_device(device);  // HPGB_DEVICE_ALL: this entry point's scratch lives on one GPU
    if (D < 0 || !total || (D > 0 && (!plans || !disc_offsets))) return HPGB_E_BAD_ARGUMENT;
    *total = 0;
    if (D == 0) {
        if (disc_offsets) disc_offsets[0] = 0;
        return HPGB_OK;
    }
    for (int64_t d = 1; d < D; d++)
        if (plans[d].nside != plans[0].nside || plans[d].fct != plans[0].fct) return HPGB_E_BAD_ARGUMENT;
    const int ndev = hpgb_device_count();
    if (ndev <= 0) return HPGB_E_NO_DEVICE;
    if (device < 0 || device >= ndev) return HPGB_E_BAD_ARGUMENT;
    std::vector<int64_t> ring_start((size_t)D + 1);
    const int64_t nring = hpgb_query_circle_ring_batch_prepare(plans, D, ring_start.data());
    const int64_t m = 2 * nring + 2 * D;
    int prev = 0;
    cudaGetDevice(&prev);
    CKQ(cudaSetDevice(device));
    DevBuf pool;
    hpgb_disc_t *d_plans = nullptr;
    int64_t *d_rs = nullptr, *d_ranges = nullptr, *d_offs = nullptr, *d_work = nullptr, *d_do = nullptr;
    hpgb_status_t *st = nullptr;
    CKQ(pool.alloc(&d_plans, (size_t)D * sizeof(hpgb_disc_t)));
    CKQ(pool.alloc(&d_rs, (size_t)(D + 1) * 8));
    CKQ(pool.alloc(&d_ranges, (size_t)m * 16));
    CKQ(pool.alloc(&d_offs, (size_t)(m + 1) * 8));
    CKQ(pool.alloc(&d_do, (size_t)(D + 1) * 8));
    CKQ(pool.alloc(&st, sizeof(hpgb_status_t)));
    if (plans[0].fct > 1) CKQ(pool.alloc(&d_work, (size_t)64 * (size_t)(nring + 1)));
    CKQ(cudaMemcpyAsync(d_plans, plans, (size_t)D * sizeof(hpgb_disc_t), cudaMemcpyHostToDevice, 0));
    CKQ(cudaMemcpyAsync(d_rs, ring_start.data(), (size_t)(D + 1) * 8, cudaMemcpyHostToDevice, 0));
    CKQ(cudaMemsetAsync(st, 0xFF, sizeof(hpgb_status_t), 0));
    int rc = hpgb_query_circle_ring_batch_ranges(d_plans, d_rs, D, nring, plans[0].nside, plans[0].fct, d_ranges, d_work, nullptr);
    if (rc == HPGB_OK) rc = launch_ranges_offsets(d_ranges, m, 0, d_offs, st, 0);
    if (rc != HPGB_OK) {
        cudaSetDevice(prev);
        return rc;
    }
    int64_t grid = (D + 1 + HPGB_BLOCK - 1) / HPGB_BLOCK;
    if (grid > 1024) grid = 1024;
    k_qb_disc_offsets<<<(unsigned)grid, HPGB_BLOCK, 0, 0>>>(d_rs, D, d_offs, d_do);
    g_hpgb_launches.fetch_add(1, std::memory_order_relaxed);
    CKQ(cudaMemcpy(disc_offsets, d_do, (size_t)(D + 1) * 8, cudaMemcpyDeviceToHost));
    *total = disc_offsets[D];
    cudaSetDevice(prev);
    if (!out) return HPGB_OK;
    if (cap < *total) return HPGB_E_BAD_ARGUMENT;
    if (*total == 0) return HPGB_OK;
    hpgb_pipe_arr arrs[1] = {{nullptr, (char *)out, 8}};
    return hpgb_host_pipeline(device, *total, arrs, 1,
                              [&](void **dp, int64_t base, int64_t cn, hpgb_status_t *, cudaStream_t s) {
                                  return launch_ranges_expand(d_ranges, d_offs, m, base, cn, (int64_t *)dp[0], s);
                              },
                              nullptr);
}

}  // extern "C"

// =============================================================================================
// Many discs at once, NEST scheme (extension): the breadth-first search of k_qn_level over a frontier of
// (disc, pixel) pairs -- every order is ONE launch for the whole batch -- then the ranges are ordered by
// (disc, lo) with two stable radix sorts and expanded in one pass.  All discs share nside and fact.
// =============================================================================================
namespace {

__global__ void __launch_bounds__(HPGB_BLOCK) k_qnb_seed(int64_t *__restrict__ fpix, int32_t *__restrict__ fdisc,
                                                         const int32_t *__restrict__ active, int64_t nactive) {
    const int64_t stride = (int64_t)gridDim.x * HPGB_BLOCK;
    for (int64_t i = (int64_t)blockIdx.x * HPGB_BLOCK + threadIdx.x; i < 12 * nactive; i += stride) {
        fpix[i] = i % 12;
        fdisc[i] = active[i / 12];
    }
}

__global__ void __launch_bounds__(HPGB_BLOCK) k_qnb_level(const hpgb_qn_plan *__restrict__ plans, const hpgb_geom g, const int o,
                                                          const int64_t *__restrict__ fpix, const int32_t *__restrict__ fdisc,
                                                          const int64_t n, int64_t *__restrict__ npix, int32_t *__restrict__ ndisc,
                                                          int64_t *__restrict__ rlo, int64_t *__restrict__ rhi,
                                                          int32_t *__restrict__ rdisc, unsigned long long *cnt) {
    const double *T = hpgb_stage_table();
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t stride = (int64_t)gridDim.x * HPGB_BLOCK;
    const int64_t rounds = (n + stride - 1) / stride;
    for (int64_t k = 0; k < rounds; k++) {
        const int64_t i = k * stride + (int64_t)blockIdx.x * HPGB_BLOCK + threadIdx.x;
        int v = 0;
        int32_t d = 0;
        int64_t pix = 0, lo = 0, hi = 0;
        if (i < n) {
            pix = fpix[i];
            d = fdisc[i];
            const hpgb_qn_plan *P = plans + d;
            // the fields of the plan this order needs, read through the pointer (520-byte plans stay in L2)
            hpgb_qn_plan q;
            q.order = P->order, q.omax = P->omax, q.inclusive = P->inclusive, q.whole_sphere = 0;
            q.ptg_z = P->ptg_z, q.ptg_phi = P->ptg_phi, q.cosrad = P->cosrad;
            q.crpdr[o] = P->crpdr[o], q.crmdr[o] = P->crmdr[o];
            v = hpgb_qn_visit(q, g, T, o, pix, lo, hi);
        }
        const unsigned m1 = __ballot_sync(0xFFFFFFFFu, v == 1), m2 = __ballot_sync(0xFFFFFFFFu, v == 2);
        unsigned long long b1 = 0, b2 = 0;
        if (lane == 0) {
            if (m1) b1 = atomicAdd(cnt + 1, (unsigned long long)__popc(m1));
            if (m2) b2 = atomicAdd(cnt, 4ull * (unsigned long long)__popc(m2));
        }
        b1 = __shfl_sync(0xFFFFFFFFu, b1, 0);
        b2 = __shfl_sync(0xFFFFFFFFu, b2, 0);
        if (v == 1) {
            const unsigned long long r = b1 + (unsigned long long)__popc(m1 & lt);
            rlo[r] = lo;
            rhi[r] = hi;
            rdisc[r] = d;
        } else if (v == 2) {
            const unsigned long long c0 = b2 + 4ull * (unsigned long long)__popc(m2 & lt);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                npix[c0 + j] = 4 * pix + j;
                ndisc[c0 + j] = d;
            }
        }
    }
}

__global__ void __launch_bounds__(HPGB_BLOCK) k_qnb_iota(int32_t *p, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * HPGB_BLOCK;
    for (int64_t i = (int64_t)blockIdx.x * HPGB_BLOCK + threadIdx.x; i < n; i += stride) p[i] = (int32_t)i;
}
__global__ void __launch_bounds__(HPGB_BLOCK) k_qnb_gather_disc(const int32_t *__restrict__ disc, const int32_t *__restrict__ perm,
                                                                int64_t n, int32_t *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * HPGB_BLOCK;
    for (int64_t i = (int64_t)blockIdx.x * HPGB_BLOCK + threadIdx.x; i < n; i += stride) out[i] = disc[perm[i]];
}
// rows in (disc, lo) order; a repeated (disc, lo) -- the breadth-first stand-in for "unwind the stack" -- is emptied
__global__ void __launch_bounds__(HPGB_BLOCK) k_qnb_table(const int64_t *__restrict__ lo, const int64_t *__restrict__ hi,
                                                          const int32_t *__restrict__ perm, const int32_t *__restrict__ sdisc,
                                                          int64_t m, longlong2 *__restrict__ table) {
    const int64_t stride = (int64_t)gridDim.x * HPGB_BLOCK;
    for (int64_t i = (int64_t)blockIdx.x * HPGB_BLOCK + threadIdx.x; i < m; i += stride) {
        const int32_t p = perm[i];
        const int64_t a = lo[p];
        const bool dup = i > 0 && sdisc[i - 1] == sdisc[i] && lo[perm[i - 1]] == a;
        table[i] = make_longlong2(a, dup ? a : hi[p]);
    }
}
// disc_offs[d] = first output pixel of disc d = offs[first row whose disc >= d]
__global__ void __launch_bounds__(HPGB_BLOCK) k_qnb_disc_offsets(const int32_t *__restrict__ sdisc, int64_t m, const int64_t *__restrict__ offs,
                                                                 int64_t D, int64_t *__restrict__ disc_offs) {
    const int64_t stride = (int64_t)gridDim.x * HPGB_BLOCK;
    for (int64_t d = (int64_t)blockIdx.x * HPGB_BLOCK + threadIdx.x; d <= D; d += stride) {
        int64_t lo = 0, hi = m;  // lower bound of d in sdisc[0..m)
        while (lo < hi) {
            const int64_t mid = lo + ((hi - lo) >> 1);
            if ((int64_t)sdisc[mid] < d) lo = mid + 1; else hi = mid;
        }
        disc_offs[d] = offs[lo];
    }
}

inline int64_t grid_for(int64_t n) {
    int64_t grid = (n + HPGB_BLOCK - 1) / HPGB_BLOCK;
    const int64_t cap = (int64_t)hpgb_sm_count() * 8;
    if (grid > cap) grid = cap;
    return grid < 1 ? 1 : grid;
}

int qnb_build(const std::vector<hpgb_qn_plan> &plans, int64_t nside, QnHandle *h) {
    const int64_t D = (int64_t)plans.size();
    int prev = 0;
    cudaGetDevice(&prev);
    CKQ(cudaSetDevice(h->device));
    DevBuf pool;
    const int64_t npix = 12 * nside * nside;
    const int omax = plans[0].omax;
    hpgb_qn_plan *d_plans = nullptr;
    CKQ(pool.alloc(&d_plans, (size_t)D * sizeof(hpgb_qn_plan)));
    CKQ(cudaMemcpyAsync(d_plans, plans.data(), (size_t)D * sizeof(hpgb_qn_plan), cudaMemcpyHostToDevice, 0));
    // whole-sphere discs skip the search: their single range goes straight into the list
    std::vector<int32_t> active;
    std::vector<int64_t> w_lo, w_hi;
    std::vector<int32_t> w_disc;
    for (int64_t d = 0; d < D; d++) {
        if (plans[d].whole_sphere) {
            w_lo.push_back(0), w_hi.push_back(npix), w_disc.push_back((int32_t)d);
        } else {
            active.push_back((int32_t)d);
        }
    }
    std::vector<int64_t *> lvl_lo, lvl_hi;
    std::vector<int32_t *> lvl_disc;
    std::vector<int64_t> lvl_n;
    int64_t R = 0;
    if (!w_lo.empty()) {
        int64_t *a = nullptr, *b = nullptr;
        int32_t *c = nullptr;
        const size_t k = w_lo.size();
        CKQ(pool.alloc(&a, k * 8));
        CKQ(pool.alloc(&b, k * 8));
        CKQ(pool.alloc(&c, k * 4));
        CKQ(cudaMemcpyAsync(a, w_lo.data(), k * 8, cudaMemcpyHostToDevice, 0));
        CKQ(cudaMemcpyAsync(b, w_hi.data(), k * 8, cudaMemcpyHostToDevice, 0));
        CKQ(cudaMemcpyAsync(c, w_disc.data(), k * 4, cudaMemcpyHostToDevice, 0));
        lvl_lo.push_back(a), lvl_hi.push_back(b), lvl_disc.push_back(c), lvl_n.push_back((int64_t)k);
        R += (int64_t)k;
    }
    int64_t n = 12 * (int64_t)active.size();
    if (n > 0) {
        unsigned long long *cnt = nullptr;
        int32_t *d_active = nullptr, *cdisc = nullptr;
        int64_t *cpix = nullptr;
        CKQ(pool.alloc(&cnt, 16));
        CKQ(pool.alloc(&d_active, active.size() * 4));
        CKQ(cudaMemcpyAsync(d_active, active.data(), active.size() * 4, cudaMemcpyHostToDevice, 0));
        CKQ(pool.alloc(&cpix, (size_t)n * 8));
        CKQ(pool.alloc(&cdisc, (size_t)n * 4));
        k_qnb_seed<<<(unsigned)grid_for(n), HPGB_BLOCK, 0, 0>>>(cpix, cdisc, d_active, (int64_t)active.size());
        int launches = 1;
        for (int o = 0; o <= omax && n > 0; o++) {
            int64_t *np_ = nullptr, *rlo = nullptr, *rhi = nullptr;
            int32_t *nd_ = nullptr, *rdisc = nullptr;
            CKQ(pool.alloc(&np_, (size_t)n * 32));
            CKQ(pool.alloc(&nd_, (size_t)n * 16));
            CKQ(pool.alloc(&rlo, (size_t)n * 8));
            CKQ(pool.alloc(&rhi, (size_t)n * 8));
            CKQ(pool.alloc(&rdisc, (size_t)n * 4));
            CKQ(cudaMemsetAsync(cnt, 0, 16, 0));
            const hpgb_geom g = hpgb_make_geom((int64_t)1 << o, 1);
            k_qnb_level<<<(unsigned)grid_for(n), HPGB_BLOCK, 0, 0>>>(d_plans, g, o, cpix, cdisc, n, np_, nd_, rlo, rhi, rdisc, cnt);
            launches++;
            unsigned long long hc[2];
            CKQ(cudaMemcpy(hc, cnt, 16, cudaMemcpyDeviceToHost));
            lvl_lo.push_back(rlo), lvl_hi.push_back(rhi), lvl_disc.push_back(rdisc), lvl_n.push_back((int64_t)hc[1]);
            R += (int64_t)hc[1];
            hpgb_dev_free(cpix), pool.release(cpix);
            hpgb_dev_free(cdisc), pool.release(cdisc);
            cpix = np_, cdisc = nd_;
            n = (int64_t)hc[0];
        }
        g_hpgb_launches.fetch_add(launches, std::memory_order_relaxed);
    }
    h->m = R;
    h->total = 0;
    h->n_discs = D;
    CKQ(hpgb_dev_alloc((void **)&h->disc_offs, (size_t)(D + 1) * 8));
    if (R == 0) {
        CKQ(cudaMemsetAsync(h->disc_offs, 0, (size_t)(D + 1) * 8, 0));
    } else {
        if (R > 0x7FFFFFFF) {
            cudaSetDevice(prev);
            return HPGB_E_BAD_ARGUMENT;  // 32-bit row permutation
        }
        int64_t *lo = nullptr, *hi = nullptr, *slo = nullptr;
        int32_t *disc = nullptr, *perm0 = nullptr, *perm1 = nullptr, *perm2 = nullptr, *kd = nullptr, *sd = nullptr;
        CKQ(pool.alloc(&lo, (size_t)R * 8));
        CKQ(pool.alloc(&hi, (size_t)R * 8));
        CKQ(pool.alloc(&slo, (size_t)R * 8));
        CKQ(pool.alloc(&disc, (size_t)R * 4));
        CKQ(pool.alloc(&perm0, (size_t)R * 4));
        CKQ(pool.alloc(&perm1, (size_t)R * 4));
        CKQ(pool.alloc(&perm2, (size_t)R * 4));
        CKQ(pool.alloc(&kd, (size_t)R * 4));
        CKQ(pool.alloc(&sd, (size_t)R * 4));
        int64_t at = 0;
        for (size_t l = 0; l < lvl_n.size(); l++) {
            if (lvl_n[l] > 0) {
                CKQ(cudaMemcpyAsync(lo + at, lvl_lo[l], (size_t)lvl_n[l] * 8, cudaMemcpyDeviceToDevice, 0));
                CKQ(cudaMemcpyAsync(hi + at, lvl_hi[l], (size_t)lvl_n[l] * 8, cudaMemcpyDeviceToDevice, 0));
                CKQ(cudaMemcpyAsync(disc + at, lvl_disc[l], (size_t)lvl_n[l] * 4, cudaMemcpyDeviceToDevice, 0));
            }
            at += lvl_n[l];
        }
        const unsigned gr = (unsigned)grid_for(R);
        k_qnb_iota<<<gr, HPGB_BLOCK, 0, 0>>>(perm0, R);
        void *tmp = nullptr;
        size_t t1 = 0, t2 = 0;
        CKQ(cub::DeviceRadixSort::SortPairs(nullptr, t1, lo, slo, perm0, perm1, R, 0, 62, (cudaStream_t)0));
        CKQ(cub::DeviceRadixSort::SortPairs(nullptr, t2, kd, sd, perm1, perm2, R, 0, 31, (cudaStream_t)0));
        const size_t tb = t1 > t2 ? t1 : t2;
        CKQ(pool.alloc(&tmp, tb));
        size_t tt = tb;
        CKQ(cub::DeviceRadixSort::SortPairs(tmp, tt, lo, slo, perm0, perm1, R, 0, 62, (cudaStream_t)0));  // by lo
        k_qnb_gather_disc<<<gr, HPGB_BLOCK, 0, 0>>>(disc, perm1, R, kd);
        tt = tb;
        CKQ(cub::DeviceRadixSort::SortPairs(tmp, tt, kd, sd, perm1, perm2, R, 0, 31, (cudaStream_t)0));    // then (stable) by disc
        CKQ(hpgb_dev_alloc((void **)&h->ranges, (size_t)R * 16));
        CKQ(hpgb_dev_alloc((void **)&h->offs, (size_t)(R + 1) * 8));
        k_qnb_table<<<gr, HPGB_BLOCK, 0, 0>>>(lo, hi, perm2, sd, R, reinterpret_cast<longlong2 *>(h->ranges));
        hpgb_status_t *st = nullptr;
        CKQ(pool.alloc(&st, sizeof(hpgb_status_t)));
        CKQ(cudaMemsetAsync(st, 0xFF, sizeof(hpgb_status_t), 0));
        const int rc = launch_ranges_offsets(h->ranges, R, 0, h->offs, st, 0);
        if (rc != HPGB_OK) {
            cudaSetDevice(prev);
            return rc;
        }
        k_qnb_disc_offsets<<<(unsigned)grid_for(D + 1), HPGB_BLOCK, 0, 0>>>(sd, R, h->offs, D, h->disc_offs);
        g_hpgb_launches.fetch_add(6, std::memory_order_relaxed);
        CKQ(cudaMemcpy(&h->total, h->offs + R, 8, cudaMemcpyDeviceToHost));
    }
    cudaError_t e = cudaDeviceSynchronize();
    cudaSetDevice(prev);
    return e == cudaSuccess ? HPGB_OK : HPGB_E_CUDA + (int)e;
}

}  // namespace

extern "C" {

int hpgb_query_circle_nest_batch_open(int64_t nside, const double *a, const double *b, const double *radius, int64_t D,
                                      int inclusive, int64_t fact, int lonlat, int degrees, int device, void **handle,
                                      int64_t *m, int64_t *total, int64_t *first_bad) {
    device = hpgb_host_first_device(device);  // HPGB_DEVICE_ALL: this entry point's scratch lives on one GPU
    if (!handle || !m || !total || D < 0 || (D > 0 && (!a || !b || !radius))) return HPGB_E_BAD_ARGUMENT;
    *handle = nullptr;
    if (first_bad) *first_bad = -1;
    if (D > 0x7FFFFFFF) return HPGB_E_BAD_ARGUMENT;
    std::vector<hpgb_qn_plan> plans((size_t)D);
    double dr_cache[HPGB_QN_MAX_ORDER + 1];
    for (int o = 0; o <= HPGB_QN_MAX_ORDER; o++) dr_cache[o] = -1.0;
    for (int64_t d = 0; d < D; d++) {  // argument checks per disc, in the wrapper's order (hpgeom/hpgeom.c:793-838)
        double theta, phi, r = radius[d];
        bool ok;
        if (lonlat) {
            ok = degrees ? hpgb_angles_exact<true, true>(a[d], b[d], theta, phi) : hpgb_angles_exact<true, false>(a[d], b[d], theta, phi);
            if (ok && degrees) r *= HPGB_PI / 180.0;
        } else {
            ok = hpgb_angles_exact<false, false>(a[d], b[d], theta, phi);
        }
        int rc = HPGB_OK;
        if (!ok) rc = HPGB_E_BAD_ANGLE;
        else if (r <= 0) rc = HPGB_E_BAD_RADIUS;
        else if ((rc = hpgb_check_nside(nside, 1)) != 0) {
        } else if (inclusive && (fact <= 0 || fact * nside > ((int64_t)1 << 29) || (fact & (fact - 1)))) rc = HPGB_E_BAD_FACT;
        else rc = hpgb_qn_make_plan(nside, theta, phi, r, inclusive ? fact : 0, &plans[(size_t)d], dr_cache);
        if (rc != HPGB_OK) {
            if (first_bad) *first_bad = d;
            return rc;
        }
    }
    const int ndev = hpgb_device_count();
    if (ndev <= 0) return HPGB_E_NO_DEVICE;
    if (device < 0 || device >= ndev) return HPGB_E_BAD_ARGUMENT;
    QnHandle *h = new QnHandle();
    h->device = device;
    int rc = HPGB_OK;
    if (D > 0) rc = qnb_build(plans, nside, h);
    else {
        int prev = 0;
        cudaGetDevice(&prev);
        cudaSetDevice(device);
        if (hpgb_dev_alloc((void **)&h->disc_offs, 8) != cudaSuccess || cudaMemset(h->disc_offs, 0, 8) != cudaSuccess) rc = HPGB_E_CUDA;
        cudaSetDevice(prev);
    }
    if (rc != HPGB_OK) {
        hpgb_query_nest_close(h);
        return rc;
    }
    *handle = h;
    *m = h->m;
    *total = h->total;
    return HPGB_OK;
}

// the n_discs + 1 boundaries of the pixel lists of a batched handle (host buffer)
int hpgb_query_nest_batch_offsets(void *handle, int64_t *disc_offsets) {
    QnHandle *h = (QnHandle *)handle;
    if (!h || !h->disc_offs || !disc_offsets) return HPGB_E_BAD_ARGUMENT;
    int prev = 0;
    cudaGetDevice(&prev);
    CKQ(cudaSetDevice(h->device));
    CKQ(cudaMemcpy(disc_offsets, h->disc_offs, (size_t)(h->n_discs + 1) * 8, cudaMemcpyDeviceToHost));
    cudaSetDevice(prev);
    return HPGB_OK;
}

}  // extern "C"
