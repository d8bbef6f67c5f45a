// k_query.cu -- query_circle in the RING scheme (reference: hpgeom/hpgeom.c:765-864 and query_disc,
// hpgeom/healpix_geom.c:632-747): per-disc plan on the host, one GPU thread per ring for the pixel
// intervals (hpgb_query.h), expansion by the kernels of k_ranges.cu.  See include/hpgb.h.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <atomic>

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
        if (work) cudaFree(work);
        if (ranges) cudaFree(ranges);
        if (offs) cudaFree(offs);
        if (st) cudaFree(st);
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
    CKQ(cudaMalloc(&d.ranges, (size_t)plan->m * 16));
    CKQ(cudaMalloc(&d.offs, (size_t)(plan->m + 1) * 8));
    CKQ(cudaMalloc(&d.st, sizeof(hpgb_status_t)));
    CKQ(cudaMemsetAsync(d.st, 0xFF, sizeof(hpgb_status_t), 0));
    if (qc_work_bytes(plan) > 0) CKQ(cudaMalloc(&d.work, (size_t)qc_work_bytes(plan)));
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
    if (!total) return HPGB_E_BAD_ARGUMENT;
    DevTable d;
    return build_table(plan, device, d, total);
}

int hpgb_host_query_circle_ring_fill(const hpgb_disc_t *plan, int64_t *out, int64_t total, int device) {
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
