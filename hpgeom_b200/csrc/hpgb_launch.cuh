// hpgb_launch.cuh -- the streaming kernel skeletons every conversion shares, and launch helpers.
//
// Data layout in HBM: structure-of-arrays, one C-contiguous int64/float64 array per operand
// (exactly the operands the reference's NpyIter walks, hpgeom/hpgeom.c:203-218), touched once.
// Kernel shape:
//   * k_map2: each thread owns PAIRS of consecutive elements and moves them with 128-bit
//     ld.global.cs / st.global.cs (streaming: evict-first, the data is never re-read);
//     grid-stride over pairs with two independent pairs in flight per thread so that every
//     SM keeps >= 32 B/thread of loads outstanding; 64-bit element indices (n may exceed 2^31);
//     grid = #SMs x resident CTAs/SM (occupancy API), never more CTAs than there is work;
//   * k_map1: scalar fallback for pointers that are not 16-byte aligned (torch slices) and for
//     per-element nside arrays;
//   * element errors: each thread keeps the smallest failing index in a register and issues at
//     most ONE atomicMin at exit, only if it saw a failure (no cost on the valid path).
#ifndef HPGB_LAUNCH_CUH
#define HPGB_LAUNCH_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/hpgb.h"
#include "hpgb_tables.h"

#define HPGB_BLOCK 256

extern std::atomic<int64_t> g_hpgb_launches;

struct hpgb_bad_t {
    unsigned long long idx;
    __device__ __forceinline__ hpgb_bad_t() : idx(HPGB_NO_BAD_ELEMENT) {}
    __device__ __forceinline__ void flag(int64_t i) {
        if ((unsigned long long)i < idx) idx = (unsigned long long)i;
    }
    __device__ __forceinline__ void commit(hpgb_status_t *st) const {
        if (idx != HPGB_NO_BAD_ELEMENT) atomicMin(&st->first_bad, idx);
    }
};

// Op interface:
//   typename Op::In2;  In2 load2(int64_t i) const;                 // loads elements i, i+1 (i even)
//   typename Op::Ctx;  Ctx prologue() const;                       // per-CTA setup (e.g. stage the trig table)
//   void run2(int64_t i, const In2&, const Ctx&, hpgb_bad_t&) const; // converts + stores i, i+1
//   void run4(i0, v0, i1, v1, ctx, bad) const;                      // two pairs (ILP across both)
//   void run1(int64_t i, const Ctx&, hpgb_bad_t&) const;            // scalar element
template <class Op>
__global__ void __launch_bounds__(HPGB_BLOCK) k_map2(const Op op, const int64_t n, hpgb_status_t *st) {
    const typename Op::Ctx ctx = op.prologue();
    const int64_t npair = n >> 1;
    const int64_t stride = (int64_t)gridDim.x * HPGB_BLOCK;
    int64_t p = (int64_t)blockIdx.x * HPGB_BLOCK + threadIdx.x;
    hpgb_bad_t bad;
    for (; p + stride < npair; p += 2 * stride) {
        const typename Op::In2 v0 = op.load2(2 * p);
        const typename Op::In2 v1 = op.load2(2 * (p + stride));
        op.run4(2 * p, v0, 2 * (p + stride), v1, ctx, bad);
    }
    if (p < npair) op.run2(2 * p, op.load2(2 * p), ctx, bad);
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) op.run1(n - 1, ctx, bad);
    bad.commit(st);
}

template <class Op>
__global__ void __launch_bounds__(HPGB_BLOCK) k_map1(const Op op, const int64_t n, hpgb_status_t *st) {
    const typename Op::Ctx ctx = op.prologue();
    const int64_t stride = (int64_t)gridDim.x * HPGB_BLOCK;
    hpgb_bad_t bad;
    for (int64_t i = (int64_t)blockIdx.x * HPGB_BLOCK + threadIdx.x; i < n; i += stride) op.run1(i, ctx, bad);
    bad.commit(st);
}

// the 52 x 4 double-double sin/cos table of hpgb_math.h, staged in shared memory per CTA
// (1664 B; random row gathers from it are LSU work that overlaps the fp64 pipe)
static __device__ const double g_hpgb_sincos[4 * HPGB_SINCOS_ROWS] = HPGB_SINCOS_TABLE_INIT;
struct hpgb_no_ctx {};
// default run4 = two run2 calls
#define HPGB_DEFAULT_RUN4                                                                                      \
    __device__ __forceinline__ void run4(int64_t i0, const In2 &v0, int64_t i1, const In2 &v1, const Ctx &c, \
                                         hpgb_bad_t &bad) const {                                             \
        run2(i0, v0, c, bad);                                                                                 \
        run2(i1, v1, c, bad);                                                                                 \
    }
__device__ __forceinline__ const double *hpgb_stage_table() {
    __shared__ double sT[4 * HPGB_SINCOS_ROWS];
    for (int i = threadIdx.x; i < 4 * HPGB_SINCOS_ROWS; i += HPGB_BLOCK) sT[i] = g_hpgb_sincos[i];
    __syncthreads();
    return sT;
}

int hpgb_sm_count();  // SMs of the current device (148 on B200)

template <class K>
static int hpgb_resident_ctas(K kernel) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, HPGB_BLOCK, 0);
    return per_sm > 0 ? per_sm : 1;
}

static inline bool hpgb_aligned16(const void *p) { return (((uintptr_t)p) & 15u) == 0; }

// launch the scalar kernel (misaligned operands, per-element nside)
template <class Op>
static int hpgb_launch_map1(const Op &op, int64_t n, hpgb_status_t *st, cudaStream_t s) {
    if (n <= 0) return HPGB_OK;
    static int occ1 = 0;
    if (!occ1) occ1 = hpgb_resident_ctas(k_map1<Op>);
    int64_t want = (n + HPGB_BLOCK - 1) / HPGB_BLOCK;
    int64_t grid = (int64_t)hpgb_sm_count() * occ1;
    if (want < grid) grid = want;
    k_map1<Op><<<(unsigned)grid, HPGB_BLOCK, 0, s>>>(op, n, st);
    g_hpgb_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? HPGB_OK : HPGB_E_CUDA + (int)e;
}

// launch `op` over n elements; vec_ok = all operand pointers are 16-byte aligned
template <class Op>
static int hpgb_launch_map(const Op &op, int64_t n, bool vec_ok, hpgb_status_t *st, cudaStream_t s) {
    if (n <= 0) return HPGB_OK;
    if (!vec_ok) return hpgb_launch_map1(op, n, st, s);
    static int occ2 = 0;  // per instantiation, per process (one device kind per box)
    if (!occ2) occ2 = hpgb_resident_ctas(k_map2<Op>);
    int64_t want = ((n >> 1) + HPGB_BLOCK - 1) / HPGB_BLOCK;
    int64_t grid = (int64_t)hpgb_sm_count() * occ2;
    if (want < grid) grid = want > 0 ? want : 1;
    k_map2<Op><<<(unsigned)grid, HPGB_BLOCK, 0, s>>>(op, n, st);
    g_hpgb_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? HPGB_OK : HPGB_E_CUDA + (int)e;
}

// streaming 128-bit accessors
__device__ __forceinline__ longlong2 hpgb_ld2(const int64_t *p) { return __ldcs(reinterpret_cast<const longlong2 *>(p)); }
__device__ __forceinline__ double2 hpgb_ld2(const double *p) { return __ldcs(reinterpret_cast<const double2 *>(p)); }
__device__ __forceinline__ void hpgb_st2(int64_t *p, int64_t a, int64_t b) { __stcs(reinterpret_cast<longlong2 *>(p), make_longlong2(a, b)); }
__device__ __forceinline__ void hpgb_st2(double *p, double a, double b) { __stcs(reinterpret_cast<double2 *>(p), make_double2(a, b)); }
__device__ __forceinline__ int64_t hpgb_ld1(const int64_t *p) { return __ldcs(p); }
__device__ __forceinline__ double hpgb_ld1(const double *p) { return __ldcs(p); }
__device__ __forceinline__ void hpgb_st1(int64_t *p, int64_t a) { __stcs(p, a); }
__device__ __forceinline__ void hpgb_st1(double *p, double a) { __stcs(p, a); }

#endif  // HPGB_LAUNCH_CUH
