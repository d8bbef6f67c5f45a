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
#include <type_traits>

#include "../../include/hpgb.h"
#include "hpgb_tables.h"

#define HPGB_BLOCK 256


extern std::atomic<int64_t> g_hpgb_launches;

// ops that defer work to the end of the streaming kernel (e.g. a queue of rare elements) declare
// `static constexpr bool HAS_EPILOGUE = true` and provide run4u(...) -- run4 for a converged CTA, may use
// warp collectives -- and epilogue(ctx, bad); k_map2 / k_map1 keep calling run4 / run2 / run1
template <class T, class = void>
struct hpgb_has_epilogue : std::false_type {};
template <class T>
struct hpgb_has_epilogue<T, std::void_t<decltype(T::HAS_EPILOGUE)>> : std::bool_constant<T::HAS_EPILOGUE> {};

// ops whose per-CTA tables live in the streaming kernel's dynamic shared memory (behind the stage ring) declare
// `static constexpr bool SMEM_PROLOGUE = true` and provide prologue_extra(unsigned char *extra) next to prologue()
template <class T, class = void>
struct hpgb_has_smem_prologue : std::false_type {};
template <class T>
struct hpgb_has_smem_prologue<T, std::void_t<decltype(T::SMEM_PROLOGUE)>> : std::bool_constant<T::SMEM_PROLOGUE> {};

struct hpgb_bad_t {
    unsigned long long idx;
    int aux;  // per-thread scratch an op may carry through the kernel (e.g. the length of its warp's queue)
    __device__ __forceinline__ hpgb_bad_t() : idx(HPGB_NO_BAD_ELEMENT), aux(0) {}
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
template <class Op, int BLOCK = HPGB_BLOCK>
__global__ void __launch_bounds__(BLOCK) k_map2(const Op op, const int64_t n, hpgb_status_t *st) {
    const typename Op::Ctx ctx = op.prologue();
    const int64_t npair = n >> 1;
    const int64_t stride = (int64_t)gridDim.x * BLOCK;
    int64_t p = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
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

template <class Op, int BLOCK = HPGB_BLOCK>
__global__ void __launch_bounds__(BLOCK) k_map1(const Op op, const int64_t n, hpgb_status_t *st) {
    const typename Op::Ctx ctx = op.prologue();
    const int64_t stride = (int64_t)gridDim.x * BLOCK;
    hpgb_bad_t bad;
    for (int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x; i < n; i += stride) op.run1(i, ctx, bad);
    bad.commit(st);
}

// the 52 x 4 double-double sin/cos table of hpgb_math.h followed by the 17-entry atan(k/16) table of
// the vector_to_pixel fast path, staged in shared memory per CTA (1800 B; random row gathers from it
// are LSU work that overlaps the fp64 pipe)
#define HPGB_TABLE_DOUBLES (4 * HPGB_SINCOS_ROWS + HPGB_ATAN16_ROWS)
static __device__ const double g_hpgb_sincos[4 * HPGB_SINCOS_ROWS] = HPGB_SINCOS_TABLE_INIT;
static __device__ const double g_hpgb_atan16[HPGB_ATAN16_ROWS] = HPGB_ATAN16_TABLE_INIT;
struct hpgb_no_ctx {};
// default run4 = two run2 calls
#define HPGB_DEFAULT_RUN4                                                                                      \
    __device__ __forceinline__ void run4(int64_t i0, const In2 &v0, int64_t i1, const In2 &v1, const Ctx &c, \
                                         hpgb_bad_t &bad) const {                                             \
        run2(i0, v0, c, bad);                                                                                 \
        run2(i1, v1, c, bad);                                                                                 \
    }
__device__ __forceinline__ const double *hpgb_stage_table() {
    __shared__ double sT[HPGB_TABLE_DOUBLES];
    for (int i = threadIdx.x; i < 4 * HPGB_SINCOS_ROWS; i += blockDim.x) sT[i] = g_hpgb_sincos[i];
    for (int i = threadIdx.x; i < HPGB_ATAN16_ROWS; i += blockDim.x) sT[4 * HPGB_SINCOS_ROWS + i] = g_hpgb_atan16[i];
    __syncthreads();
    return sT;
}

// -------------------------------------------------------------------------------------------
// k_stream: TMA-fed persistent streaming kernel (the fast path for 16-byte aligned operands).
//
// One elected thread per CTA keeps a ring of STAGES shared-memory tiles full with 1-D bulk
// async copies (cp.async.bulk.shared.global + mbarrier complete_tx -> SASS UBLKCP); a tile holds
// TILE = 1024*H consecutive elements of each input array.  Producer/consumer hand-over is two
// mbarriers per stage and NO CTA-wide barrier:
//   full[s]  (count 1, + transaction bytes) : the TMA engine completes it when the tile has landed;
//   empty[s] (count 8 = warps per CTA)      : each warp arrives once it has pulled its elements out
//                                             of the stage with conflict-free 128-bit LDS.
// The eight warps therefore drift freely (a warp that hits the rare exact path does not stall the
// other seven); only the producer thread ever waits on `empty`, and it refills the stage that was
// consumed one iteration EARLIER, whose readers have normally long gone.  Results go straight to
// HBM with 128-bit streaming stores.  Global-load latency is hidden by the ring, not by occupancy:
// no registers are tied up by loads in flight and the kernel stays busy even at the 2-3 CTAs/SM the
// fp64 register footprint allows.
// Within sub-block h of a tile, thread t owns elements {2t, 2t+1} and {512 + 2t, 512 + 2t + 1}, so
// each warp LDS / STG instruction covers one contiguous 512-byte span.
// Op additionally provides:  static constexpr int NIN, MINB (min resident CTAs/SM -> register cap);
//                            const void *src(int k) const;
//                            In2 load2s(const unsigned char *stage, int tile_elems, int elem) const;
//                            static constexpr bool TILE_HOOK; static constexpr int EXTRA_SMEM;
//   TILE_HOOK ops replace the per-thread run4 calls by  tile<TILE>(stage, extra_smem, first_element,
//   tid, ctx, bad, empty_barrier)  and may rewrite the stage in place (see OpReorder, ring -> nest).
// -------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t hpgb_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void hpgb_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(hpgb_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void hpgb_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(hpgb_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void hpgb_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(hpgb_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void hpgb_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
#if defined(HPGB_TMA_EVICT_FIRST)
    // the tile is read exactly once: mark its L2 lines evict-first so that they, not the dirty
    // output lines waiting for write-back, are the victims
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     hpgb_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(hpgb_smem_u32(bar)), "l"(pol)
                 : "memory");
#else
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     hpgb_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(hpgb_smem_u32(bar))
                 : "memory");
#endif
}
__device__ __forceinline__ void hpgb_mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = hpgb_smem_u32(bar);
    uint32_t ok;
    do {
#if defined(HPGB_TRYWAIT_HINT_NS)  // A/B: ask the hardware to suspend the warp for up to this long per attempt
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity), "r"((uint32_t)HPGB_TRYWAIT_HINT_NS)
            : "memory");
#else
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
#endif
    } while (!ok);
}

// BLOCK = threads per CTA (a sub-block is 4 * BLOCK elements: thread t owns {2t, 2t+1} and {2 BLOCK + 2t, + 1});
// MINB = min resident CTAs/SM (register cap); CONTIG = false: tiles are dealt round-robin to the CTAs (the grid
// sweeps the arrays as one moving window), true: every CTA owns one contiguous run of tiles.
template <class Op, int STAGES, int H, int BLOCK = HPGB_BLOCK, int MINB = Op::MINB, bool CONTIG = false>
__global__ void __launch_bounds__(BLOCK, MINB) k_stream(const Op op, const int64_t n, hpgb_status_t *st) {
    constexpr int NIN = Op::NIN;
    constexpr int SUB = 4 * BLOCK;
    constexpr int TILE = SUB * H;
    constexpr uint32_t ARR_BYTES = TILE * 8;
    static_assert(!Op::TILE_HOOK || BLOCK == HPGB_BLOCK, "tile-hook ops are written for HPGB_BLOCK threads");
    extern __shared__ __align__(128) unsigned char hpgb_stage_smem[];
    __shared__ uint64_t full[STAGES], empty[STAGES];
    unsigned char *const extra_smem = hpgb_stage_smem + (size_t)STAGES * NIN * ARR_BYTES;  // Op::EXTRA_SMEM bytes
    const typename Op::Ctx ctx = [&]() {
        if constexpr (hpgb_has_smem_prologue<Op>::value)
            return op.prologue_extra(extra_smem);
        else
            return op.prologue();
    }();
    const int tid = threadIdx.x;
    const int64_t ntiles = n / TILE;
    // this CTA's tiles: t0, t0 + tstep, ... < tend
    int64_t t0, tend, tstep;
    if (CONTIG) {
        const int64_t per = ntiles / gridDim.x, rem = ntiles % gridDim.x;
        t0 = (int64_t)blockIdx.x * per + ((int64_t)blockIdx.x < rem ? (int64_t)blockIdx.x : rem);
        tend = t0 + per + ((int64_t)blockIdx.x < rem ? 1 : 0);
        tstep = 1;
    } else {
        t0 = blockIdx.x;
        tend = ntiles;
        tstep = gridDim.x;
    }
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            hpgb_mbar_init(&full[s], 1);
            hpgb_mbar_init(&empty[s], BLOCK / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int s, int64_t tile) {
        hpgb_mbar_expect_tx(&full[s], NIN * ARR_BYTES);
#pragma unroll
        for (int k = 0; k < NIN; k++)
            hpgb_bulk_g2s(hpgb_stage_smem + (size_t)(s * NIN + k) * ARR_BYTES,
                          (const unsigned char *)op.src(k) + (size_t)tile * ARR_BYTES, ARR_BYTES, &full[s]);
    };
    if (tid == 0) {
        int64_t t = t0;
        for (int s = 0; s < STAGES && t < tend; s++, t += tstep) issue(s, t);
    }
    hpgb_bad_t bad;
    int s = 0, sprev = STAGES - 1;
    uint32_t phase = 0, phase_prev = 1;  // parity of full[s] now / of empty[sprev] for the refill
    bool first = true;
    for (int64_t t = t0; t < tend; t += tstep) {
        if (tid == 0 && !first) {  // refill the stage every warp left one iteration ago
            const int64_t tn = t + (int64_t)(STAGES - 1) * tstep;
#if defined(HPGB_STREAM_NOMEM)
            // compute ceiling of an op (never built into the library): the ring is filled ONCE, every later iteration
            // converts the same resident tiles again and stores into this CTA's own window of `out` (L2 resident)
            if (false) {
#else
            if (tn < tend) {
#endif
                hpgb_mbar_wait(&empty[sprev], phase_prev);
                issue(sprev, tn);
            }
        }
        first = false;
#if defined(HPGB_STREAM_NOMEM)  // tools/ubench/a2p_sweep.cu only: compute ceiling, see below
        if (t < t0 + (int64_t)STAGES * tstep)
#endif
        hpgb_mbar_wait(&full[s], phase);
        const unsigned char *stage = hpgb_stage_smem + (size_t)s * NIN * ARR_BYTES;
        if constexpr (Op::TILE_HOOK) {
            // the op processes the whole tile itself (warp-cooperative) and arrives on empty[s]
            op.template tile<TILE>(const_cast<unsigned char *>(stage), extra_smem, t * TILE, tid, ctx, bad, &empty[s]);
        } else {
#pragma unroll
            for (int h = 0; h < H; h++) {
                const typename Op::In2 v0 = op.load2s(stage, TILE, h * SUB + 2 * tid);
                const typename Op::In2 v1 = op.load2s(stage, TILE, h * SUB + SUB / 2 + 2 * tid);
                if (h == H - 1) {  // this warp is done with the stage
                    __syncwarp();
                    if ((tid & 31) == 0) hpgb_mbar_arrive(&empty[s]);
                }
#if defined(HPGB_STREAM_NOMEM)
                const int64_t e0 = ((int64_t)blockIdx.x * STAGES + s) * TILE + h * SUB + 2 * tid;
#else
                const int64_t e0 = t * TILE + h * SUB + 2 * tid;
#endif
                if constexpr (hpgb_has_epilogue<Op>::value)
                    op.run4u(e0, v0, e0 + SUB / 2, v1, ctx, bad);  // every thread of the CTA is here: warp collectives allowed
                else
                    op.run4(e0, v0, e0 + SUB / 2, v1, ctx, bad);
            }
        }
        sprev = s;
        phase_prev = phase;
        if (++s == STAGES) {
            s = 0;
            phase ^= 1u;
        }
    }
    if constexpr (hpgb_has_epilogue<Op>::value) op.epilogue(ctx, bad);  // e.g. drain a queue of deferred elements
    if (blockIdx.x == 0)  // fewer than TILE leftover elements
        for (int64_t i = ntiles * TILE + tid; i < n; i += BLOCK) op.run1(i, ctx, bad);
    bad.commit(st);
}

int hpgb_sm_count();  // SMs of the current device (148 on B200)

// Launch configuration cached PER DEVICE: the opt-in for more than 48 KB of dynamic shared memory
// (cudaFuncAttributeMaxDynamicSharedMemorySize) is a per-device attribute of the function, so a process
// that uses cuda:0 and then cuda:1 must set it on each.  Slot = current device; 0 = not set up yet.
// (Relaxed atomics: two threads racing on the first call compute and store the same value.)
struct hpgb_occ_cache {
    std::atomic<int> v[64];
    hpgb_occ_cache() {
        for (auto &x : v) x.store(0, std::memory_order_relaxed);
    }
};
static inline int hpgb_current_device() {
    int d = 0;
    cudaGetDevice(&d);
    return (d < 0 || d >= 64) ? 0 : d;
}
// resident CTAs/SM of `kernel` on the current device (sets the dynamic-smem opt-in first); >= 1
template <class K>
static int hpgb_resident_ctas(K kernel, int dyn_smem = 0, int block = HPGB_BLOCK) {
    int per_sm = 0;
    if (dyn_smem > 0) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, dyn_smem);
    return per_sm > 0 ? per_sm : 1;
}
template <class K>
static int hpgb_resident_ctas_cached(hpgb_occ_cache &cache, K kernel, int dyn_smem = 0, int block = HPGB_BLOCK) {
    const int dev = hpgb_current_device();
    int occ = cache.v[dev].load(std::memory_order_relaxed);
    if (!occ) {
        occ = hpgb_resident_ctas(kernel, dyn_smem, block);
        cache.v[dev].store(occ, std::memory_order_relaxed);
    }
    return occ;
}

static inline bool hpgb_aligned16(const void *p) { return (((uintptr_t)p) & 15u) == 0; }

// launch the scalar kernel (misaligned operands, per-element nside); dyn_smem = bytes of dynamic
// shared memory the op's prologue() stages its tables into; BLOCK = threads per CTA (ops with big
// tables run ONE large CTA per SM so the table is staged once per SM)
template <class Op, int BLOCK = HPGB_BLOCK>
static int hpgb_launch_map1(const Op &op, int64_t n, hpgb_status_t *st, cudaStream_t s, int dyn_smem = 0) {
    if (n <= 0) return HPGB_OK;
    static hpgb_occ_cache cache;  // per instantiation
    const int occ1 = hpgb_resident_ctas_cached(cache, k_map1<Op, BLOCK>, dyn_smem, BLOCK);
    int64_t want = (n + BLOCK - 1) / BLOCK;
    int64_t grid = (int64_t)hpgb_sm_count() * occ1;
    if (want < grid) grid = want;
    k_map1<Op, BLOCK><<<(unsigned)grid, BLOCK, dyn_smem, s>>>(op, n, st);
    g_hpgb_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? HPGB_OK : HPGB_E_CUDA + (int)e;
}

// launch `op` over n elements; vec_ok = all operand pointers are 16-byte aligned
template <class Op, int BLOCK = HPGB_BLOCK>
static int hpgb_launch_map(const Op &op, int64_t n, bool vec_ok, hpgb_status_t *st, cudaStream_t s, int dyn_smem = 0) {
    if (n <= 0) return HPGB_OK;
    if (!vec_ok) return hpgb_launch_map1<Op, BLOCK>(op, n, st, s, dyn_smem);
    static hpgb_occ_cache cache;
    const int occ2 = hpgb_resident_ctas_cached(cache, k_map2<Op, BLOCK>, dyn_smem, BLOCK);
    int64_t want = ((n >> 1) + BLOCK - 1) / BLOCK;
    int64_t grid = (int64_t)hpgb_sm_count() * occ2;
    if (want < grid) grid = want > 0 ? want : 1;
    k_map2<Op, BLOCK><<<(unsigned)grid, BLOCK, dyn_smem, s>>>(op, n, st);
    g_hpgb_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? HPGB_OK : HPGB_E_CUDA + (int)e;
}

// launch the TMA streaming kernel (operands 16-byte aligned); falls back to k_map2 for small n
template <class Op, int STAGES, int H = 1, int BLOCK = HPGB_BLOCK, int MINB = Op::MINB, bool CONTIG = false>
static int hpgb_launch_stream(const Op &op, int64_t n, hpgb_status_t *st, cudaStream_t s) {
    constexpr int TILE = 4 * BLOCK * H;
    if (n < 4 * TILE) return hpgb_launch_map(op, n, true, st, s);
    const int smem = STAGES * Op::NIN * TILE * 8 + Op::EXTRA_SMEM;
    static hpgb_occ_cache cache;
    const int occ = hpgb_resident_ctas_cached(cache, k_stream<Op, STAGES, H, BLOCK, MINB, CONTIG>, smem, BLOCK);
    int64_t grid = (int64_t)hpgb_sm_count() * occ;
    const int64_t ntiles = n / TILE;
    if (ntiles < grid) grid = ntiles;
    k_stream<Op, STAGES, H, BLOCK, MINB, CONTIG><<<(unsigned)grid, BLOCK, smem, s>>>(op, n, st);
    g_hpgb_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? HPGB_OK : HPGB_E_CUDA + (int)e;
}

// launch `op` on the fastest skeleton its operands allow: the TMA streaming kernel when every pointer
// is 16-byte aligned, else the scalar kernel
template <class Op, int STAGES = 3, int H = 2>
static int hpgb_launch_auto(const Op &op, int64_t n, bool vec_ok, hpgb_status_t *st, cudaStream_t s) {
    if (vec_ok) return hpgb_launch_stream<Op, STAGES, H>(op, n, st, s);
    return hpgb_launch_map(op, n, false, st, s);
}

// members a streaming op needs when its only input array is int64 pixels / two or three float64 arrays
#define HPGB_STREAM_IN_I64(ptr, minb)                                                                       \
    static constexpr int NIN = 1;                                                                           \
    static constexpr int MINB = minb;                                                                       \
    static constexpr bool TILE_HOOK = false;                                                                \
    static constexpr int EXTRA_SMEM = 0;                                                                    \
    __device__ __forceinline__ const void *src(int) const { return ptr; }                                  \
    __device__ __forceinline__ In2 load2s(const unsigned char *stage, int, int e) const {                  \
        return *reinterpret_cast<const longlong2 *>(stage + 8 * e);                                         \
    }
#define HPGB_STREAM_IN_F64x2(p0, p1, minb)                                                                  \
    static constexpr int NIN = 2;                                                                           \
    static constexpr int MINB = minb;                                                                       \
    static constexpr bool TILE_HOOK = false;                                                                \
    static constexpr int EXTRA_SMEM = 0;                                                                    \
    __device__ __forceinline__ const void *src(int k) const { return k == 0 ? (const void *)p0 : (const void *)p1; } \
    __device__ __forceinline__ In2 load2s(const unsigned char *stage, int tile, int e) const {              \
        return In2{*reinterpret_cast<const double2 *>(stage + 8 * e),                                       \
                   *reinterpret_cast<const double2 *>(stage + 8 * tile + 8 * e)};                           \
    }
#define HPGB_STREAM_IN_F64x3(p0, p1, p2, minb)                                                              \
    static constexpr int NIN = 3;                                                                           \
    static constexpr int MINB = minb;                                                                       \
    static constexpr bool TILE_HOOK = false;                                                                \
    static constexpr int EXTRA_SMEM = 0;                                                                    \
    __device__ __forceinline__ const void *src(int k) const {                                              \
        return k == 0 ? (const void *)p0 : (k == 1 ? (const void *)p1 : (const void *)p2);                  \
    }                                                                                                       \
    __device__ __forceinline__ In2 load2s(const unsigned char *stage, int tile, int e) const {              \
        return In2{*reinterpret_cast<const double2 *>(stage + 8 * e),                                       \
                   *reinterpret_cast<const double2 *>(stage + 8 * tile + 8 * e),                            \
                   *reinterpret_cast<const double2 *>(stage + 16 * tile + 8 * e)};                          \
    }

// streaming 128-bit accessors
__device__ __forceinline__ longlong2 hpgb_ld2(const int64_t *p) { return __ldcs(reinterpret_cast<const longlong2 *>(p)); }
__device__ __forceinline__ double2 hpgb_ld2(const double *p) { return __ldcs(reinterpret_cast<const double2 *>(p)); }
#if defined(HPGB_ST_MODE) && HPGB_ST_MODE == 1   // plain write-back stores
__device__ __forceinline__ void hpgb_st2(int64_t *p, int64_t a, int64_t b) { *reinterpret_cast<longlong2 *>(p) = make_longlong2(a, b); }
__device__ __forceinline__ void hpgb_st2(double *p, double a, double b) { *reinterpret_cast<double2 *>(p) = make_double2(a, b); }
#elif defined(HPGB_ST_MODE) && HPGB_ST_MODE == 2  // cache-global
__device__ __forceinline__ void hpgb_st2(int64_t *p, int64_t a, int64_t b) { __stcg(reinterpret_cast<longlong2 *>(p), make_longlong2(a, b)); }
__device__ __forceinline__ void hpgb_st2(double *p, double a, double b) { __stcg(reinterpret_cast<double2 *>(p), make_double2(a, b)); }
#elif defined(HPGB_ST_MODE) && HPGB_ST_MODE == 3  // write-through
__device__ __forceinline__ void hpgb_st2(int64_t *p, int64_t a, int64_t b) { __stwt(reinterpret_cast<longlong2 *>(p), make_longlong2(a, b)); }
__device__ __forceinline__ void hpgb_st2(double *p, double a, double b) { __stwt(reinterpret_cast<double2 *>(p), make_double2(a, b)); }
#else
__device__ __forceinline__ void hpgb_st2(int64_t *p, int64_t a, int64_t b) { __stcs(reinterpret_cast<longlong2 *>(p), make_longlong2(a, b)); }
__device__ __forceinline__ void hpgb_st2(double *p, double a, double b) { __stcs(reinterpret_cast<double2 *>(p), make_double2(a, b)); }
#endif
// 256-bit streaming stores (sm_100: st.global.v4.b64 -> STG.E.EF.256): a lane writes one whole 32-byte sector -- a (4,)
// row of int64 / float64 -- in ONE request instead of two half-sector ones.  p must be 32-byte aligned.
__device__ __forceinline__ void hpgb_st4(int64_t *p, int64_t a, int64_t b, int64_t c, int64_t d) {
    asm volatile("st.global.cs.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}
__device__ __forceinline__ void hpgb_st4(double *p, double a, double b, double c, double d) {
    asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
static inline bool hpgb_aligned32(const void *p) { return (((uintptr_t)p) & 31u) == 0; }
__device__ __forceinline__ int64_t hpgb_ld1(const int64_t *p) { return __ldcs(p); }
__device__ __forceinline__ double hpgb_ld1(const double *p) { return __ldcs(p); }
__device__ __forceinline__ void hpgb_st1(int64_t *p, int64_t a) { __stcs(p, a); }
__device__ __forceinline__ void hpgb_st1(double *p, double a) { __stcs(p, a); }

// Row outputs ((n,8) neighbours, (n,4)+(n,4) interpolation): a lane holds NU consecutive 16-byte
// units, the warp's 32*NU units are contiguous in HBM.  Storing them lane by lane would make every
// warp store touch 32 half-filled sectors; instead the units take a turn through shared memory
// (XOR-swizzled, conflict-free both ways) so that each store instruction writes 512 contiguous
// bytes.  Full warps only (callers fall back to direct stores otherwise).  NU = 4 or 8.
// store policy of the coalesced row stores (A/B switch): 0 = streaming (evict-first), 1 = plain write-back
#ifndef HPGB_ROW_ST
#define HPGB_ROW_ST 0
#endif
__device__ __forceinline__ void hpgb_row_st(longlong2 *p, const longlong2 &v) {
#if HPGB_ROW_ST == 1
    *p = v;
#else
    __stcs(p, v);
#endif
}
template <int NU>
__device__ __forceinline__ void hpgb_warp_store_units(longlong2 *warp_dst, const longlong2 (&u)[NU], longlong2 *warp_smem) {
    const int lane = threadIdx.x & 31;
    constexpr int M = NU - 1;
    const int sw = NU == 8 ? (lane & 7) : ((lane >> 1) & 3);
#pragma unroll
    for (int j = 0; j < NU; j++) warp_smem[lane * NU + (j ^ sw)] = u[j];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < NU; k++) {
        const int idx = k * 32 + lane, src = idx / NU, j = idx & M;
        const int ssw = NU == 8 ? (src & 7) : ((src >> 1) & 3);
        hpgb_row_st(warp_dst + idx, warp_smem[src * NU + (j ^ ssw)]);
    }
    __syncwarp();
}

#endif  // HPGB_LAUNCH_CUH
