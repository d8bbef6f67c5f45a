// a2p_sweep.cu -- measurement tool (not product): the headline angle_to_pixel kernel under different
// launch shapes of the streaming skeleton, next to two ceilings of the SAME skeleton:
//   copy     : same tiles, same TMA ring, same stores, trivial arithmetic  -> what HBM allows at 16 B in + 8 B out
//   nomem    : (build with -DHPGB_STREAM_NOMEM) the ring is filled once, the arithmetic and the stores run
//              at full length                                            -> what the instruction stream allows
// and under different placements of the three operand arrays in HBM.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -diag-suppress 177 \
//        [-DHPGB_A2P_MAGIC] [-DHPGB_A2P_QUEUE -DHPGB_A2P_QBLOCK=384] [-DHPGB_STREAM_NOMEM] \
//        tools/ubench/a2p_sweep.cu -o tools/ubench/a2p_sweep
//   tools/ubench/a2p_sweep [log2n=30] [reps=10]
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../hpgeom_b200/csrc/k_a2p.cu"

std::atomic<int64_t> g_hpgb_launches{0};
int hpgb_sm_count() {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v > 0 ? v : 148;
}

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// uniform on the sphere, degrees (BASELINE config 5 distribution)
__global__ void k_fill(double *lon, double *lat, int64_t n, uint64_t seed) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t a = mix64(seed + 2 * (uint64_t)i), b = mix64(seed + 2 * (uint64_t)i + 1);
        const double u1 = (double)(a >> 11) * 0x1p-53, u2 = (double)(b >> 11) * 0x1p-53;
        lon[i] = 360.0 * u1;
        lat[i] = asin(2.0 * u2 - 1.0) * 57.29577951308232;
    }
}
__global__ void k_sum(const int64_t *p, int64_t n, unsigned long long *acc) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long s = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        s += (unsigned long long)p[i] * (unsigned long long)(2 * (i & 1023) + 1);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(acc, s);
}

// the memory ceiling of the skeleton: same operands, same tiles, (almost) no arithmetic
struct OpCopy {
    const double *a, *b;
    int64_t *out;
    struct In2 {
        double2 a, b;
    };
    typedef hpgb_no_ctx Ctx;
    HPGB_STREAM_IN_F64x2(a, b, 2)
    __device__ __forceinline__ Ctx prologue() const { return Ctx(); }
    __device__ __forceinline__ In2 load2(int64_t i) const { return In2{hpgb_ld2(a + i), hpgb_ld2(b + i)}; }
    __device__ __forceinline__ void run2(int64_t i, const In2 &v, const Ctx &, hpgb_bad_t &) const {
        hpgb_st2(out + i, __double_as_longlong(v.a.x) ^ __double_as_longlong(v.b.x),
                 __double_as_longlong(v.a.y) ^ __double_as_longlong(v.b.y));
    }
    HPGB_DEFAULT_RUN4
    __device__ __forceinline__ void run1(int64_t i, const Ctx &, hpgb_bad_t &) const {
        out[i] = __double_as_longlong(a[i]) ^ __double_as_longlong(b[i]);
    }
};

typedef OpAng2Pix<true, true, true, false, true> OpHead;  // nest, lonlat, degrees, scalar nside >= 2^16

struct Timing {
    float ms;
    unsigned long long sum;
};

cudaEvent_t e0, e1;
unsigned long long *d_acc;
hpgb_status_t *d_st;

template <class Op, int ST, int H, int BLOCK, int MINB, bool CONTIG>
Timing run_cfg(const Op &op, int64_t n, int reps, int64_t *out, const char *name, const char *layout) {
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, k_stream<Op, ST, H, BLOCK, MINB, CONTIG>));
    const int smem = ST * Op::NIN * 4 * BLOCK * H * 8 + Op::EXTRA_SMEM;
    CK(cudaFuncSetAttribute(k_stream<Op, ST, H, BLOCK, MINB, CONTIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_stream<Op, ST, H, BLOCK, MINB, CONTIG>, BLOCK, smem));
    CK(cudaMemset(d_st, 0xFF, sizeof(hpgb_status_t)));
    for (int i = 0; i < 3; i++) {
        int rc = hpgb_launch_stream<Op, ST, H, BLOCK, MINB, CONTIG>(op, n, d_st, 0);
        if (rc) {
            printf("%-34s launch failed rc=%d\n", name, rc);
            return Timing{0, 0};
        }
    }
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; i++) hpgb_launch_stream<Op, ST, H, BLOCK, MINB, CONTIG>(op, n, d_st, 0);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    CK(cudaMemset(d_acc, 0, 8));
    k_sum<<<148 * 8, 256>>>(out, n, d_acc);
    unsigned long long sum = 0, bad = 0;
    CK(cudaMemcpy(&sum, d_acc, 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&bad, d_st, 8, cudaMemcpyDeviceToHost));
    printf("%-34s %-9s regs %3d occ %d smem %6d | %8.3f ms %7.1f Gpts/s %6.0f GB/s  sum %016llx%s\n", name, layout, fa.numRegs, occ,
           smem, ms, n / ms / 1e6, 24.0 * n / ms / 1e6, sum, bad == ~0ull ? "" : "  (status: bad element!)");
    fflush(stdout);
    return Timing{ms, sum};
}

}  // namespace

int main(int argc, char **argv) {
    const int log2n = argc > 1 ? atoi(argv[1]) : 30;
    const int reps = argc > 2 ? atoi(argv[2]) : 10;
    const int64_t n = (int64_t)1 << log2n;
    const int64_t nside = (int64_t)1 << 29;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaMalloc(&d_acc, 8));
    CK(cudaMalloc(&d_st, sizeof(hpgb_status_t)));
    // layouts: "carved" = lon | lat | pix back to back in one allocation (bench.py round 1);
    //          "separate" = three cudaMalloc'ed arrays (what three torch.empty calls give);
    //          "skewed" = one allocation, lat shifted by 1 MiB + 16 KiB, pix by 3 MiB + 48 KiB
    char *pool = nullptr, *sa = nullptr, *sb = nullptr, *so = nullptr;
    const size_t arr = (size_t)n * 8, pad = (size_t)8 << 20;
    CK(cudaMalloc(&pool, 3 * arr + 2 * pad));
    CK(cudaMalloc(&sa, arr));
    CK(cudaMalloc(&so, arr + (2 << 20)));  // allocation order differs from operand order on purpose
    CK(cudaMalloc(&sb, arr));
    struct Layout {
        const char *name;
        double *lon, *lat;
        int64_t *out;
    } layouts[3] = {
        {"carved", (double *)pool, (double *)(pool + arr), (int64_t *)(pool + 2 * arr)},
        {"separate", (double *)sa, (double *)sb, (int64_t *)so},
        {"skewed", (double *)pool, (double *)(pool + arr + (1 << 20) + (16 << 10)), (int64_t *)(pool + 2 * arr + (4 << 20) + (48 << 10))},
    };
    printf("# a2p_sweep n=2^%d reps=%d variant:%s%s%s  pool %p sep %p %p %p\n", log2n, reps,
#if defined(HPGB_A2P_MAGIC)
           " MAGIC",
#else
           "",
#endif
#if defined(HPGB_A2P_QUEUE)
           " QUEUE",
#else
           "",
#endif
#if defined(HPGB_STREAM_NOMEM)
           " NOMEM(compute ceiling)",
#else
           "",
#endif
           (void *)pool, (void *)sa, (void *)sb, (void *)so);
    const hpgb_geom g = hpgb_make_geom(nside, 1);
    const int nl = argc > 3 ? atoi(argv[3]) : 3;
    for (int l = 0; l < nl; l++) {
        const Layout &L = layouts[l];
        k_fill<<<148 * 8, 256>>>(L.lon, L.lat, n, 12345);
        CK(cudaDeviceSynchronize());
        OpHead op{L.lon, L.lat, L.out, nullptr, 0, g, hpgb_make_a2p_fast<true, true>(g), hpgb_make_rk(g)};
        OpCopy cp{L.lon, L.lat, L.out};
        run_cfg<OpCopy, 3, 2, 256, 2, false>(cp, n, reps, L.out, "copy   st3 h2 b256 minb2", L.name);
        const Timing base = run_cfg<OpHead, 3, 2, 256, 2, false>(op, n, reps, L.out, "a2p    st3 h2 b256 minb2 (base)", L.name);
#define CFG(ST, H, B, M, C, NAME)                                                              \
    do {                                                                                       \
        const Timing t_ = run_cfg<OpHead, ST, H, B, M, C>(op, n, reps, L.out, NAME, L.name);   \
        if (t_.sum != base.sum) printf("    ^^^^ RESULT DIFFERS FROM BASE\n");                 \
    } while (0)
        CFG(3, 2, 128, 4, false, "a2p    st3 h2 b128 minb4");
        CFG(3, 2, 128, 3, false, "a2p    st3 h2 b128 minb3");
        CFG(2, 2, 320, 2, false, "a2p    st2 h2 b320 minb2");
        CFG(2, 2, 256, 3, false, "a2p    st2 h2 b256 minb3");
        CFG(3, 2, 256, 1, false, "a2p    st3 h2 b256 minb1 (occ 2 by smem)");
        CFG(3, 1, 128, 6, false, "a2p    st3 h1 b128 minb6");
        CFG(3, 2, 192, 2, false, "a2p    st3 h2 b192 minb2");
#undef CFG
    }
    return 0;
}
