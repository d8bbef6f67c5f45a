// pipes2.cu -- measurement tool (not product): do the fp64 pipe and the integer ALU pipe of sm_100a overlap, what do
// the fp64 conversions cost, and at which SM clock does a long fp64 + integer kernel really run?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/ubench/pipes2.cu -o tools/ubench/pipes2
// Rates are per SMSP and per ACTUAL clock: every kernel reads clock64() and globaltimer at its start and end.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 2048

__device__ __forceinline__ uint64_t gtime() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, double seed, uint32_t iseed, uint64_t *clk) {
    const uint64_t c0 = clock64(), t0 = gtime();
    double a0 = threadIdx.x + seed, a1 = a0 * 3 + 1, a2 = a0 * 5 + 2, a3 = a0 * 7 + 3;
    uint32_t b0 = threadIdx.x ^ iseed, b1 = b0 * 3 + 1, b2 = b0 * 5 + 2, b3 = b0 * 7 + 3;
    long long l0 = 0, l1 = 0;
#pragma unroll 1
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (MODE == 0) {  // DFMA only
                asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(a0) : "d"(seed));
                asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(a1) : "d"(seed));
                asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(a2) : "d"(seed));
                asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(a3) : "d"(seed));
            } else if (MODE == 1) {  // 2 DFMA + 2 LOP3
                asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(a0) : "d"(seed));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b0) : "r"(b1), "r"(iseed));
                asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(a2) : "d"(seed));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b2) : "r"(b3), "r"(iseed));
            } else if (MODE == 2) {  // 2 DFMA + 2 IMAD
                asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(a0) : "d"(seed));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b0) : "r"(iseed), "r"(b1));
                asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(a2) : "d"(seed));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b2) : "r"(iseed), "r"(b3));
            } else if (MODE == 3) {  // 4 DFMA + 4 LOP3 + ... counted as 4: DFMA, LOP3, IMAD, LOP3
                asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(a0) : "d"(seed));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b0) : "r"(b1), "r"(iseed));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b2) : "r"(iseed), "r"(b3));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b1) : "r"(b3), "r"(iseed));
            } else if (MODE == 4) {  // F2I.S64.F64 (cvt.rzi.s64.f64)
                asm volatile("cvt.rzi.s64.f64 %0, %1;" : "=l"(l0) : "d"(a0));
                asm volatile("cvt.rzi.s64.f64 %0, %1;" : "=l"(l1) : "d"(a1));
                a0 += (double)0;  // keep a dependence-free stream; results folded below
                b0 ^= (uint32_t)l0;
                b1 ^= (uint32_t)l1;
                asm volatile("cvt.rzi.s64.f64 %0, %1;" : "=l"(l0) : "d"(a2));
                asm volatile("cvt.rzi.s64.f64 %0, %1;" : "=l"(l1) : "d"(a3));
                b2 ^= (uint32_t)l0;
                b3 ^= (uint32_t)l1;
            } else if (MODE == 5) {  // FRND.F64.FLOOR
                asm volatile("cvt.rmi.f64.f64 %0, %0;" : "+d"(a0));
                asm volatile("cvt.rmi.f64.f64 %0, %0;" : "+d"(a1));
                asm volatile("cvt.rmi.f64.f64 %0, %0;" : "+d"(a2));
                asm volatile("cvt.rmi.f64.f64 %0, %0;" : "+d"(a3));
            } else if (MODE == 6) {  // DADD
                asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a0) : "d"(seed));
                asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a1) : "d"(seed));
                asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a2) : "d"(seed));
                asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a3) : "d"(seed));
            } else if (MODE == 7) {  // DSETP + SEL
                asm volatile("{.reg .pred p; setp.lt.f64 p, %1, %2; selp.u32 %0, %0, %3, p;}" : "+r"(b0) : "d"(a0), "d"(seed), "r"(iseed));
                asm volatile("{.reg .pred p; setp.lt.f64 p, %1, %2; selp.u32 %0, %0, %3, p;}" : "+r"(b1) : "d"(a1), "d"(seed), "r"(iseed));
                asm volatile("{.reg .pred p; setp.lt.f64 p, %1, %2; selp.u32 %0, %0, %3, p;}" : "+r"(b2) : "d"(a2), "d"(seed), "r"(iseed));
                asm volatile("{.reg .pred p; setp.lt.f64 p, %1, %2; selp.u32 %0, %0, %3, p;}" : "+r"(b3) : "d"(a3), "d"(seed), "r"(iseed));
            } else if (MODE == 8) {  // 2 DFMA + 2 FFMA (fp32 on the fma pipe)
                float f0 = __uint_as_float(b0), f1 = __uint_as_float(b1);
                asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(a0) : "d"(seed));
                asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f0) : "f"(1.0001f));
                asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(a2) : "d"(seed));
                asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f1) : "f"(1.0001f));
                b0 = __float_as_uint(f0);
                b1 = __float_as_uint(f1);
            } else if (MODE == 9) {  // 1 DFMA + 1 F2I.S64 + 2 LOP3
                asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(a0) : "d"(seed));
                asm volatile("cvt.rzi.s64.f64 %0, %1;" : "=l"(l0) : "d"(a1));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b0) : "r"(b1), "r"(iseed));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b2) : "r"(b3), "r"((uint32_t)l0));
            } else if (MODE == 10) {  // I2F.F64.S64
                asm volatile("cvt.rn.f64.s64 %0, %1;" : "=d"(a0) : "l"((long long)b0));
                asm volatile("cvt.rn.f64.s64 %0, %1;" : "=d"(a1) : "l"((long long)b1));
                asm volatile("cvt.rn.f64.s64 %0, %1;" : "=d"(a2) : "l"((long long)b2));
                asm volatile("cvt.rn.f64.s64 %0, %1;" : "=d"(a3) : "l"((long long)b3));
                b0 += (uint32_t)__double2loint(a0);
                b1 += (uint32_t)__double2loint(a1);
                b2 += (uint32_t)__double2loint(a2);
                b3 += (uint32_t)__double2loint(a3);
            } else if (MODE == 11) {  // the a2p mix per point: 36 fp64, 42 alu, 9 imad, 4 xu -> here 4 fp64 + 5 alu + 1 imad per group
                asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(a0) : "d"(seed));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b0) : "r"(b1), "r"(iseed));
                asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(a1) : "d"(seed));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b1) : "r"(b2), "r"(iseed));
                asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(a2) : "d"(seed));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b2) : "r"(b3), "r"(iseed));
                asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(a3) : "d"(seed));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b3) : "r"(b0), "r"(iseed));
                asm volatile("shf.r.clamp.b32 %0, %0, %1, %2;" : "+r"(b0) : "r"(b2), "r"(iseed));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b1) : "r"(iseed), "r"(b3));
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + (double)(b0 ^ b1 ^ b2 ^ b3) + (double)(l0 + l1);
    if (threadIdx.x == 0) {
        clk[2 * blockIdx.x] = clock64() - c0;
        clk[2 * blockIdx.x + 1] = gtime() - t0;
    }
}

template <int MODE>
void run(const char *name, double *d, uint64_t *dclk, int sms, int per_group, int warps_per_smsp = 8, int reps = 1) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int ctas = warps_per_smsp / 2;  // 256-thread CTAs: 2 warps per SMSP each
    const int grid = sms * ctas;
    k<MODE><<<grid, 256>>>(d, 1.0000001, 3, dclk);
    cudaEventRecord(e0);
    for (int r = 0; r < reps; r++) k<MODE><<<grid, 256>>>(d, 1.0000001, 3, dclk);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    uint64_t h[2];
    cudaMemcpy(h, dclk, 16, cudaMemcpyDeviceToHost);
    const double mhz = (double)h[0] / (double)h[1] * 1e3;
    const double winst = (double)grid * 8 * ITER * 8 * per_group;
    const double per_smsp_clk = winst / (sms * 4.0) / (double)h[0] * ((double)h[1] * 1e-6 / ms);
    printf("%-34s %8.3f ms  SM clock %6.0f MHz  %.3f warp-inst/clk/SMSP  (%d warps/SMSP)\n", name, ms, mhz, per_smsp_clk, warps_per_smsp);
}

int main(int argc, char **argv) {
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *d;
    uint64_t *dclk;
    cudaMalloc(&d, sms * 8 * 256 * 8);
    cudaMalloc(&dclk, sms * 8 * 16);
    run<0>("DFMA", d, dclk, sms, 4);
    run<6>("DADD", d, dclk, sms, 4);
    run<1>("2 DFMA + 2 LOP3", d, dclk, sms, 4);
    run<2>("2 DFMA + 2 IMAD", d, dclk, sms, 4);
    run<3>("DFMA + 2 LOP3 + IMAD", d, dclk, sms, 4);
    run<8>("2 DFMA + 2 FFMA", d, dclk, sms, 4);
    run<4>("F2I.S64.F64", d, dclk, sms, 4);
    run<5>("FRND.F64.FLOOR", d, dclk, sms, 4);
    run<10>("I2F.F64.S64", d, dclk, sms, 4);
    run<7>("DSETP + SEL (2 inst)", d, dclk, sms, 8);
    run<9>("DFMA + F2I + 2 LOP3", d, dclk, sms, 4);
    run<11>("a2p-like mix (4 DFMA 5 ALU 1 IMAD)", d, dclk, sms, 10);
    run<11>("a2p-like mix, 4 warps/SMSP", d, dclk, sms, 10, 4);
    run<11>("a2p-like mix, 2 warps/SMSP", d, dclk, sms, 10, 2);
    // clock under a sustained fp64 + integer load: the same kernel back to back for ~0.3 s, per-launch clock
    printf("sustained a2p-like mix, 4 warps/SMSP:\n");
    for (int r = 0; r < 12; r++) run<11>("  burst of 8 launches", d, dclk, sms, 10, 4, 8);
    return 0;
}
