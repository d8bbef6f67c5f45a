// Micro-benchmark: issue throughput of integer instructions on sm_100a (which pipe, what rate).
// Each test runs N_ITER iterations of UNROLL independent chains per thread, 1024 threads/SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 4096

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed, uint32_t mul) {
    uint32_t a0 = threadIdx.x + seed, a1 = a0 * 3 + 1, a2 = a0 * 5 + 2, a3 = a0 * 7 + 3;
    uint32_t b0 = a0 ^ 0x1234, b1 = a1 ^ 0x777, b2 = a2 ^ 0x999, b3 = a3 ^ 0x4242;
#pragma unroll 1
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (MODE == 0) {  // LOP3 only (alu)
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a0) : "r"(b0), "r"(seed));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a1) : "r"(b1), "r"(seed));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a2) : "r"(b2), "r"(seed));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a3) : "r"(b3), "r"(seed));
            } else if (MODE == 1) {  // IMAD lo (fma)
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a0) : "r"(mul), "r"(b0));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a1) : "r"(mul), "r"(b1));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a2) : "r"(mul), "r"(b2));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a3) : "r"(mul), "r"(b3));
            } else if (MODE == 2) {  // IMAD.HI
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a0) : "r"(mul));
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a1) : "r"(mul));
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a2) : "r"(mul));
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a3) : "r"(mul));
            } else if (MODE == 3) {  // SHF (alu)
                asm volatile("shf.r.clamp.b32 %0, %0, %1, %2;" : "+r"(a0) : "r"(b0), "r"(seed));
                asm volatile("shf.r.clamp.b32 %0, %0, %1, %2;" : "+r"(a1) : "r"(b1), "r"(seed));
                asm volatile("shf.r.clamp.b32 %0, %0, %1, %2;" : "+r"(a2) : "r"(b2), "r"(seed));
                asm volatile("shf.r.clamp.b32 %0, %0, %1, %2;" : "+r"(a3) : "r"(b3), "r"(seed));
            } else if (MODE == 4) {  // 2 LOP3 + 2 IMAD (dual pipe)
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a0) : "r"(b0), "r"(seed));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a1) : "r"(mul), "r"(b1));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a2) : "r"(b2), "r"(seed));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a3) : "r"(mul), "r"(b3));
            } else if (MODE == 5) {  // 2 LOP3 + 2 IMAD.HI
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a0) : "r"(b0), "r"(seed));
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a1) : "r"(mul));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a2) : "r"(b2), "r"(seed));
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a3) : "r"(mul));
            } else if (MODE == 6) {  // IMAD.WIDE
                uint64_t w0 = ((uint64_t)b0 << 32) | a0, w1 = ((uint64_t)b1 << 32) | a1, w2 = ((uint64_t)b2 << 32) | a2, w3 = ((uint64_t)b3 << 32) | a3;
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w0) : "r"(a0), "r"(mul));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w1) : "r"(a1), "r"(mul));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w2) : "r"(a2), "r"(mul));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w3) : "r"(a3), "r"(mul));
                a0 = (uint32_t)w0; b0 = (uint32_t)(w0 >> 32); a1 = (uint32_t)w1; b1 = (uint32_t)(w1 >> 32);
                a2 = (uint32_t)w2; b2 = (uint32_t)(w2 >> 32); a3 = (uint32_t)w3; b3 = (uint32_t)(w3 >> 32);
            } else if (MODE == 7) {  // PRMT
                asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a0) : "r"(b0), "r"(seed));
                asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a1) : "r"(b1), "r"(seed));
                asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a2) : "r"(b2), "r"(seed));
                asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a3) : "r"(b3), "r"(seed));
            } else if (MODE == 8) {  // ISETP+SEL
                asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %1, %2, p;}" : "+r"(a0) : "r"(b0), "r"(seed));
                asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %1, %2, p;}" : "+r"(a1) : "r"(b1), "r"(seed));
                asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %1, %2, p;}" : "+r"(a2) : "r"(b2), "r"(seed));
                asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %1, %2, p;}" : "+r"(a3) : "r"(b3), "r"(seed));
            } else if (MODE == 9) {  // 3 LOP3 + 1 IMAD
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a0) : "r"(b0), "r"(seed));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a1) : "r"(b1), "r"(seed));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a2) : "r"(b2), "r"(seed));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a3) : "r"(mul), "r"(b3));
            } else if (MODE == 10) {  // 2 LOP3 + 1 IMAD + 1 DFMA-free: popc (which pipe?)
                asm volatile("popc.b32 %0, %0;" : "+r"(a0));
                asm volatile("popc.b32 %0, %0;" : "+r"(a1));
                asm volatile("popc.b32 %0, %0;" : "+r"(a2));
                asm volatile("popc.b32 %0, %0;" : "+r"(a3));
            } else if (MODE == 11) {  // brev
                asm volatile("brev.b32 %0, %0;" : "+r"(a0));
                asm volatile("brev.b32 %0, %0;" : "+r"(a1));
                asm volatile("brev.b32 %0, %0;" : "+r"(a2));
                asm volatile("brev.b32 %0, %0;" : "+r"(a3));
            } else if (MODE == 12) {  // 2 LOP3 + 2 IMAD.WIDE
                uint64_t w1 = ((uint64_t)b1 << 32) | a1, w3 = ((uint64_t)b3 << 32) | a3;
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a0) : "r"(b0), "r"(seed));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w1) : "r"(a1), "r"(mul));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a2) : "r"(b2), "r"(seed));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w3) : "r"(a3), "r"(mul));
                a1 = (uint32_t)w1; b1 = (uint32_t)(w1 >> 32); a3 = (uint32_t)w3; b3 = (uint32_t)(w3 >> 32);
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ b0 ^ b1 ^ b2 ^ b3;
}

template <int MODE>
void run(const char *name, uint32_t *d, int sms) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = sms * 4;  // 4 CTAs x 256 = 1024 threads/SM = 8 warps/SMSP
    k<MODE><<<grid, 256>>>(d, 1, 3);
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(d, 1, 3);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double winst = (double)grid * 8 * ITER * 8 * 4;  // warp instructions
    const double per_smsp_clk = winst / (sms * 4.0) / (ms * 1e-3 * clk * 1e3);
    printf("%-28s %8.3f ms   %.3f warp-inst/clk/SMSP (at nominal %d MHz)\n", name, ms, per_smsp_clk, clk / 1000);
}

int main() {
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t *d;
    cudaMalloc(&d, sms * 4 * 256 * 4);
    run<0>("LOP3", d, sms);
    run<1>("IMAD.lo", d, sms);
    run<2>("IMAD.HI", d, sms);
    run<3>("SHF", d, sms);
    run<7>("PRMT", d, sms);
    run<8>("ISETP+SEL (2 inst)", d, sms);
    run<6>("IMAD.WIDE", d, sms);
    run<10>("POPC", d, sms);
    run<11>("BREV", d, sms);
    run<4>("2 LOP3 + 2 IMAD", d, sms);
    run<9>("3 LOP3 + 1 IMAD", d, sms);
    run<5>("2 LOP3 + 2 IMAD.HI", d, sms);
    run<12>("2 LOP3 + 2 IMAD.WIDE", d, sms);
    return 0;
}
