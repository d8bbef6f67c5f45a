// Validation: branch-free fp64 square root (rsqrt approximation + two coupled Newton steps + one
// residual correction, no special-case branch) against IEEE sqrt() on the operand domain of the
// pixel_to_vector hot path: x in (2^-130, 4), random mantissas, log-uniform exponents.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ double sqrt_tame(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    double r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    return fma(fma(-g, g, x), h, g);
}

__device__ __forceinline__ uint64_t rng(uint64_t &s) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    return s;
}

__global__ void k(unsigned long long *bad, unsigned long long *n, int iters, uint64_t seed) {
    uint64_t s = seed + 0x9E3779B97F4A7C15ull * (blockIdx.x * blockDim.x + threadIdx.x + 1);
    unsigned long long nb = 0;
    for (int i = 0; i < iters; i++) {
        const uint64_t r1 = rng(s), r2 = rng(s);
        const int e = 1023 - 130 + (int)(r2 % 132);
        const double x = __longlong_as_double((long long)(((uint64_t)e << 52) | (r1 >> 12)));
        if (sqrt_tame(x) != sqrt(x)) nb++;
        // the actual expressions: (1 - z)(1 + z) with z = k * fact1-like grid, t (2 - t)
        const double z = (double)(int32_t)(r1 >> 33) * (1.0 / 2147483648.0) * 0.99;
        const double a = (1.0 - z) * (1.0 + z);
        if (sqrt_tame(a) != sqrt(a)) nb++;
        const double t = __longlong_as_double((long long)(((uint64_t)(1023 - 60 + (int)(r2 % 54)) << 52) | (r2 >> 12)));
        const double b = t * (2.0 - t);
        if (sqrt_tame(b) != sqrt(b)) nb++;
    }
    atomicAdd(bad, nb);
    atomicAdd(n, 3ull * iters);
}

int main() {
    unsigned long long *d, h[2] = {0, 0};
    cudaMalloc(&d, 16);
    cudaMemset(d, 0, 16);
    for (int rep = 0; rep < 8; rep++) k<<<148 * 8, 256>>>(d, d + 1, 4096, 777 + rep);
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("sqrt_tame vs IEEE: %llu mismatches in %llu square roots\n", h[0], h[1]);
    return h[0] != 0;
}
