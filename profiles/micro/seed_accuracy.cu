// Accuracy of the MUFU.RCP64H / MUFU.RSQ64H seeds (rcp.approx.ftz.f64 / rsqrt.approx.ftz.f64) and of the one-step cubic
// (Halley-type) refinements used by geodesy.cuh, against IEEE division / sqrt.   nvcc -arch=sm_100a -O3 seed_accuracy.cu
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ double seed_rcp(double a) { double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a)); return y; }
__device__ __forceinline__ double seed_rsq(double a) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a)); return y; }

__device__ __forceinline__ double rcp3(double a) {  // y(1 + e + e^2), e = 1 - a y : cubic
    const double y = seed_rcp(a);
    const double e = fma(-a, y, 1.0);
    return fma(y, fma(e, e, e), y);
}
__device__ __forceinline__ double rsq3(double a) {  // y(1 + e/2 + 3e^2/8), e = 1 - a y^2 : cubic
    const double y = seed_rsq(a);
    const double e = fma(-a * y, y, 1.0);
    return fma(y, e * fma(0.375, e, 0.5), y);
}
__device__ __forceinline__ double rcp2(double a) {  // one Newton step
    const double y = seed_rcp(a);
    return fma(y, fma(-a, y, 1.0), y);
}

__global__ void k(unsigned long long n, double lo, double hi, double *out) {
    double m[5] = {0, 0, 0, 0, 0};
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long x = 0x9E3779B97F4A7C15ull * (i + 1);
        x ^= x >> 12; x ^= x << 25; x ^= x >> 27; x *= 0x2545F4914F6CDD1Dull;
        const double u = (double)(x >> 11) * 0x1p-53;
        const double a = lo * exp(u * log(hi / lo));
        const double r = 1.0 / a, s = 1.0 / sqrt(a);
        m[0] = fmax(m[0], fabs(seed_rcp(a) - r) / r);
        m[1] = fmax(m[1], fabs(seed_rsq(a) - s) / s);
        m[2] = fmax(m[2], fabs(rcp3(a) - r) / r);
        m[3] = fmax(m[3], fabs(rsq3(a) - s) / s);
        m[4] = fmax(m[4], fabs(rcp2(a) - r) / r);
    }
    for (int j = 0; j < 5; ++j) {
        unsigned long long b = __double_as_longlong(m[j]);
        atomicMax((unsigned long long *)out + j, b);
    }
}

int main() {
    double *d, h[5];
    cudaMalloc(&d, sizeof(h));
    const double ranges[4][2] = {{0.5, 2.0}, {1e-3, 1e3}, {1e10, 1e16}, {0.99, 1.01}};
    for (auto &r : ranges) {
        cudaMemset(d, 0, sizeof(h));
        k<<<148 * 8, 256>>>(1ull << 30, r[0], r[1], d);
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("a in [%g, %g]: seed rcp %.3e (2^%.1f)  seed rsqrt %.3e (2^%.1f)  rcp cubic %.3e  rsqrt cubic %.3e  rcp newton1 %.3e\n", r[0], r[1], h[0],
               log2(h[0]), h[1], log2(h[1]), h[2], h[3], h[4]);
    }
    return cudaGetLastError() != cudaSuccess;
}
