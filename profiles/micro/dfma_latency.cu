// Dependent-issue latency and per-SMSP throughput of DFMA / DADD on sm_100a (B200): sizes the ILP x occupancy the fp64 ray kernels need.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_latency dfma_latency.cu && ./dfma_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void k(double *out, long long *cyc, int iters, double a, double b) {
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 1e-3 + c;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) x[c] = fma(x[c], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int CHAINS>
void run(int warps_per_block) {
    double *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    k<CHAINS><<<148, 32 * warps_per_block>>>(out, cyc, iters, 0.999, 1e-3);
    k<CHAINS><<<148, 32 * warps_per_block>>>(out, cyc, iters, 0.999, 1e-3);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double per = (double)h[0] / (iters * 8.0);
    printf("chains %d warps/SM %2d: %.2f cycles per step of %d DFMA/warp -> %.2f cycles per warp-DFMA per SMSP (%d warps/SMSP)\n", CHAINS, warps_per_block,
           per, CHAINS, per / (CHAINS * (warps_per_block / 4.0 > 1 ? warps_per_block / 4.0 : 1)), warps_per_block / 4 > 0 ? warps_per_block / 4 : 1);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<1>(1); run<2>(1); run<4>(1); run<8>(1);
    run<1>(4); run<1>(8); run<1>(16); run<1>(32);
    run<2>(8); run<2>(16); run<4>(16);
    return 0;
}
