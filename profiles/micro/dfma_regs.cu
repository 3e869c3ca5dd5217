// DFMA issue rate on sm_100a as a function of how many operands are distinct register pairs (vs constant-bank operands):
// does register-file bandwidth cap the FP64 pipe below 1 warp instruction per 2 cycles per SMSP?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_regs dfma_regs.cu && ./dfma_regs
#include <cstdio>
#include <cuda_runtime.h>

// MODE 0: x = fma(x, a, b)   a, b kernel parameters (constant bank)       -> 1 register operand
// MODE 1: x = fma(x, y, b)   y a per-thread register                      -> 2 register operands
// MODE 2: x = fma(x, y, z)   y, z per-thread registers                    -> 3 register operands
// MODE 3: x = fma(y, z, x)   accumulate form, 3 register operands (the trilinear / Horner shape)
// MODE 4/5: one of the three register operands is the same register in consecutive instructions (.reuse)
template <int MODE, int CHAINS>
__global__ void k(double *out, const double *in, long long *cyc, int iters, double a, double b) {
    double x[CHAINS], y[CHAINS], z[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
        x[c] = in[threadIdx.x + 32 * c];
        y[c] = in[threadIdx.x + 32 * c + 1024];
        z[c] = in[threadIdx.x + 32 * c + 2048];
    }
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) {
                if (MODE == 0) x[c] = fma(x[c], a, b);
                if (MODE == 1) x[c] = fma(x[c], y[c], b);
                if (MODE == 2) x[c] = fma(x[c], y[c], z[c]);
                if (MODE == 3) x[c] = fma(y[c], z[(c + 1) % CHAINS], x[c]);
                if (MODE == 4) x[c] = fma(x[c], y[0], z[c]);   // operand B shared by consecutive DFMAs: operand reuse cache
                if (MODE == 5) x[c] = fma(y[0], z[c], x[c]);   // operand A shared, accumulate form
            }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE, int CHAINS>
void run(int warps_per_block) {
    double *out, *in; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8); cudaMalloc(&in, 4096 * 8);
    double h_in[4096];
    for (int i = 0; i < 4096; ++i) h_in[i] = 0.5 + 1e-4 * i;
    cudaMemcpy(in, h_in, sizeof(h_in), cudaMemcpyHostToDevice);
    const int iters = 2000;
    k<MODE, CHAINS><<<148, 32 * warps_per_block>>>(out, in, cyc, iters, 0.999, 1e-3);
    k<MODE, CHAINS><<<148, 32 * warps_per_block>>>(out, in, cyc, iters, 0.999, 1e-3);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double per_step = (double)h[0] / (iters * 8.0);
    const double warps_per_smsp = warps_per_block / 4.0;
    printf("mode %d chains %d warps/SMSP %.0f: %.2f cycles per warp-DFMA per SMSP\n", MODE, CHAINS, warps_per_smsp, per_step / (CHAINS * warps_per_smsp));
    cudaFree(out); cudaFree(cyc); cudaFree(in);
}

int main() {
    run<0, 4>(16); run<1, 4>(16); run<2, 4>(16); run<3, 4>(16);
    run<0, 2>(12); run<1, 2>(12); run<2, 2>(12); run<3, 2>(12);
    run<2, 8>(8); run<3, 8>(8);
    run<4, 4>(16); run<5, 4>(16); run<4, 8>(8); run<5, 8>(8);
    return 0;
}
