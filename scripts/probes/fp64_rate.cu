// DFMA issue rate per SM: W warps per CTA (one CTA per SM), 16 independent accumulators per thread.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void rate(double *out, long long *clk, int iters) {
    double a[16];
    for (int k = 0; k < 16; ++k) a[k] = threadIdx.x * 1e-3 + k;
    const double b = 1.0000001, c = 1e-9;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = fma(a[k], b, c);
    }
    __syncthreads();
    const long long t1 = clock64();
    double s = 0; for (int k = 0; k < 16; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
int main() {
    double *out; long long *clk; cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&clk, 8);
    const int iters = 4096;
    for (int warps : {1, 2, 4, 8, 16, 32}) {
        rate<<<148, warps * 32>>>(out, clk, iters); rate<<<148, warps * 32>>>(out, clk, iters);
        long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
        const double fma_per_clk = (double)warps * 32 * 16 * iters / (double)h;
        printf("warps/SM %2d: %.1f DFMA lanes per clock per SM  (= %.1f TFLOP/s at 148 SMs x 1.965 GHz)\n", warps, fma_per_clk, fma_per_clk * 2 * 148 * 1.965e9 / 1e12);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
