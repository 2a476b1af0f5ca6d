// Dependent-chain latencies of the instructions on the coordinate-descent register chain (one warp, one CTA).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double sel(double r, double l1, double a, double b, double c) {
    double out;
    asm("{\n\t.reg .pred p, q;\n\tsetp.gt.f64 p, %1, %2;\n\tsetp.lt.f64 q, %1, %3;\n\tselp.f64 %0, %5, %6, q;\n\tselp.f64 %0, %4, %0, p;\n\t}"
        : "=d"(out) : "d"(r), "d"(l1), "d"(-l1), "d"(a), "d"(b), "d"(c));
    return out;
}
__global__ void probe(double *out, long long *clk, double x0, double l1, int n) {
    const int lane = threadIdx.x;
    double a = x0 + lane, b = 1.0 + 1e-9 * lane, c = 0.5;
    long long t0, t1;
    // 1: DFMA chain
    t0 = clock64();
    for (int i = 0; i < n; ++i) { a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); }
    t1 = clock64(); if (lane == 0) clk[0] = t1 - t0;
    // 2: SHFL (64-bit) chain
    t0 = clock64();
    for (int i = 0; i < n; ++i) { a = __shfl_sync(~0u, a, 1); a = __shfl_sync(~0u, a, 2); a = __shfl_sync(~0u, a, 3); a = __shfl_sync(~0u, a, 4); }
    t1 = clock64(); if (lane == 0) clk[1] = t1 - t0;
    // 3: DADD chain
    t0 = clock64();
    for (int i = 0; i < n; ++i) { a = a + b; a = a + c; a = a + b; a = a + c; }
    t1 = clock64(); if (lane == 0) clk[2] = t1 - t0;
    // 4: setp+selp chain (the select of the soft threshold)
    t0 = clock64();
    for (int i = 0; i < n; ++i) { a = sel(a, l1, b, c, a); a = sel(a, l1, c, b, a); a = sel(a, l1, b, c, a); a = sel(a, l1, c, b, a); }
    t1 = clock64(); if (lane == 0) clk[3] = t1 - t0;
    // 5: the full step: shfl -> fma(r) -> 2 fma -> select
    double r = a, inv = 0.37, kp = -0.2, kn = 0.2, s = 1e-3;
    t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double dp = fma(r, inv, kp), dn = fma(r, inv, kn);
            const double dc = sel(r, l1, dp, dn, c);
            const double di = __shfl_sync(~0u, dc, (i * 4 + k) & 31);
            r = fma(-di, s, r);
        }
    }
    t1 = clock64(); if (lane == 0) clk[4] = t1 - t0;
    // 6: DMUL chain
    t0 = clock64();
    for (int i = 0; i < n; ++i) { r = r * b; r = r * b; r = r * b; r = r * b; }
    t1 = clock64(); if (lane == 0) clk[5] = t1 - t0;
    // 7: FFMA chain (fp32) for comparison
    float f = (float)r, g = 1.0001f, h = 0.5f;
    t0 = clock64();
    for (int i = 0; i < n; ++i) { f = fmaf(f, g, h); f = fmaf(f, g, h); f = fmaf(f, g, h); f = fmaf(f, g, h); }
    t1 = clock64(); if (lane == 0) clk[6] = t1 - t0;
    // 8: 32-bit SHFL chain
    int q = lane;
    t0 = clock64();
    for (int i = 0; i < n; ++i) { q = __shfl_sync(~0u, q, 1); q = __shfl_sync(~0u, q + 1, 2); q = __shfl_sync(~0u, q + 1, 3); q = __shfl_sync(~0u, q + 1, 4); }
    t1 = clock64(); if (lane == 0) clk[7] = t1 - t0;
    out[lane] = a + r + f + q;
}
int main() {
    double *out; long long *clk;
    cudaMalloc(&out, 32 * 8 * 64); cudaMalloc(&clk, 64 * 8);
    const int n = 256;
    for (int warps = 1; warps <= 1; ++warps) {
        probe<<<1, 32>>>(out, clk, 1.0, 0.25, n);
        probe<<<1, 32>>>(out, clk, 1.0, 0.25, n);
        long long h[8];
        cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
        const char *nm[8] = {"DFMA", "SHFL64", "DADD", "SETP+SELP(f64)", "CD step", "DMUL", "FFMA", "SHFL32(+IADD)"};
        for (int k = 0; k < 8; ++k) printf("%-16s %.1f clk per op\n", nm[k], (double)h[k] / (4.0 * n));
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
