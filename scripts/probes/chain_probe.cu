// What does the 32-step register chain of the coordinate-descent kernels cost, alone and next to other warps?
// A faithful copy of the dense register phase (cd_cluster.cuh: candidate / cd_soft_select / shuffle / residual FMA,
// diagonal sub-block in shared memory) in a 12-warp CTA: `n_seq` register warps (warp ids 0, 4, 8, ... = one SM
// sub-partition, or 0, 1, 2, ... when spread = 1) run `blocks` blocks each; the other warps either idle or run a
// panel-like FP64 stream (loads from a global buffer + 8 FMAs per 16 bytes).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double soft_select(double r, double l1, double dpos, double dneg, double negw) {
    double out;
    asm("{\n\t.reg .pred p, q;\n\tsetp.gt.f64 p, %1, %2;\n\tsetp.lt.f64 q, %1, %3;\n\tselp.f64 %0, %5, %6, q;\n\tselp.f64 %0, %4, %0, p;\n\t}"
        : "=d"(out) : "d"(r), "d"(l1), "d"(-l1), "d"(dpos), "d"(dneg), "d"(negw));
    return out;
}
__global__ void __launch_bounds__(384, 1)
chain(const double *__restrict__ Q, double *out, long long *clk, int blocks, int n_seq, int spread, int panel, int variant) {
    __shared__ double S[2][1024];
    __shared__ double w_s[4][64], Qw_s[4][64];
    __shared__ __align__(16) double cst[4][32][4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 2048; i += 384) (&S[0][0])[i] = (i % 33 == 0) ? 1.0 : 1e-3 * ((i * 7) % 13 - 6);
    if (tid < 256) { (&w_s[0][0])[tid] = 0.01 * (tid % 7 - 3); (&Qw_s[0][0])[tid] = 0.02 * (tid % 5 - 2); }
    __syncthreads();
    const bool is_seq = spread ? (warp < n_seq) : ((warp & 3) == 0 && (warp >> 2) < n_seq);
    const int m = spread ? warp : (warp >> 2);
    if (is_seq) {
        const double l1 = 0.013, l2 = 0.4;
        double acc = 0.0;
        long long t_loop = 0;
        const long long t0 = clock64();
        for (int b = 0; b < blocks; ++b) {
            const double *Sb = S[b & 1];
            const double d_l = 1.0 + 1e-3 * lane, q_l = 0.05 * (lane - 16) + 1e-6 * b;
            const double w_l = w_s[m][lane + 32 * (b & 1)];
            double Qw_l = Qw_s[m][lane + 32 * (b & 1)] + acc * 1e-9;
            const double inv_l = 1.0 / (d_l + l2);
            const double a_l = fma(w_l, d_l, q_l);
            const double negw = -w_l;
            const double k_pos = fma(-l1, inv_l, -w_l), k_neg = fma(l1, inv_l, -w_l);
            double r_l = a_l - Qw_l, delta_l = 0.0;
            const long long tb = clock64();
            if (variant == 0) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const double s_il = Sb[i * 32 + lane];
                    const double dpos = fma(r_l, inv_l, k_pos), dneg = fma(r_l, inv_l, k_neg);
                    const double dc = soft_select(r_l, l1, dpos, dneg, negw);
                    const double di = __shfl_sync(0xffffffffu, dc, i);
                    if (lane == i) delta_l = dc;
                    r_l = fma(-di, s_il, r_l);
                }
            } else if (variant == 3) {
                // variant 3: the pivot chain runs REDUNDANTLY in every lane (uniform): lane t+1's residual is broadcast one step
                // early (before the current pivot's delta is known) and every lane applies the last update itself, so the
                // shuffle leaves the dependent chain: delta_{t-1} -> FMA -> soft threshold -> delta_t.  Same operations in the
                // same order on the same operands: same bits.
                cst[m][lane][0] = inv_l; cst[m][lane][1] = k_pos; cst[m][lane][2] = k_neg; cst[m][lane][3] = negw;
                __syncwarp();
                double bc = __shfl_sync(0xffffffffu, r_l, 0), d_prev = 0.0, s_prev = 0.0;
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                    const double bc_next = __shfl_sync(0xffffffffu, r_l, (t + 1) & 31);
                    const double2 c01 = *reinterpret_cast<const double2 *>(&cst[m][t][0]);
                    const double2 c23 = *reinterpret_cast<const double2 *>(&cst[m][t][2]);
                    const double s_tl = Sb[t * 32 + lane];
                    const double s_next = Sb[t * 32 + ((t + 1) & 31)];
                    const double rho = fma(-d_prev, s_prev, bc);
                    const double dpos = fma(rho, c01.x, c01.y), dneg = fma(rho, c01.x, c23.x);
                    const double d_t = soft_select(rho, l1, dpos, dneg, c23.y);
                    r_l = fma(-d_t, s_tl, r_l);
                    if (lane == t) delta_l = d_t;
                    d_prev = d_t; s_prev = s_next; bc = bc_next;
                }
            } else if (variant == 1) {
                // variant 1: choose the CONSTANT of the linear branch first (two compares in parallel, one select), then ONE
                // FMA, then the select of the "stays at zero" branch: same expressions, hence the same bits, as variant 0
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const double s_il = Sb[i * 32 + lane];
                    double dc;
                    asm("{\n\t.reg .pred p, q, pq;\n\t.reg .f64 k, d;\n\t"
                        "setp.gt.f64 p, %1, %2;\n\tsetp.lt.f64 q, %1, %3;\n\tor.pred pq, p, q;\n\t"
                        "selp.f64 k, %5, %6, p;\n\tfma.rn.f64 d, %1, %4, k;\n\tselp.f64 %0, d, %7, pq;\n\t}"
                        : "=d"(dc) : "d"(r_l), "d"(l1), "d"(-l1), "d"(inv_l), "d"(k_pos), "d"(k_neg), "d"(negw));
                    const double di = __shfl_sync(0xffffffffu, dc, i);
                    if (lane == i) delta_l = dc;
                    r_l = fma(-di, s_il, r_l);
                }
            } else {
                // variant 2: as 1, but the zero branch through the FMA itself (inv -> 0, constant -> -w): no final select
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const double s_il = Sb[i * 32 + lane];
                    double dc;
                    asm("{\n\t.reg .pred p, q, pq;\n\t.reg .f64 k, v;\n\t"
                        "setp.gt.f64 p, %1, %2;\n\tsetp.lt.f64 q, %1, %3;\n\tor.pred pq, p, q;\n\t"
                        "selp.f64 k, %6, %7, q;\n\tselp.f64 k, %5, k, p;\n\tselp.f64 v, %4, 0d0000000000000000, pq;\n\t"
                        "fma.rn.f64 %0, %1, v, k;\n\t}"
                        : "=d"(dc) : "d"(r_l), "d"(l1), "d"(-l1), "d"(inv_l), "d"(k_pos), "d"(k_neg), "d"(negw));
                    const double di = __shfl_sync(0xffffffffu, dc, i);
                    if (lane == i) delta_l = dc;
                    r_l = fma(-di, s_il, r_l);
                }
            }
            t_loop += clock64() - tb;
            w_s[m][lane + 32 * (b & 1)] = w_l + delta_l;
            acc += delta_l + r_l;
        }
        const long long t1 = clock64();
        if (lane == 0) { clk[2 * m] = t1 - t0; clk[2 * m + 1] = t_loop; }
        out[tid] = acc;
    } else if (panel) {
        // panel-like stream: 8 loads of 16 bytes in flight, 8 FMAs per load, until the register warps are done (fixed count)
        const double2 *Q2 = reinterpret_cast<const double2 *>(Q) + (size_t)blockIdx.x * (1 << 20) + tid;
        double2 a[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
        const int iters = blocks * panel;
        for (int it = 0; it < iters; ++it) {
            double2 v[8];
#pragma unroll
            for (int g = 0; g < 8; ++g) v[g] = __ldg(Q2 + (size_t)((it * 8 + g) & 1023) * 512);
#pragma unroll
            for (int g = 0; g < 8; ++g)
#pragma unroll
                for (int mm = 0; mm < 4; ++mm) { a[mm].x = fma(1e-3 * (mm + 1), v[g].x, a[mm].x); a[mm].y = fma(1e-3 * (mm + 1), v[g].y, a[mm].y); }
        }
        out[384 + tid] = a[0].x + a[1].y + a[2].x + a[3].y;
    }
}
int main() {
    double *Q, *out; long long *clk;
    cudaMalloc(&Q, (size_t)148 * (1 << 20) * 16 + (1 << 24)); cudaMemset(Q, 0, (size_t)148 * (1 << 20) * 16 + (1 << 24));
    cudaMalloc(&out, 1 << 16); cudaMalloc(&clk, 64 * 8);
    const int blocks = 2000;
    printf("clocks per 32-coordinate block (total / inner loop only), register warp 0\n");
    for (int variant : {0, 3})
        for (int n_seq : {1, 4})
            for (int spread : {0, 1})
                for (int panel : {0, 4}) {
                    const int grid = 148;
                    chain<<<grid, 384>>>(Q, out, clk, blocks, n_seq, spread, panel, variant);
                    chain<<<grid, 384>>>(Q, out, clk, blocks, n_seq, spread, panel, variant);
                    long long h[8];
                    cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
                    double chk[2];
                    cudaMemcpy(chk, out, sizeof chk, cudaMemcpyDeviceToHost);
                    printf("variant %d  register warps %d (%s)  panel warps %-10s : %7.0f / %7.0f clk per block   check %.17g\n", variant, n_seq,
                           spread ? "one per sub-partition" : "same sub-partition   ", panel == 0 ? "idle" : "streaming",
                           (double)h[0] / blocks, (double)h[1] / blocks, chk[0]);
                }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
