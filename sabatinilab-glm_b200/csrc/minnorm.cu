// Minimum-norm least squares for rank-deficient / near-singular normal equations — the alpha == 0 branch of the
// reference (backend/sglm.py:96-101 -> sklearn LinearRegression -> scipy.linalg.lstsq(X, y, cond=tol), tol = 1e-6:
// singular values of the centred X below tol * s_max are dropped and the minimum-norm solution is returned,
// sklearn/linear_model/_base.py:750-753).  Lag designs are routinely rank deficient (one-hot indicator groups that
// sum to the intercept, duplicated lags), and a Cholesky factorisation of such a Gram matrix "succeeds" on pivots of
// rounding-noise size — so the solver flags small pivots and this kernel takes over, entirely on the device:
//
//   one-sided Jacobi (Hestenes) on the rows of U = A = X_c' X_c: plane rotations R = ... R_2 R_1 applied from the left
//   until the rows of U = R A are mutually orthogonal; then |lambda_i| = ||U_i||, eigenvector v_i = row i of R,
//   lambda_i = U_i . v_i, and  w = sum over { i : ||U_i|| > tol^2 max_k ||U_k|| } of  v_i (v_i . q) / lambda_i
//   (eigenvalues of X_c' X_c are the squared singular values of X_c: s_i > tol s_max  <=>  lambda_i > tol^2 lambda_max).
//
// Round-robin pair ordering: C/2 disjoint row pairs per round, one CTA per pair, all CTAs of a cooperative launch
// meet at a grid barrier between rounds; U and V (2 x 8 C^2 bytes: 64 MB at C = 2000) stay in L2.  A rare path
// (cost ~ C^3 x sweeps) that replaces the host read-back + library eigensolver of round 1; the systems whose
// Cholesky pivots were all healthy leave immediately (device-side status check, no host synchronisation).
#include "common.cuh"

namespace sglm {

constexpr int MN_THREADS = 256;

struct MnCtl {                 // device scalars
    unsigned bar;              // grid barrier arrivals (monotonic)
    unsigned rotations;        // rotations applied in the current sweep
    unsigned long long smax;   // bits of max_i ||U_i||
    double fro2;               // ||A||_F^2 (invariant under the rotations): scale of the null-row floor
    int sweeps, rank;
};

__device__ __forceinline__ void mn_grid_barrier(unsigned *bar, unsigned n_blocks, unsigned &gen) {
    __syncthreads();
    ++gen;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        const unsigned target = gen * n_blocks;
        const long long t0 = clock64();
        while (*reinterpret_cast<volatile unsigned *>(bar) < target) {
            __nanosleep(32);
            if (clock64() - t0 > 20000000000LL) asm volatile("trap;");      // a protocol bug must not hang the GPU
        }
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ double mn_block_sum(double v, double *sh) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < MN_THREADS / 32; ++w) s += sh[w];
    return s;
}

// status: device int32 (of the Cholesky solve).  The kernel runs only when *status != 0 and writes 0 back.
__global__ void __launch_bounds__(MN_THREADS)
minnorm_jacobi_kernel(const double *__restrict__ A, long long lda, const double *__restrict__ q, int n, double rcond,
                      double shift, int status_mask, int max_sweeps, int *__restrict__ status, double *__restrict__ w_out,
                      double *__restrict__ U, double *__restrict__ V, long long ld, double *__restrict__ coef,
                      MnCtl *__restrict__ ctl) {
    if ((*reinterpret_cast<volatile int *>(status) & status_mask) == 0) return;      // uniform over the grid: nobody reaches a barrier
    __shared__ double sh[3][MN_THREADS / 32];
    __shared__ double rot[2];
    const unsigned nb = gridDim.x;
    unsigned gen = 0;
    const int tid = threadIdx.x;
    // U = A (rows), V = I
    double f2 = 0.0;
    for (long long e = (long long)blockIdx.x * MN_THREADS + tid; e < (long long)n * n; e += (long long)nb * MN_THREADS) {
        const int i = (int)(e / n), j = (int)(e - (long long)i * n);
        const double a = A[i * lda + j];
        U[i * ld + j] = a;
        V[i * ld + j] = (i == j) ? 1.0 : 0.0;
        f2 = fma(a, a, f2);
    }
    f2 = mn_block_sum(f2, sh[0]);
    if (tid == 0) atomicAdd(&ctl->fro2, f2);
    mn_grid_barrier(&ctl->bar, nb, gen);
    // Rows that have shrunk to rounding noise (norm <= 1e-15 ||A||_F: the null space of a rank-deficient Gram, far
    // below the cut-off rcond^2 lambda_max >= 1e-12 ||A||_F / sqrt(n)) take no further part: rotating noise against
    // noise never settles, and these directions are dropped from the solution anyway.
    const double null_floor = 1e-30 * *reinterpret_cast<volatile double *>(&ctl->fro2);
    const int m = n + (n & 1);                     // even number of players (the last one is a bye when n is odd)
    const int half = m / 2;
    int sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        for (int r = 0; r < m - 1; ++r) {
            for (int k = blockIdx.x; k < half; k += nb) {
                int a, b;
                if (k == 0) { a = m - 1; b = r; }
                else { a = (r + k) % (m - 1); b = (r - k + (m - 1)) % (m - 1); }
                if (a >= n || b >= n) continue;     // bye
                const int p = min(a, b), qq = max(a, b);
                double *Up = U + (long long)p * ld, *Uq = U + (long long)qq * ld;
                double al = 0.0, be = 0.0, ga = 0.0;
                for (int j = tid; j < n; j += MN_THREADS) {
                    const double x = __ldcg(Up + j), y = __ldcg(Uq + j);
                    al = fma(x, x, al); be = fma(y, y, be); ga = fma(x, y, ga);
                }
                al = warp_sum(al); be = warp_sum(be); ga = warp_sum(ga);
                __syncthreads();
                if ((tid & 31) == 0) { sh[0][tid >> 5] = al; sh[1][tid >> 5] = be; sh[2][tid >> 5] = ga; }
                __syncthreads();
                if (tid == 0) {
                    double sa = 0, sb = 0, sg = 0;
                    for (int w = 0; w < MN_THREADS / 32; ++w) { sa += sh[0][w]; sb += sh[1][w]; sg += sh[2][w]; }
                    double c = 1.0, s = 0.0;
                    // rotate when the rows are not orthogonal to working precision (relative to their norms)
                    if (fmin(sa, sb) > null_floor && fabs(sg) > 1e-14 * sqrt(sa * sb)) {
                        const double zeta = (sb - sa) / (2.0 * sg);
                        const double t = (fabs(zeta) > 1e100) ? 0.5 / zeta
                                                              : copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                        c = 1.0 / sqrt(1.0 + t * t);
                        s = c * t;
                        atomicAdd(&ctl->rotations, 1u);
                    }
                    rot[0] = c; rot[1] = s;
                }
                __syncthreads();
                const double c = rot[0], s = rot[1];
                if (s != 0.0) {
                    double *Vp = V + (long long)p * ld, *Vq = V + (long long)qq * ld;
                    for (int j = tid; j < n; j += MN_THREADS) {
                        const double x = __ldcg(Up + j), y = __ldcg(Uq + j);
                        Up[j] = c * x - s * y;
                        Uq[j] = s * x + c * y;
                        const double vx = __ldcg(Vp + j), vy = __ldcg(Vq + j);
                        Vp[j] = c * vx - s * vy;
                        Vq[j] = s * vx + c * vy;
                    }
                }
            }
            mn_grid_barrier(&ctl->bar, nb, gen);
        }
        const unsigned rots = *reinterpret_cast<volatile unsigned *>(&ctl->rotations);
        mn_grid_barrier(&ctl->bar, nb, gen);         // everybody has read the count
        if (blockIdx.x == 0 && tid == 0) ctl->rotations = 0u;
        mn_grid_barrier(&ctl->bar, nb, gen);
        if (rots == 0u) break;
    }
    // sigma_i = ||U_i||, lambda_i = U_i . V_i, coef_i = (V_i . q) / lambda_i for the kept directions
    for (int i = blockIdx.x; i < n; i += nb) {
        double ss = 0.0;
        for (int j = tid; j < n; j += MN_THREADS) { const double x = __ldcg(U + (long long)i * ld + j); ss = fma(x, x, ss); }
        ss = mn_block_sum(ss, sh[0]);
        if (tid == 0) atomicMax(&ctl->smax, (unsigned long long)__double_as_longlong(sqrt(ss)));
    }
    mn_grid_barrier(&ctl->bar, nb, gen);
    const double smax = __longlong_as_double((long long)*reinterpret_cast<volatile unsigned long long *>(&ctl->smax));
    // least squares (shift == 0): relative cut-off rcond^2 * lambda_max; ridge (shift = alpha > 0): the absolute
    // cut-off s > 1e-15 of sklearn's SVD fallback (_ridge.py:_solve_svd)
    const double thr = (shift > 0.0) ? 1e-30 : rcond * rcond * smax;
    for (int i = blockIdx.x; i < n; i += nb) {
        double ss = 0.0, lam = 0.0, vq = 0.0;
        for (int j = tid; j < n; j += MN_THREADS) {
            const double u = __ldcg(U + (long long)i * ld + j), v = __ldcg(V + (long long)i * ld + j);
            ss = fma(u, u, ss); lam = fma(u, v, lam); vq = fma(v, q[j], vq);
        }
        ss = mn_block_sum(ss, sh[0]); lam = mn_block_sum(lam, sh[1]); vq = mn_block_sum(vq, sh[2]);
        if (tid == 0) {
            const bool keep = sqrt(ss) > thr && lam != 0.0;
            coef[i] = keep ? vq / (lam + shift) : 0.0;
            if (keep) atomicAdd(&ctl->rank, 1);
        }
    }
    mn_grid_barrier(&ctl->bar, nb, gen);
    for (int j = blockIdx.x * MN_THREADS + tid; j < n; j += nb * MN_THREADS) {
        double acc = 0.0;
        for (int i = 0; i < n; ++i) acc = fma(__ldcg(coef + i), __ldcg(V + (long long)i * ld + j), acc);
        w_out[j] = acc;
    }
    if (blockIdx.x == 0 && tid == 0) { ctl->sweeps = sweep; *status = 0; }
}

}  // namespace sglm

using namespace sglm;

extern "C" size_t sglm_ols_minnorm_workspace_bytes(int32_t C) {
    const size_t ld = ((size_t)C + 7) & ~(size_t)7;
    return (2 * (size_t)C * ld + (size_t)C + 32) * sizeof(double);
}

// status (device int32): the solve runs only when (*status & status_mask) != 0 (the Cholesky solver flagged a
// non-positive pivot: bit 0, or a rounding-noise pivot: bit 1) and clears *status; otherwise the launch returns at once.
// shift = 0: least squares with the relative cut-off rcond; shift = alpha > 0: the Ridge system (Qc + alpha I) w = qc
// through the same decomposition (sklearn falls back to its SVD solver when the Cholesky solve of a Ridge fails).
extern "C" int sglm_ols_minnorm_f64(const double *Qc, int64_t ldq, const double *qc, int32_t C, double rcond,
                                    double shift, int32_t status_mask, int32_t *status, double *w, void *work,
                                    size_t work_bytes, void *stream) {
    SGLM_CHECK_ARG(C > 0 && ldq >= C, SGLM_E_SHAPE, "ols_minnorm: bad shape");
    SGLM_CHECK_ARG(Qc && qc && status && w && work, SGLM_E_INVALID_ARG, "ols_minnorm: null pointer");
    SGLM_CHECK_ARG(work_bytes >= sglm_ols_minnorm_workspace_bytes(C), SGLM_E_WORKSPACE, "ols_minnorm: workspace too small");
    SGLM_CHECK_ARG(rcond >= 0.0 && rcond < 1.0, SGLM_E_INVALID_ARG, "ols_minnorm: rcond out of range");
    cudaStream_t st = (cudaStream_t)stream;
    const long long ld = ((long long)C + 7) & ~7LL;
    double *U = (double *)work, *V = U + (size_t)C * ld, *coef = V + (size_t)C * ld;
    MnCtl *ctl = (MnCtl *)(coef + C);
    SGLM_CUDA_OK(cudaMemsetAsync(ctl, 0, sizeof(MnCtl), st));
    int per_sm = 0;
    SGLM_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, minnorm_jacobi_kernel, MN_THREADS, 0));
    SGLM_CHECK_ARG(per_sm >= 1, SGLM_E_CUDA, "ols_minnorm: kernel does not fit an SM");
    int grid = std::min(sm_count() * std::min(per_sm, 4), std::max(1, (C + 1) / 2));
    const double *A = Qc; long long lda = ldq; const double *qv = qc; int n = C; double rc = rcond; int max_sweeps = 40;
    void *args[] = {&A, &lda, &qv, &n, &rc, &shift, &status_mask, &max_sweeps, &status, &w, &U, &V, (void *)&ld, &coef, &ctl};
    SGLM_CUDA_OK(cudaLaunchCooperativeKernel((void *)minnorm_jacobi_kernel, dim3(grid), dim3(MN_THREADS), args, 0, st));
    return SGLM_OK;
}
