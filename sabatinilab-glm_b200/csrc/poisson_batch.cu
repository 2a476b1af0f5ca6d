// Batched Poisson GLM grid (BASELINE configs[3]; north-star piece 4): every (fold, alpha) fit of a sweep advances
// together.  Replaces the per-(fold, alpha) TweedieRegressor(power=1).fit calls of the reference
// (backend/sglm.py:112-115, :241; sklearn/linear_model/_glm/glm.py:185-339) — each of which makes two passes over X
// per loss / gradient evaluation — by one Newton-type iteration over the whole batch of B models:
//
//   eta  = X Wt + b          one fp64 GEMM  [T x C] x [C x B]: X is read ONCE per iteration for all B models
//   R    = rw (mu - y)       fused elementwise epilogue + per-model sums (objective, sum mu, sum r), deterministic
//   Gw   = X' R              one fp64 GEMM  [C x T] x [T x B] (split over T, partials added in fixed order)
//   step                     per-model kernels: objective check / step halving, right-hand side H~ w - g, batched
//                            triangular solves with the cached factor of (H~ + alpha n I), intercept, convergence
//
// H~ = X' diag(rw mu_ref) X is the weighted Gram of ONE reference model per fold (tcgen05 digit-plane Gram,
// gram_tc.cu) shared by all alphas of the fold: exact gradient + approximate Hessian = the same fixed point as
// Newton's iteration, reached at a linear rate; the host refreshes H~ when the steps stop contracting fast.
// No host synchronisation inside an iteration (one small status read-back per iteration decides about refreshes).
//
// Rooflines: the two GEMMs are FP64-pipe bound (2 T C B flops each; B200: ~40 TFLOP/s), the epilogue is HBM bound
// (16 T B bytes).
#include <algorithm>

#include "common.cuh"

namespace sglm {

constexpr int PB_BM = 128, PB_BN = 64, PB_BK = 16, PB_THREADS = 256;
constexpr int PB_LDA = PB_BM + 4;      // padded k-major tiles (16-byte aligned rows)
constexpr int PB_LDB = PB_BN + 4;
constexpr size_t PB_GEMM_SMEM = (size_t)2 * PB_BK * (PB_LDA + PB_LDB) * sizeof(double);

// 8 x 4 register tile per thread from k-major shared tiles
__device__ __forceinline__ void pb_tile_fma(const double *As, const double *Bs, int ty, int tx, double (&acc)[8][4]) {
#pragma unroll
    for (int k = 0; k < PB_BK; ++k) {
        double a[8], b[4];
        const double2 *ap = reinterpret_cast<const double2 *>(As + k * PB_LDA + ty * 8);
        const double2 *bp = reinterpret_cast<const double2 *>(Bs + k * PB_LDB + tx * 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const double2 v = ap[i]; a[2 * i] = v.x; a[2 * i + 1] = v.y; }
#pragma unroll
        for (int j = 0; j < 2; ++j) { const double2 v = bp[j]; b[2 * j] = v.x; b[2 * j + 1] = v.y; }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
}

// Eta[M x N] = A[M x K] * B[K x N]   (A = X row-major, rows of the design; B = Wt row-major [C][ldb]; N % 64 == 0)
__global__ void __launch_bounds__(PB_THREADS, 2)
pb_gemm_nn_kernel(const double *__restrict__ A, long long lda, long long M, int K, const double *__restrict__ B,
                  long long ldb, double *__restrict__ Cc, long long ldc) {
    extern __shared__ __align__(16) double pb_smem[];
    double (*As)[PB_BK * PB_LDA] = reinterpret_cast<double (*)[PB_BK * PB_LDA]>(pb_smem);
    double (*Bs)[PB_BK * PB_LDB] = reinterpret_cast<double (*)[PB_BK * PB_LDB]>(pb_smem + 2 * PB_BK * PB_LDA);
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const long long m0 = (long long)blockIdx.x * PB_BM;
    const int n0 = blockIdx.y * PB_BN;
    double acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    // A tile: thread -> (row r = tid / 2, 8 consecutive k from (tid & 1) * 8), stored transposed (k-major)
    const int ar = tid >> 1, ak = (tid & 1) * 8;
    // B tile: thread -> (k = tid / 16, 4 consecutive n from (tid & 15) * 4)
    const int bk = tid >> 4, bn = (tid & 15) * 4;
    const bool a_vec = ((lda & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    auto load = [&](int k0, double (&ra)[8], double (&rb)[4]) {
        const long long row = m0 + ar;
        if (row < M && a_vec && k0 + ak + 8 <= K) {
            const double2 *p = reinterpret_cast<const double2 *>(A + row * lda + k0 + ak);
#pragma unroll
            for (int i = 0; i < 4; ++i) { const double2 v = __ldg(p + i); ra[2 * i] = v.x; ra[2 * i + 1] = v.y; }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) ra[i] = (row < M && k0 + ak + i < K) ? A[row * lda + k0 + ak + i] : 0.0;
        }
        if (k0 + bk < K) {
            const double2 *p = reinterpret_cast<const double2 *>(B + (long long)(k0 + bk) * ldb + n0 + bn);
            const double2 v0 = __ldg(p), v1 = __ldg(p + 1);
            rb[0] = v0.x; rb[1] = v0.y; rb[2] = v1.x; rb[3] = v1.y;
        } else { rb[0] = rb[1] = rb[2] = rb[3] = 0.0; }
    };
    auto store = [&](int buf, const double (&ra)[8], const double (&rb)[4]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) As[buf][(ak + i) * PB_LDA + ar] = ra[i];
        double2 *q = reinterpret_cast<double2 *>(&Bs[buf][bk * PB_LDB + bn]);
        q[0] = make_double2(rb[0], rb[1]); q[1] = make_double2(rb[2], rb[3]);
    };
    double ra[8], rb[4];
    load(0, ra, rb);
    store(0, ra, rb);
    __syncthreads();
    int buf = 0;
    for (int k0 = 0; k0 < K; k0 += PB_BK) {
        const bool more = k0 + PB_BK < K;
        if (more) load(k0 + PB_BK, ra, rb);
        pb_tile_fma(As[buf], Bs[buf], ty, tx, acc);
        if (more) store(buf ^ 1, ra, rb);
        __syncthreads();
        buf ^= 1;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long row = m0 + ty * 8 + i;
        if (row < M) {
            double2 *q = reinterpret_cast<double2 *>(Cc + row * ldc + n0 + tx * 4);
            q[0] = make_double2(acc[i][0], acc[i][1]);
            q[1] = make_double2(acc[i][2], acc[i][3]);
        }
    }
}

// part[split][M x N] = A[k-range, M]' * B[k-range, N]   (A = X [T][lda], M = design columns; B = R [T][ldb])
__global__ void __launch_bounds__(PB_THREADS, 2)
pb_gemm_tn_kernel(const double *__restrict__ A, long long lda, long long Kt, int M, const double *__restrict__ B,
                  long long ldb, int N, long long rows_per_split, double *__restrict__ part, long long ldp) {
    extern __shared__ __align__(16) double pb_smem[];
    double (*As)[PB_BK * PB_LDA] = reinterpret_cast<double (*)[PB_BK * PB_LDA]>(pb_smem);
    double (*Bs)[PB_BK * PB_LDB] = reinterpret_cast<double (*)[PB_BK * PB_LDB]>(pb_smem + 2 * PB_BK * PB_LDA);
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int m0 = blockIdx.x * PB_BM, n0 = blockIdx.y * PB_BN;
    const long long t_lo = (long long)blockIdx.z * rows_per_split, t_hi = min(Kt, t_lo + rows_per_split);
    double acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    const int ak = tid >> 4, am = (tid & 15) * 8;     // A tile [16 t][128 m]: 8 consecutive m per thread
    const int bk = tid >> 4, bn = (tid & 15) * 4;     // B tile [16 t][64 n]
    const bool a_vec = ((lda & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((m0 & 1) == 0);
    auto load = [&](long long t0, double (&ra)[8], double (&rb)[4]) {
        const long long t = t0 + ak;
        if (t < t_hi && a_vec && m0 + am + 8 <= M) {
            const double2 *p = reinterpret_cast<const double2 *>(A + t * lda + m0 + am);
#pragma unroll
            for (int i = 0; i < 4; ++i) { const double2 v = __ldg(p + i); ra[2 * i] = v.x; ra[2 * i + 1] = v.y; }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) ra[i] = (t < t_hi && m0 + am + i < M) ? A[t * lda + m0 + am + i] : 0.0;
        }
        if (t < t_hi) {
            const double2 *p = reinterpret_cast<const double2 *>(B + t * ldb + n0 + bn);
            const double2 v0 = __ldg(p), v1 = __ldg(p + 1);
            rb[0] = v0.x; rb[1] = v0.y; rb[2] = v1.x; rb[3] = v1.y;
        } else { rb[0] = rb[1] = rb[2] = rb[3] = 0.0; }
    };
    auto store = [&](int buf, const double (&ra)[8], const double (&rb)[4]) {
        double2 *qa = reinterpret_cast<double2 *>(&As[buf][ak * PB_LDA + am]);
#pragma unroll
        for (int i = 0; i < 4; ++i) qa[i] = make_double2(ra[2 * i], ra[2 * i + 1]);
        double2 *q = reinterpret_cast<double2 *>(&Bs[buf][bk * PB_LDB + bn]);
        q[0] = make_double2(rb[0], rb[1]); q[1] = make_double2(rb[2], rb[3]);
    };
    double ra[8], rb[4];
    if (t_lo < t_hi) {
        load(t_lo, ra, rb);
        store(0, ra, rb);
        __syncthreads();
        int buf = 0;
        for (long long t0 = t_lo; t0 < t_hi; t0 += PB_BK) {
            const bool more = t0 + PB_BK < t_hi;
            if (more) load(t0 + PB_BK, ra, rb);
            pb_tile_fma(As[buf], Bs[buf], ty, tx, acc);
            if (more) store(buf ^ 1, ra, rb);
            __syncthreads();
            buf ^= 1;
        }
    }
    double *out = part + (long long)blockIdx.z * M * ldp;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = m0 + ty * 8 + i;
        if (row < M) {
            double2 *q = reinterpret_cast<double2 *>(out + (long long)row * ldp + n0 + tx * 4);
            q[0] = make_double2(acc[i][0], acc[i][1]);
            q[1] = make_double2(acc[i][2], acc[i][3]);
        }
    }
}

__global__ void __launch_bounds__(256)
pb_reduce_splits_kernel(const double *__restrict__ part, long long n_elem, int n_split, double *__restrict__ out) {
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < n_elem; e += (long long)gridDim.x * 256) {
        double s = 0.0;
        for (int k = 0; k < n_split; ++k) s += part[(long long)k * n_elem + e];      // fixed order
        out[e] = s;
    }
}

// Elementwise pass over Eta [T][ldb] (thread <-> model column, CTA <-> row range): eta += b, mu = exp(eta).
//   mode 0 (iteration): Eta <- rw (mu - y) in place; partial sums {rw (mu - y eta), rw mu, rw, rw (mu - y)}
//   mode 1 (scores)   : no write; 2 x 8 sums per model for the row weights rw_a / rw_b (train / test):
//                       {n, r^2, y, y^2, y eta, mu, y log y, r} with r = y - mu (as sglm_score_f64)
// Y [T][ldy] holds the response columns (one per distinct `roll`), RW [n_w][ldrw] the row-weight vectors
// (rw id < 0: all rows).  Partial sums per CTA -> pb_finish_sums_kernel adds them in fixed order.
constexpr int PB_NS_IT = 4, PB_NS_SC = 16;
template <int MODE>
__global__ void __launch_bounds__(128)
pb_epilogue_kernel(double *__restrict__ Eta, long long ldb, int n_models, long long T, const double *__restrict__ Y,
                   long long ldy, const int *__restrict__ ycol, const double *__restrict__ RW, long long ldrw,
                   const int *__restrict__ rw_a, const int *__restrict__ rw_b, const double *__restrict__ bvec,
                   const int *__restrict__ status, double *__restrict__ partials) {
    constexpr int NS = MODE == 0 ? PB_NS_IT : PB_NS_SC;
    const int mdl = blockIdx.y * 128 + threadIdx.x;
    const bool live = mdl < n_models;
    const long long rows_per = (T + gridDim.x - 1) / gridDim.x;
    const long long t0 = (long long)blockIdx.x * rows_per, t1 = min(T, t0 + rows_per);
    double acc[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) acc[i] = 0.0;
    if (live && !(MODE == 0 && status && status[mdl] != 0)) {
        const double b = bvec[mdl];
        const double *y = Y + ycol[mdl];
        const double *wa = rw_a[mdl] >= 0 ? RW + (long long)rw_a[mdl] * ldrw : nullptr;
        const double *wb = (MODE == 1 && rw_b[mdl] >= 0) ? RW + (long long)rw_b[mdl] * ldrw : nullptr;
        for (long long t = t0; t < t1; ++t) {
            const double eta = Eta[t * ldb + mdl] + b;
            const double yt = y[t * ldy];
            if (MODE == 0) {
                const double m = wa ? wa[t] : 1.0;
                double r = 0.0;
                if (m != 0.0) {
                    const double mu = exp(eta);
                    r = m * (mu - yt);
                    acc[0] += m * (mu - yt * eta);
                    acc[1] += m * mu;
                    acc[2] += m;
                    acc[3] += r;
                }
                Eta[t * ldb + mdl] = r;
            } else {
                const double ma = wa ? wa[t] : 1.0, mb = (MODE == 1 && rw_b[mdl] < -1) ? 0.0 : (wb ? wb[t] : 1.0);
                if (ma != 0.0 || mb != 0.0) {
                    const double mu = exp(eta), r = yt - mu;
                    const double ylogy = (yt > 0.0) ? yt * log(yt) : 0.0;
                    const double v[8] = {1.0, r * r, yt, yt * yt, yt * eta, mu, ylogy, r};
#pragma unroll
                    for (int i = 0; i < 8; ++i) { acc[i] += ma * v[i]; acc[8 + i] += mb * v[i]; }
                }
            }
        }
    }
    if (live) {
        double *dst = partials + ((long long)blockIdx.x * n_models + mdl) * NS;
#pragma unroll
        for (int i = 0; i < NS; ++i) dst[i] = acc[i];
    }
}

__global__ void __launch_bounds__(256)
pb_finish_sums_kernel(const double *__restrict__ partials, int n_part, long long n_vals, double *__restrict__ sums) {
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < n_vals; e += (long long)gridDim.x * 256) {
        double s = 0.0;
        for (int p = 0; p < n_part; ++p) s += partials[(long long)p * n_vals + e];
        sums[e] = s;
    }
}

// ---- per-model step logic (one CTA per model) -----------------------------------------------------------------
struct PbState {
    double *W, *Wprev, *Wnew, *rhs;          // [B][ldw]
    double *b, *bprev, *fprev, *fcur, *last_step, *step_out, *ratio_out;   // [B]
    int *halv, *n_iter, *status, *flag, *has_prev;                          // [B]
    const double *alpha, *n_tot, *tol;       // [B]
    const int *fit_icpt, *max_iter, *hess_id;
    const double *const *HQ;                 // [n_hess] centred weighted Gram Qc (C x ldq)
    const double *const *Hxbar;              // [n_hess] weighted column means
    const double *Hh11;                      // [n_hess] sum of the weights
    const double *sums;                      // [B][4] from the epilogue
    const double *Gw;                        // [C][ldb] gradient X'R
    const double *hscale;                    // [B] the model's Hessian = hscale * (the group's H~): a fold's first step
                                             // borrows the full-data Hessian scaled by its share of sum(y)
    long long ldw, ldq, ldb;
    int C, B;
};

__device__ __forceinline__ double pb_block_sum(double v, double *sh) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += sh[w];
    return s;
}
__device__ __forceinline__ double pb_block_max(double v, double *sh) {
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = sh[0];
    for (int w = 1; w < 8; ++w) s = fmax(s, sh[w]);
    return s;
}

// objective check / step halving, else the right-hand side of the chord step  (H~ + a n I) w_new = H~ w - g
__global__ void __launch_bounds__(256)
pb_rhs_kernel(PbState s) {
    __shared__ double sh[8];
    extern __shared__ __align__(16) double wsh[];          // [C] current w
    const int m = blockIdx.x, tid = threadIdx.x;
    if (s.status[m] != 0) { if (tid == 0) s.flag[m] = 0; return; }
    double *w = s.W + (long long)m * s.ldw;
    const double *wp = s.Wprev + (long long)m * s.ldw;
    double ww = 0.0;
    for (int j = tid; j < s.C; j += 256) { const double v = w[j]; wsh[j] = v; ww = fma(v, v, ww); }
    ww = pb_block_sum(ww, sh);
    const double f = s.sums[4 * m + 0] / s.n_tot[m] + 0.5 * s.alpha[m] * ww;
    const double fp = s.fprev[m];
    const bool worse = s.has_prev[m] && !(f <= fp + 1e-12 * fmax(1.0, fabs(fp)));
    if (worse && s.halv[m] < 30) {
        for (int j = tid; j < s.C; j += 256) w[j] = 0.5 * (wsh[j] + wp[j]);
        if (tid == 0) { s.b[m] = 0.5 * (s.b[m] + s.bprev[m]); s.halv[m] += 1; s.flag[m] = 2; }
        return;
    }
    if (s.n_iter[m] >= s.max_iter[m]) { if (tid == 0) { s.status[m] = 2; s.flag[m] = 0; } return; }
    const int h = s.hess_id[m];
    const double *Q = s.HQ[h];
    const double *xbar = s.Hxbar[h];
    const double g_b = s.sums[4 * m + 3];
    const bool icpt = s.fit_icpt[m] != 0;
    const double inv_hs = 1.0 / s.hscale[m];
    double *rhs = s.rhs + (long long)m * s.ldw;
    const int warp = tid >> 5, lane = tid & 31;
    // (hs Q + a n I) w_new = hs Q w - g + xbar g_b   <=>   (Q + (a n / hs) I) w_new = Q w + (-g + xbar g_b) / hs ; the cached
    // factor belongs to Q + a n' I with a n' ~ a n / hs (exact for hs = 1)
    for (int j = warp; j < s.C; j += 8) {
        const double *row = Q + (long long)j * s.ldq;
        double d = 0.0;
        for (int k = lane; k < s.C; k += 32) d = fma(row[k], wsh[k], d);
        d = warp_sum(d);
        if (lane == 0) rhs[j] = d + (-s.Gw[(long long)j * s.ldb + m] + (icpt ? xbar[j] * g_b : 0.0)) * inv_hs;
    }
    if (tid == 0) { s.halv[m] = 0; s.fcur[m] = f; s.flag[m] = 1; }
}

// intercept, step size, convergence
__global__ void __launch_bounds__(256)
pb_finish_kernel(PbState s) {
    __shared__ double sh[8];
    const int m = blockIdx.x, tid = threadIdx.x;
    if (s.flag[m] != 1) { if (tid == 0) s.step_out[m] = s.flag[m] == 2 ? -1.0 : 0.0; return; }
    double *w = s.W + (long long)m * s.ldw, *wp = s.Wprev + (long long)m * s.ldw;
    const double *wn = s.Wnew + (long long)m * s.ldw;
    const int h = s.hess_id[m];
    const double *xbar = s.Hxbar[h];
    const bool icpt = s.fit_icpt[m] != 0;
    double dot = 0.0, dw = 0.0, wmax = 0.0;
    for (int j = tid; j < s.C; j += 256) {
        const double a = wn[j], o = w[j];
        dot = fma(icpt ? xbar[j] : 0.0, a - o, dot);
        dw = fmax(dw, fabs(a - o));
        wmax = fmax(wmax, fabs(a));
    }
    dot = pb_block_sum(dot, sh);
    dw = pb_block_max(dw, sh);
    wmax = pb_block_max(wmax, sh);
    const double b_old = s.b[m];
    const double b_new = icpt ? b_old - s.sums[4 * m + 3] / (s.Hh11[h] * s.hscale[m]) - dot : 0.0;
    for (int j = tid; j < s.C; j += 256) { wp[j] = w[j]; w[j] = wn[j]; }
    if (tid == 0) {
        const double step = fmax(dw, fabs(b_new - b_old)) / fmax(1.0, wmax);
        const double last = s.last_step[m];
        const bool have = last > 0.0 && last < 1e300;
        const double ratio = have ? step / last : 1.0;
        // contraction estimate = the worse of the last two ratios (1.0 = not known yet: the first two steps after a
        // Hessian refresh cannot declare convergence unless the step itself is below the target)
        const double rho = fmin(fmax(ratio, s.ratio_out[m]), 0.9);
        s.bprev[m] = b_old; s.b[m] = b_new; s.fprev[m] = s.fcur[m]; s.has_prev[m] = 1;
        s.n_iter[m] += 1;
        // linear convergence (inexact Hessian): after a step that contracted by rho the iterate is within
        // step * rho / (1 - rho) of the optimum
        if (step * rho / (1.0 - rho) <= fmin(s.tol[m], 1e-8) || step <= 1e-12) s.status[m] = 1;
        s.step_out[m] = step;
        s.ratio_out[m] = ratio;
        s.last_step[m] = step;
    }
}

}  // namespace sglm

using namespace sglm;

extern "C" size_t sglm_pb_gemm_tn_workspace_bytes(int64_t T, int32_t C, int64_t ldb) {
    const long long rows_per = 8192;
    const long long n_split = std::max<long long>(1, ceil_div<long long>(std::max<long long>(T, 1), rows_per));
    return (size_t)n_split * (size_t)C * (size_t)ldb * sizeof(double);
}

// Eta [T][ldb] = X [T][C] * Wt [C][ldb]   (ldb a multiple of 64; the columns beyond the models are zero in Wt)
extern "C" int sglm_pb_eta_f64(const double *X, int64_t ldx, int64_t T, int32_t C, const double *Wt, int64_t ldb,
                               double *Eta, void *stream) {
    SGLM_CHECK_ARG(T >= 0 && C > 0 && ldx >= C && ldb >= 64 && ldb % 64 == 0, SGLM_E_SHAPE, "pb_eta: bad shape");
    if (T == 0) return SGLM_OK;
    SGLM_CHECK_ARG(X && Wt && Eta, SGLM_E_INVALID_ARG, "pb_eta: null pointer");
    SGLM_CHECK_ARG(((uintptr_t)Wt & 15) == 0 && ((uintptr_t)Eta & 15) == 0, SGLM_E_ALIGN, "pb_eta: 16-byte alignment");
    dim3 grid((unsigned)ceil_div<long long>(T, PB_BM), (unsigned)(ldb / PB_BN));
    SGLM_CUDA_OK(cudaFuncSetAttribute(pb_gemm_nn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PB_GEMM_SMEM));
    pb_gemm_nn_kernel<<<grid, PB_THREADS, PB_GEMM_SMEM, (cudaStream_t)stream>>>(X, ldx, T, C, Wt, ldb, Eta, ldb);
    SGLM_LAUNCH_OK("pb_gemm_nn_kernel");
    return SGLM_OK;
}

// ------------------------------------------------------------------ quadratic forms of many vectors: v_m' A v_m
// out[m] = sum_i Vt[i][m] (A Vt)[i][m]: one fp64 GEMM (A is read once per 64 vectors instead of once per 8 as in
// quadform_kernel) followed by column dots in fixed row chunks (deterministic).  Scores of a CV grid from the
// statistics: RSS(model, row set) = v' G[set] v (backend/sglm.py:305-312 computes them by three prediction passes).
constexpr int QG_PARTS = 16;
__global__ void __launch_bounds__(128)
coldot_kernel(const double *__restrict__ Vt, const double *__restrict__ Y, long long n, long long ldb, int n_models,
              double *__restrict__ part) {
    const int m = blockIdx.x * 128 + threadIdx.x;
    if (m >= n_models) return;
    const long long per = (n + QG_PARTS - 1) / QG_PARTS;
    const long long i0 = (long long)blockIdx.y * per, i1 = min(n, i0 + per);
    double acc = 0.0;
    for (long long i = i0; i < i1; ++i) acc = fma(Vt[i * ldb + m], Y[i * ldb + m], acc);
    part[(long long)blockIdx.y * n_models + m] = acc;
}
__global__ void __launch_bounds__(128)
coldot_finish_kernel(const double *__restrict__ part, int n_models, double *__restrict__ out) {
    const int m = blockIdx.x * 128 + threadIdx.x;
    if (m >= n_models) return;
    double s = 0.0;
    for (int p = 0; p < QG_PARTS; ++p) s += part[(long long)p * n_models + m];
    out[m] = s;
}

// Vt [n][ldb]: the vectors as COLUMNS (ldb % 64 == 0, columns >= n_models zero); work: 2 * n * ldb doubles is not needed —
// Y [n][ldb] and part [16][n_models] are caller buffers.
extern "C" int sglm_quadform_gemm_f64(const double *A, int64_t lda, int32_t n, const double *Vt, int64_t ldb,
                                      int32_t n_models, double *Y, double *part, double *out, void *stream) {
    SGLM_CHECK_ARG(n > 0 && lda >= n && n_models > 0 && ldb >= n_models && ldb % 64 == 0, SGLM_E_SHAPE, "quadform_gemm: bad shape");
    SGLM_CHECK_ARG(A && Vt && Y && part && out, SGLM_E_INVALID_ARG, "quadform_gemm: null pointer");
    SGLM_CHECK_ARG(((uintptr_t)Vt & 15) == 0 && ((uintptr_t)Y & 15) == 0, SGLM_E_ALIGN, "quadform_gemm: 16-byte alignment");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)ceil_div<long long>(n, PB_BM), (unsigned)(ldb / PB_BN));
    SGLM_CUDA_OK(cudaFuncSetAttribute(pb_gemm_nn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PB_GEMM_SMEM));
    pb_gemm_nn_kernel<<<grid, PB_THREADS, PB_GEMM_SMEM, st>>>(A, lda, n, n, Vt, ldb, Y, ldb);
    SGLM_LAUNCH_OK("pb_gemm_nn_kernel(quadform)");
    dim3 dgrid((unsigned)ceil_div(n_models, 128), QG_PARTS);
    coldot_kernel<<<dgrid, 128, 0, st>>>(Vt, Y, n, ldb, n_models, part);
    SGLM_LAUNCH_OK("coldot_kernel");
    coldot_finish_kernel<<<ceil_div(n_models, 128), 128, 0, st>>>(part, n_models, out);
    SGLM_LAUNCH_OK("coldot_finish_kernel");
    return SGLM_OK;
}

// Gw [C][ldb] = X' R   (R [T][ldb]); deterministic: fixed T chunks of 8192 rows, partials added in order
extern "C" int sglm_pb_xt_r_f64(const double *X, int64_t ldx, int64_t T, int32_t C, const double *R, int64_t ldb,
                                double *Gw, void *workspace, size_t workspace_bytes, void *stream) {
    SGLM_CHECK_ARG(T >= 0 && C > 0 && ldx >= C && ldb >= 64 && ldb % 64 == 0, SGLM_E_SHAPE, "pb_xt_r: bad shape");
    SGLM_CHECK_ARG(X && R && Gw && workspace, SGLM_E_INVALID_ARG, "pb_xt_r: null pointer");
    SGLM_CHECK_ARG(workspace_bytes >= sglm_pb_gemm_tn_workspace_bytes(T, C, ldb), SGLM_E_WORKSPACE, "pb_xt_r: workspace too small");
    const long long rows_per = 8192;
    const int n_split = (int)std::max<long long>(1, ceil_div<long long>(std::max<long long>(T, 1), rows_per));
    dim3 grid((unsigned)ceil_div(C, PB_BM), (unsigned)(ldb / PB_BN), (unsigned)n_split);
    SGLM_CUDA_OK(cudaFuncSetAttribute(pb_gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PB_GEMM_SMEM));
    pb_gemm_tn_kernel<<<grid, PB_THREADS, PB_GEMM_SMEM, (cudaStream_t)stream>>>(X, ldx, T, C, R, ldb, (int)ldb, rows_per,
                                                                   (double *)workspace, ldb);
    SGLM_LAUNCH_OK("pb_gemm_tn_kernel");
    const long long n_elem = (long long)C * ldb;
    pb_reduce_splits_kernel<<<(unsigned)std::min<long long>(ceil_div<long long>(n_elem, 256), sm_count() * 8LL), 256, 0,
                              (cudaStream_t)stream>>>((const double *)workspace, n_elem, n_split, Gw);
    SGLM_LAUNCH_OK("pb_reduce_splits_kernel");
    return SGLM_OK;
}

static int pb_epi_grid(long long T) { return (int)std::max<long long>(1, std::min<long long>(sm_count() * 8LL, ceil_div<long long>(T, 64))); }

extern "C" size_t sglm_pb_epilogue_workspace_bytes(int64_t T, int32_t n_models) {
    return (size_t)pb_epi_grid(T) * (size_t)std::max(n_models, 1) * PB_NS_SC * sizeof(double);
}

// mode 0: Eta <- rw (mu - y) in place, sums [B][4]; mode 1: scores, sums [B][16]  (see pb_epilogue_kernel)
extern "C" int sglm_pb_epilogue_f64(double *Eta, int64_t ldb, int32_t n_models, int64_t T, const double *Y,
                                    int64_t ldy, const int32_t *ycol, const double *RW, int64_t ldrw,
                                    const int32_t *rw_a, const int32_t *rw_b, const double *b,
                                    const int32_t *status, int32_t mode, double *sums, void *workspace,
                                    size_t workspace_bytes, void *stream) {
    SGLM_CHECK_ARG(T >= 0 && n_models > 0 && ldb >= n_models && ldy >= 1, SGLM_E_SHAPE, "pb_epilogue: bad shape");
    SGLM_CHECK_ARG(Eta && Y && ycol && rw_a && b && sums && workspace && (mode == 0 || rw_b), SGLM_E_INVALID_ARG,
                   "pb_epilogue: null pointer");
    SGLM_CHECK_ARG(workspace_bytes >= sglm_pb_epilogue_workspace_bytes(T, n_models), SGLM_E_WORKSPACE,
                   "pb_epilogue: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int gx = pb_epi_grid(T);
    dim3 grid((unsigned)gx, (unsigned)ceil_div(n_models, 128));
    const int NS = mode == 0 ? PB_NS_IT : PB_NS_SC;
    if (mode == 0)
        pb_epilogue_kernel<0><<<grid, 128, 0, st>>>(Eta, ldb, n_models, T, Y, ldy, ycol, RW, ldrw, rw_a, rw_b, b, status,
                                                    (double *)workspace);
    else
        pb_epilogue_kernel<1><<<grid, 128, 0, st>>>(Eta, ldb, n_models, T, Y, ldy, ycol, RW, ldrw, rw_a, rw_b, b, status,
                                                    (double *)workspace);
    SGLM_LAUNCH_OK("pb_epilogue_kernel");
    const long long n_vals = (long long)n_models * NS;
    pb_finish_sums_kernel<<<(unsigned)ceil_div<long long>(n_vals, 256), 256, 0, st>>>((const double *)workspace, gx, n_vals, sums);
    SGLM_LAUNCH_OK("pb_finish_sums_kernel");
    return SGLM_OK;
}

// One chord step of every active model: objective check / halving + right-hand sides, batched triangular solves with
// the cached factors L_of[m] of (H~ + alpha n I), intercepts + convergence.  `state_host` = the PbState fields as a
// host array of 64-bit words in declaration order (pointers, then ldw, ldq, ldb, then C | B << 32).
extern "C" int sglm_pb_step_f64(const uint64_t *state_host, const double *const *L_of, void *stream);

extern "C" int sglm_pb_state_words(void) { return (int)(sizeof(PbState) / 8); }

extern "C" int sglm_chol_solve_batched_f64(const double *const *L_of, int64_t ldq, int32_t C, const double *rhs,
                                           double *out, int64_t ld, const int32_t *flags, int32_t n_systems,
                                           void *stream);

extern "C" int sglm_pb_step_f64(const uint64_t *state_host, const double *const *L_of, void *stream) {
    SGLM_CHECK_ARG(state_host && L_of, SGLM_E_INVALID_ARG, "pb_step: null pointer");
    static_assert(sizeof(PbState) % 8 == 0, "PbState is passed as 64-bit words");
    PbState s;
    memcpy(&s, state_host, sizeof(PbState));
    SGLM_CHECK_ARG(s.C > 0 && s.B > 0 && s.ldw >= s.C && s.ldq >= s.C && s.ldb >= s.B, SGLM_E_SHAPE, "pb_step: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)s.C * sizeof(double);
    SGLM_CHECK_ARG(smem <= 200 * 1024, SGLM_E_UNSUPPORTED, "pb_step: C=%d too large", s.C);
    SGLM_CUDA_OK(cudaFuncSetAttribute(pb_rhs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pb_rhs_kernel<<<s.B, 256, smem, st>>>(s);
    SGLM_LAUNCH_OK("pb_rhs_kernel");
    const int rc = sglm_chol_solve_batched_f64(L_of, s.ldq, s.C, s.rhs, s.Wnew, s.ldw, s.flag, s.B, stream);
    if (rc != SGLM_OK) return rc;
    pb_finish_kernel<<<s.B, 256, 0, st>>>(s);
    SGLM_LAUNCH_OK("pb_finish_kernel");
    return SGLM_OK;
}
