// Batched solvers on sufficient statistics: centring, Gram coordinate descent
// (ElasticNet / Lasso), Cholesky (Ridge / OLS), intercepts and scores-from-statistics.
//
// These replace the per-(fold, alpha, l1_ratio) scikit-learn fits that the reference
// launches one at a time from Python threads (backend/sglm_cv.py:106-170 -> backend/
// sglm.py:241 -> sklearn cd_fast / _ridge): every model of the CV grid becomes one CTA
// of a single launch, many models resident per SM.
//
// Roofline of the CD kernel: memory (rows of Q streamed from L2/HBM).  Algorithmic bytes
// per model = 8*C per coordinate update that changes w (one row of Q) — the kernel counts
// those updates (info[6m+3]) so bench.py can report sum(bytes)/time.
#include <algorithm>

#include "common.cuh"

namespace sglm {

// --------------------------------------------------------------------------- centring
__global__ void __launch_bounds__(256)
center_stats_kernel(const double *__restrict__ Ap, const double *__restrict__ Am, long long ldg, int C,
                    int n_y, int y_col, int fit_intercept, double *__restrict__ Qc, long long ldq,
                    double *__restrict__ qc, double *__restrict__ xbar, double *__restrict__ diag,
                    double *__restrict__ scal) {
    const int one = C + n_y, yc = C + y_col;
    auto A = [&](int i, int j) {
        double v = Ap[(long long)i * ldg + j];
        if (Am) v -= Am[(long long)i * ldg + j];
        return v;
    };
    const double n = A(one, one);
    const double sy = A(yc, one);
    const double ybar = fit_intercept ? sy / n : 0.0;
    const int i = blockIdx.y;
    const double xi = fit_intercept ? A(i, one) / n : 0.0;
    for (int j = blockIdx.x * 256 + threadIdx.x; j < ldq; j += gridDim.x * 256) {
        double v = 0.0;
        if (j < C) {
            const double xj = fit_intercept ? A(j, one) / n : 0.0;
            v = A(i, j) - n * xi * xj;
        }
        Qc[(long long)i * ldq + j] = v;
        if (j == i) diag[i] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        qc[i] = A(i, yc) - n * xi * ybar;
        xbar[i] = xi;
        if (i == 0) {
            scal[0] = A(yc, yc) - n * ybar * ybar;
            scal[1] = n;
            scal[2] = ybar;
            scal[3] = sy;
        }
    }
}

// --------------------------------------------------------------------------- Gram CD
template <int NW>
__device__ __forceinline__ void cta_sync() {
    if (NW == 1) __syncwarp(); else __syncthreads();
}

// Reduce K partial values over the CTA.  Entries [0, K_SUM) are summed, the rest maxed.
template <int NW, int K, int K_SUM>
__device__ __forceinline__ void cta_reduce(double (&v)[K], double *scratch) {
#pragma unroll
    for (int i = 0; i < K; ++i) v[i] = (i < K_SUM) ? warp_sum(v[i]) : warp_max(v[i]);
    if (NW > 1) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < K; ++i) scratch[warp * K + i] = v[i];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < K; ++i) {
            double a = scratch[i];
            for (int w = 1; w < NW; ++w) a = (i < K_SUM) ? a + scratch[w * K + i] : fmax(a, scratch[w * K + i]);
            v[i] = a;
        }
        __syncthreads();
    }
}

constexpr int CD_CH = 4;   // 16-byte chunks of Qw owned by one thread per pass of the panel update

// One CTA (NW warps) per model.  Cyclic coordinate descent is sequential — coordinate j+1
// needs the Qw produced by coordinate j — so a one-coordinate-at-a-time kernel is bound by
// the latency of its own instruction chain (measured: ~2400 cycles per update, ncu
// profiles/r1_v2_cd_one_*).  The sweep is therefore blocked 32 coordinates at a time:
//
//   phase 1 (warp 0, registers + shuffles only): lane l owns coordinate j_l of the block,
//     its Qw[j_l] and the column Q[j_0..j_31][j_l] of the 32x32 diagonal sub-block; the 32
//     coordinates are visited in order, each delta is broadcast with a shuffle and folded
//     into the other lanes' Qw[j_l] — the exact FMA sequence of the unblocked algorithm;
//   phase 2 (all warps): the 32 deltas are applied to all of Qw at once,
//     Qw[k] = fma(delta_i, Q[j_i][k], Qw[k]) for i in block order (again the same FMA
//     sequence per element), streaming only the rows whose coefficient moved, many
//     independent 16-byte loads in flight per thread — this is the HBM/L2-bound part.
//
// The iterates are bit-identical to the one-coordinate-at-a-time formulation; only the
// schedule changes.  Shared memory per model: w, Qw, the active list (no copy of Q).
template <int NB, int CD_RG, int MINB>
__global__ void __launch_bounds__((NB + 1) * 32, MINB)
enet_cd_gram_kernel(const double *const *__restrict__ prob_Q, const double *const *__restrict__ prob_q,
                    const double *const *__restrict__ prob_diag, const double *__restrict__ prob_yy,
                    long long ldq, int C, const int *__restrict__ prob_of_model,
                    const double *__restrict__ l1_reg, const double *__restrict__ l2_reg,
                    const double *__restrict__ tol_in, const int *__restrict__ max_iter_in, int warm_start,
                    int do_screening, double *__restrict__ W, long long ldw, double *__restrict__ info) {
    constexpr int NW = NB + 1;              // warp 0: sequential phase; warps 1..NB: panel update
    constexpr int NT = NW * 32;
    constexpr int NTB = NB * 32;
    extern __shared__ __align__(16) double sm[];
    const int Cp = (C + 1) & ~1;
    double *w = sm;                         // [Cp]
    double *Qw = w + Cp;                    // [Cp]  (element C, when C is odd, is a harmless pad)
    double *scratch = Qw + Cp;              // [NW * 8]
    double *dlt = scratch + NW * 8;         // [32] deltas of the rows to stream (compacted)
    int *active = reinterpret_cast<int *>(dlt + 32);                           // [C]
    int *jb = active + C;                                                      // [32] their coordinates
    int *iscr = jb + 32;                                                       // [NW + 2]
    unsigned char *state = reinterpret_cast<unsigned char *>(iscr + NW + 2);   // [C] 2 = to-drop

    const int m = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pid = prob_of_model[m];
    const double *__restrict__ Q = prob_Q[pid];
    const double *__restrict__ q = prob_q[pid];
    const double *__restrict__ dg = prob_diag[pid];
    const double yy = prob_yy[pid];
    const double l1 = l1_reg[m], l2 = l2_reg[m];
    const double d_w_tol = tol_in[m];
    const double tol = d_w_tol * yy;
    const int max_iter = max_iter_in[m];
    const bool vec2 = ((ldq & 1) == 0) && ((reinterpret_cast<uintptr_t>(Q) & 15) == 0) && (ldq >= Cp);
    const int Cp2 = Cp >> 1;
    long long n_upd = 0, n_blk = 0;

    // Qw += a * Q[j, :]   (single-row path: warm start, screening drops)
    auto axpy_row = [&](int j, double a) {
        const double *row = Q + (long long)j * ldq;
#pragma unroll 8
        for (int k = tid; k < C; k += NT) Qw[k] += a * __ldg(row + k);
    };

    for (int j = tid; j < Cp; j += NT) {
        w[j] = (warm_start && j < C) ? W[(long long)m * ldw + j] : 0.0;
        Qw[j] = 0.0;
        if (j < C) state[j] = 0;
    }
    cta_sync<NW>();
    if (warm_start) {
        for (int j = 0; j < C; ++j) {
            const double wj = w[j];
            if (wj != 0.0) { axpy_row(j, wj); ++n_upd; }
        }
        cta_sync<NW>();
    }

    // duality gap (sklearn _cd_fast.pyx:1006-1092 gap_enet_gram); uniform result in all threads
    double dual_norm = 0.0;
    auto compute_gap = [&]() -> double {
        double v[6] = {0, 0, 0, 0, 0, 0};   // ww, wq, wQw, |w|_1, sum xta^2, max |xta|
        for (int j = tid; j < C; j += NT) {
            const double wj = w[j], Qwj = Qw[j], qj = __ldg(q + j);
            v[0] += wj * wj;
            v[1] += wj * qj;
            v[2] += wj * Qwj;
            v[3] += fabs(wj);
            const double xta = (l1 == 0.0) ? (qj - Qwj) : (qj - Qwj - l2 * wj);
            v[4] += xta * xta;
            v[5] = fmax(v[5], fabs(xta));
        }
        cta_reduce<NW, 6, 5>(v, scratch);
        const double R2 = yy + v[2] - 2.0 * v[1];
        const double Ry = yy - v[1];
        const double w22 = (l2 > 0.0) ? v[0] : 0.0;
        if (l1 == 0.0) {
            dual_norm = v[4];
            if (l2 == 0.0) return v[4];
            return R2 + 0.5 * l2 * w22 - Ry + 1.0 / (2.0 * l2) * v[4];
        }
        dual_norm = v[5];
        const double primal = 0.5 * (R2 + l2 * w22) + l1 * v[3];
        const double scale = (dual_norm > l1) ? l1 / dual_norm : 1.0;
        const double dual = -0.5 * scale * scale * (R2 + l2 * w22) + scale * Ry;
        return primal - dual;
    };

    const bool screening = do_screening && (l1 != 0.0);
    int n_active = C;
    for (int j = tid; j < C; j += NT) active[j] = j;

    // gap-safe screening (sklearn _cd_fast.pyx:1187-1208, :1259-1279).  Ordered compaction
    // of the surviving coordinates; dropped non-zero coordinates are applied afterwards in
    // ascending order (XtA is evaluated from the pre-drop Qw, as in sklearn).
    auto screen = [&](double gap, bool initial) {
        const double radius = sqrt(2.0 * fabs(gap)) / l1;
        const double denom = fmax(l1, dual_norm);
        const int n_cand = initial ? C : n_active;
        int n_new = 0, any_drop = 0;
        for (int base = 0; base < n_cand; base += NT) {
            const int idx = base + tid;
            int j = -1, keep = 0;
            if (idx < n_cand) {
                j = initial ? idx : active[idx];
                const double djj = __ldg(dg + j);
                if (initial && djj == 0.0) {
                    w[j] = 0.0;
                } else {
                    const double xta = __ldg(q + j) - Qw[j] - l2 * w[j];
                    const double d_j = (1.0 - fabs(xta / denom)) / sqrt(djj + l2);
                    if (d_j <= radius) keep = 1;
                    else if (w[j] != 0.0) { state[j] = 2; any_drop = 1; }
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            const int pre = __popc(bal & ((1u << lane) - 1));
            int woff = 0, tot = __popc(bal);
            if (NW > 1) {
                if (lane == 0) iscr[warp] = tot;
                __syncthreads();
                tot = 0;
                for (int ww = 0; ww < NW; ++ww) {
                    const int c = iscr[ww];
                    if (ww < warp) woff += c;
                    tot += c;
                }
            } else {
                __syncwarp();
            }
            if (keep) active[n_new + woff + pre] = j;
            n_new += tot;
            cta_sync<NW>();
        }
        n_active = n_new;
        any_drop = __syncthreads_or(any_drop);
        if (any_drop) {
            for (int j = 0; j < C; ++j) {
                if (state[j] == 2) {          // uniform: state is only written before the barrier above
                    const double wj = w[j];
                    cta_sync<NW>();
                    axpy_row(j, -wj);
                    if (tid == 0) { w[j] = 0.0; state[j] = 0; }
                    ++n_upd;
                    cta_sync<NW>();
                }
            }
        }
    };

    double gap = compute_gap();
    int n_iter = 0;
    bool done = (gap >= 0.0 && gap <= tol) || max_iter <= 0;
    if (!done) {
        if (screening) screen(gap, true);
        cta_sync<NW>();
        for (n_iter = 0; n_iter < max_iter; ++n_iter) {
            double wmax_l = 0.0, dwmax_l = 0.0;            // per-lane maxima (warp 0)
            const int n_blocks = (n_active + 31) >> 5;
            // block registers of warp 0: coordinate, q, diag, 1/(diag+l2), column of the diagonal sub-block
            int j_l = 0;
            bool ok_l = false, valid_l = false;
            double q_l = 0.0, d_l = 0.0, inv_l = 0.0, den_l = 1.0;
            double col[32];
            auto load_block = [&](int b) {
                const int pos = (b << 5) + lane;
                valid_l = pos < n_active;
                j_l = valid_l ? active[pos] : 0;
                q_l = valid_l ? __ldg(q + j_l) : 0.0;
                d_l = valid_l ? __ldg(dg + j_l) : 0.0;
                ok_l = valid_l && d_l != 0.0;
                den_l = ok_l ? d_l + l2 : 1.0;
                inv_l = 1.0 / den_l;
                const int nb = min(32, n_active - (b << 5));
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int ji = (i < nb) ? active[(b << 5) + i] : 0;          // smem broadcast
                    col[i] = (i < nb && valid_l) ? __ldg(Q + (long long)ji * ldq + j_l) : 0.0;
                }
            };
            if (warp == 0 && n_blocks > 0) load_block(0);

            for (int b = 0; b < n_blocks; ++b) {
                if (warp == 0) {
                    // ---------------- phase 1: the 32 coordinates of the block, in order
                    double Qw_l = valid_l ? Qw[j_l] : 0.0;
                    const double w_l = valid_l ? w[j_l] : 0.0;
                    const double wd_l = w_l * d_l;
                    double w_new_l = w_l, delta_l = 0.0;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const double tmp = (q_l - Qw_l) + wd_l;
                        const double soft = copysign(fmax(fabs(tmp) - l1, 0.0), tmp);
                        double cand = soft * inv_l;
                        cand = fma(fma(-den_l, cand, soft), inv_l, cand);     // one Newton step: soft / den
                        const double dc = ok_l ? cand - w_l : 0.0;
                        const double di = __shfl_sync(0xffffffffu, dc, i);
                        if (lane == i) { w_new_l = ok_l ? cand : w_l; delta_l = dc; }
                        Qw_l = fma(di, col[i], Qw_l);
                    }
                    const unsigned nz = __ballot_sync(0xffffffffu, delta_l != 0.0);
                    if (ok_l) {
                        dwmax_l = fmax(dwmax_l, fabs(delta_l));
                        wmax_l = fmax(wmax_l, fabs(w_new_l));
                    }
                    if (delta_l != 0.0) {
                        w[j_l] = w_new_l;
                        const int slot = __popc(nz & ((1u << lane) - 1));          // compacted, block order
                        dlt[slot] = delta_l;
                        jb[slot] = j_l;
                    }
                    if (lane == 0) iscr[NW + 1] = __popc(nz);
                    n_upd += __popc(nz);
                    ++n_blk;
                    if (b + 1 < n_blocks) load_block(b + 1);      // in flight during phase 2
                }
                cta_sync<NW>();
                // ---------------- phase 2: Qw += sum_i delta_i Q[j_i, :]   (rows in block order)
                const int n_rows = iscr[NW + 1];
                if (n_rows > 0 && warp > 0) {
                    const int bt = tid - 32;
                    if (vec2) {
                        const double2 *Q2 = reinterpret_cast<const double2 *>(Q);
                        double2 *Qw2 = reinterpret_cast<double2 *>(Qw);
                        const long long ld2 = ldq >> 1;
                        for (int c0 = bt; c0 < Cp2; c0 += NTB * CD_CH) {
                            double2 acc[CD_CH];
#pragma unroll
                            for (int u = 0; u < CD_CH; ++u) {
                                const int c = c0 + u * NTB;
                                acc[u] = (c < Cp2) ? Qw2[c] : make_double2(0.0, 0.0);
                            }
                            for (int r0 = 0; r0 < n_rows; r0 += CD_RG) {
                                double2 v[CD_RG][CD_CH];
                                double dl[CD_RG];
#pragma unroll
                                for (int rr = 0; rr < CD_RG; ++rr) {
                                    const bool rok = r0 + rr < n_rows;
                                    dl[rr] = rok ? dlt[r0 + rr] : 0.0;
                                    const double2 *row = Q2 + (long long)jb[rok ? r0 + rr : r0] * ld2;
#pragma unroll
                                    for (int u = 0; u < CD_CH; ++u) {
                                        const int c = c0 + u * NTB;
                                        v[rr][u] = (c < Cp2) ? __ldg(row + c) : make_double2(0.0, 0.0);
                                    }
                                }
#pragma unroll
                                for (int rr = 0; rr < CD_RG; ++rr)
#pragma unroll
                                    for (int u = 0; u < CD_CH; ++u) {
                                        acc[u].x = fma(dl[rr], v[rr][u].x, acc[u].x);
                                        acc[u].y = fma(dl[rr], v[rr][u].y, acc[u].y);
                                    }
                            }
#pragma unroll
                            for (int u = 0; u < CD_CH; ++u) {
                                const int c = c0 + u * NTB;
                                if (c < Cp2) Qw2[c] = acc[u];
                            }
                        }
                    } else {
                        for (int k = bt; k < C; k += NTB) {
                            double a = Qw[k];
                            for (int r = 0; r < n_rows; ++r) a = fma(dlt[r], __ldg(Q + (long long)jb[r] * ldq + k), a);
                            Qw[k] = a;
                        }
                    }
                }
                cta_sync<NW>();
            }
            // sweep maxima (uniform in all threads)
            double mx[2] = {warp == 0 ? wmax_l : 0.0, warp == 0 ? dwmax_l : 0.0};
            cta_reduce<NW, 2, 0>(mx, scratch);
            const double w_max = mx[0], d_w_max = mx[1];
            if (w_max == 0.0 || d_w_max / w_max <= d_w_tol || n_iter == max_iter - 1) {
                gap = compute_gap();
                if (gap <= tol) { ++n_iter; done = true; break; }
                if (screening) screen(gap, false);
                cta_sync<NW>();
            }
        }
    }
    cta_sync<NW>();
    for (int j = tid; j < C; j += NT) W[(long long)m * ldw + j] = w[j];
    if (tid == 0) {
        info[6 * m + 0] = gap;
        info[6 * m + 1] = tol;
        info[6 * m + 2] = (double)n_iter;
        info[6 * m + 3] = (double)n_upd;
        info[6 * m + 4] = (double)n_upd + (double)n_blk * (32.0 * 32.0) / (double)C;   // rows fetched (equiv.)
        info[6 * m + 5] = (double)n_blk;
    }
}

// --------------------------------------------------------------------------- Cholesky (Ridge / OLS)
constexpr int RC_NB = 32;
constexpr int RC_THREADS = 256;

__global__ void __launch_bounds__(RC_THREADS)
ridge_cholesky_kernel(const double *__restrict__ Qc, long long ldq, const double *__restrict__ qc, int C,
                      const double *__restrict__ alpha, double *__restrict__ W, long long ldw,
                      int *__restrict__ status, double *__restrict__ work) {
    extern __shared__ __align__(16) double sh[];
    double *D = sh;                          // [32][33] diagonal block
    double *Pi = D + RC_NB * 33;             // [64][33]
    double *Pj = Pi + 64 * 33;               // [64][33]
    double *wv = Pj + 64 * 33;               // [C] solution vector
    __shared__ int bad;

    const int kq = blockIdx.x, tid = threadIdx.x;
    const double a = alpha[kq];
    double *M = work + (long long)kq * (C + 1) * ldq;
    if (tid == 0) bad = 0;

    // M = lower(Qc) + a I ; row C = qc'
    for (long long e = tid; e < (long long)(C + 1) * C; e += RC_THREADS) {
        const int i = (int)(e / C), j = (int)(e - (long long)i * C);
        if (i == C) M[(long long)i * ldq + j] = qc[j];
        else if (j <= i) M[(long long)i * ldq + j] = Qc[(long long)i * ldq + j] + (i == j ? a : 0.0);
    }
    __syncthreads();

    for (int k0 = 0; k0 < C; k0 += RC_NB) {
        const int nb = min(RC_NB, C - k0);
        // 1. diagonal block -> smem, unblocked Cholesky by warp 0
        for (int e = tid; e < nb * nb; e += RC_THREADS) {
            const int r = e / nb, c = e - r * nb;
            D[r * 33 + c] = (c <= r) ? M[(long long)(k0 + r) * ldq + k0 + c] : 0.0;
        }
        __syncthreads();
        if (tid < 32) {
            for (int c = 0; c < nb; ++c) {
                double d = D[c * 33 + c];
                if (!(d > 0.0)) { if (tid == 0) bad = 1; d = nan(""); }
                const double l = sqrt(d);
                __syncwarp();
                if (tid == 0) D[c * 33 + c] = l;
                for (int r = c + 1 + tid; r < nb; r += 32) D[r * 33 + c] /= l;
                __syncwarp();
                // trailing update inside the block: D[r][cc] -= D[r][c]*D[cc][c] for c < cc <= r
                for (int e = tid; e < nb * nb; e += 32) {
                    const int r = e / nb, cc = e - r * nb;
                    if (cc > c && cc <= r) D[r * 33 + cc] -= D[r * 33 + c] * D[cc * 33 + c];
                }
                __syncwarp();
            }
        }
        __syncthreads();
        for (int e = tid; e < nb * nb; e += RC_THREADS) {
            const int r = e / nb, c = e - r * nb;
            if (c <= r) M[(long long)(k0 + r) * ldq + k0 + c] = D[r * 33 + c];
        }
        // 2. panel solve: rows below the block (incl. the rhs row C): x L_kk' = M[i, k0:k0+nb]
        for (int i = k0 + nb + tid; i <= C; i += RC_THREADS) {
            double x[RC_NB];
            double *row = M + (long long)i * ldq + k0;
#pragma unroll
            for (int c = 0; c < RC_NB; ++c) x[c] = (c < nb) ? row[c] : 0.0;
#pragma unroll
            for (int c = 0; c < RC_NB; ++c) {
                if (c < nb) {
                    double s = x[c];
#pragma unroll
                    for (int r = 0; r < RC_NB; ++r)
                        if (r < c) s -= x[r] * D[c * 33 + r];
                    x[c] = s / D[c * 33 + c];
                }
            }
#pragma unroll
            for (int c = 0; c < RC_NB; ++c)
                if (c < nb) row[c] = x[c];
        }
        __syncthreads();
        // 3. trailing update with 64x64 tiles: M[i][j] -= sum_c M[i][k0+c] M[j][k0+c]
        const int lo = k0 + nb;
        if (lo <= C) {
            const int ty = tid >> 4, tx = tid & 15;
            for (int i0 = lo; i0 <= C; i0 += 64) {
                for (int e = tid; e < 64 * RC_NB; e += RC_THREADS) {
                    const int r = e >> 5, c = e & 31;
                    Pi[r * 33 + c] = (i0 + r <= C && c < nb) ? M[(long long)(i0 + r) * ldq + k0 + c] : 0.0;
                }
                for (int j0 = lo; j0 <= i0 && j0 < C; j0 += 64) {
                    __syncthreads();
                    for (int e = tid; e < 64 * RC_NB; e += RC_THREADS) {
                        const int r = e >> 5, c = e & 31;
                        Pj[r * 33 + c] = (j0 + r < C && c < nb) ? M[(long long)(j0 + r) * ldq + k0 + c] : 0.0;
                    }
                    __syncthreads();
                    double acc[4][4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int v = 0; v < 4; ++v) acc[u][v] = 0.0;
#pragma unroll 8
                    for (int c = 0; c < RC_NB; ++c) {
                        double pa[4], pb[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) pa[u] = Pi[(ty + 16 * u) * 33 + c];
#pragma unroll
                        for (int v = 0; v < 4; ++v) pb[v] = Pj[(tx + 16 * v) * 33 + c];
#pragma unroll
                        for (int u = 0; u < 4; ++u)
#pragma unroll
                            for (int v = 0; v < 4; ++v) acc[u][v] += pa[u] * pb[v];
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            const int i = i0 + ty + 16 * u, j = j0 + tx + 16 * v;
                            if (i <= C && j < C && j <= i) M[(long long)i * ldq + j] -= acc[u][v];
                        }
                }
                __syncthreads();
            }
        }
        __syncthreads();
    }

    // row C now holds y with L y = qc.  Back substitution L' w = y, blocked from the end.
    for (int j = tid; j < C; j += RC_THREADS) wv[j] = M[(long long)C * ldq + j];
    __syncthreads();
    const int n_blk = (C + RC_NB - 1) / RC_NB;
    for (int b = n_blk - 1; b >= 0; --b) {
        const int k0 = b * RC_NB, nb = min(RC_NB, C - k0);
        for (int e = tid; e < nb * nb; e += RC_THREADS) {
            const int r = e / nb, c = e - r * nb;
            D[r * 33 + c] = (c <= r) ? M[(long long)(k0 + r) * ldq + k0 + c] : 0.0;
        }
        __syncthreads();
        if (tid == 0) {
            for (int c = nb - 1; c >= 0; --c) {
                double s = wv[k0 + c];
                for (int r = c + 1; r < nb; ++r) s -= D[r * 33 + c] * wv[k0 + r];
                wv[k0 + c] = s / D[c * 33 + c];
            }
        }
        __syncthreads();
        for (int i = tid; i < k0; i += RC_THREADS) {
            double s = wv[i];
            for (int r = 0; r < nb; ++r) s -= M[(long long)(k0 + r) * ldq + i] * wv[k0 + r];
            wv[i] = s;
        }
        __syncthreads();
    }
    for (int j = tid; j < C; j += RC_THREADS) W[(long long)kq * ldw + j] = wv[j];
    if (tid == 0) status[kq] = bad;
}

// --------------------------------------------------------------------------- intercept + evaluation vectors
__global__ void __launch_bounds__(128)
finalize_models_kernel(const double *__restrict__ W, long long ldw, int C, int n_y,
                       const int *__restrict__ y_col_of_model, const double *const *__restrict__ xbar_of_model,
                       const double *__restrict__ ybar_of_model, double *__restrict__ intercept,
                       double *__restrict__ V, long long ldv) {
    const int m = blockIdx.x, tid = threadIdx.x;
    const double *w = W + (long long)m * ldw;
    const double *xb = xbar_of_model ? xbar_of_model[m] : nullptr;
    double s = 0.0;
    if (xb) for (int j = tid; j < C; j += 128) s += xb[j] * w[j];
    s = warp_sum(s);
    __shared__ double part[4];
    __shared__ double bsh;
    if ((tid & 31) == 0) part[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
        const double b = xb ? (ybar_of_model[m] - (part[0] + part[1] + part[2] + part[3])) : 0.0;
        intercept[m] = b;
        bsh = b;
    }
    __syncthreads();
    if (V) {
        double *v = V + (long long)m * ldv;
        const int yc = y_col_of_model ? y_col_of_model[m] : 0;
        for (int j = tid; j < ldv; j += 128) {
            double x = 0.0;
            if (j < C) x = -w[j];
            else if (j < C + n_y) x = (j - C == yc) ? 1.0 : 0.0;
            else if (j == C + n_y) x = -bsh;
            v[j] = x;
        }
    }
}

// --------------------------------------------------------------------------- quadratic forms  v' A v
constexpr int QF_MB = 4;       // models per CTA
constexpr int QF_THREADS = 256;

__global__ void __launch_bounds__(QF_THREADS)
quadform_kernel(const double *__restrict__ A, long long lda, int n, const double *__restrict__ V,
                long long ldv, int n_models, double *__restrict__ out) {
    extern __shared__ __align__(16) double vs[];   // [QF_MB][n]
    const int m0 = blockIdx.x * QF_MB, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nm = min(QF_MB, n_models - m0);
    for (int e = tid; e < QF_MB * n; e += QF_THREADS) {
        const int mm = e / n, j = e - mm * n;
        vs[e] = (mm < nm) ? V[(long long)(m0 + mm) * ldv + j] : 0.0;
    }
    __syncthreads();
    double acc[QF_MB] = {0, 0, 0, 0};
    for (int i = warp; i < n; i += QF_THREADS / 32) {
        const double *row = A + (long long)i * lda;
        double d[QF_MB] = {0, 0, 0, 0};
        for (int j = lane; j < n; j += 32) {
            const double aij = __ldg(row + j);
#pragma unroll
            for (int mm = 0; mm < QF_MB; ++mm) d[mm] += aij * vs[mm * n + j];
        }
#pragma unroll
        for (int mm = 0; mm < QF_MB; ++mm) {
            d[mm] = warp_sum(d[mm]);
            acc[mm] += d[mm] * vs[mm * n + i];
        }
    }
    __shared__ double red[QF_THREADS / 32][QF_MB];
    if (lane == 0)
        for (int mm = 0; mm < QF_MB; ++mm) red[warp][mm] = acc[mm];
    __syncthreads();
    if (tid < nm) {
        double s = 0.0;
        for (int w = 0; w < QF_THREADS / 32; ++w) s += red[w][tid];
        out[m0 + tid] = s;
    }
}

}  // namespace sglm

using namespace sglm;

extern "C" int sglm_center_stats_f64(const double *A_plus, const double *A_minus, int64_t ldg, int32_t C,
                                     int32_t n_y, int32_t y_col, int32_t fit_intercept, double *Qc,
                                     int64_t ldq, double *qc, double *xbar, double *diag, double *scal,
                                     void *stream) {
    SGLM_CHECK_ARG(C > 0 && n_y > 0 && y_col >= 0 && y_col < n_y, SGLM_E_SHAPE, "center_stats: bad shape");
    SGLM_CHECK_ARG(A_plus && Qc && qc && xbar && diag && scal, SGLM_E_INVALID_ARG, "center_stats: null pointer");
    SGLM_CHECK_ARG(ldg >= C + n_y + 1 && ldq >= C, SGLM_E_SHAPE, "center_stats: leading dimension too small");
    dim3 grid((unsigned)std::min<long long>(ceil_div<long long>(ldq, 256), 64), (unsigned)C);
    center_stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A_plus, A_minus, ldg, C, n_y, y_col,
                                                                 fit_intercept, Qc, ldq, qc, xbar, diag, scal);
    SGLM_LAUNCH_OK("center_stats_kernel");
    return SGLM_OK;
}

static size_t cd_smem_bytes(int C, int nw) {
    const int Cp = (C + 1) & ~1;
    return (size_t)(2 * Cp + nw * 8 + 32) * sizeof(double) + (size_t)(C + 32 + nw + 2) * sizeof(int) + (size_t)C;
}

extern "C" int sglm_enet_cd_gram_f64(const double *const *prob_Q, const double *const *prob_q,
                                     const double *const *prob_diag, const double *prob_yy, int64_t ldq, int32_t C,
                                     const int32_t *prob_of_model, const double *l1_reg,
                                     const double *l2_reg, const double *tol, const int32_t *max_iter,
                                     int32_t n_models, int32_t warm_start, int32_t do_screening,
                                     double *W, int64_t ldw, double *info, void *stream) {
    SGLM_CHECK_ARG(C > 0 && n_models >= 0 && ldq >= C && ldw >= C, SGLM_E_SHAPE, "enet_cd: bad shape");
    if (n_models == 0) return SGLM_OK;
    SGLM_CHECK_ARG(prob_Q && prob_q && prob_diag && prob_yy && prob_of_model && l1_reg && l2_reg && tol && max_iter && W && info,
                   SGLM_E_INVALID_ARG, "enet_cd: null pointer");
    const int nb = C <= 256 ? 1 : (C <= 512 ? 2 : (C <= 1024 ? 4 : 8));   // panel-update warps (+1 sequential warp)
    const int nw = nb + 1;
    const size_t smem = cd_smem_bytes(C, nw);
    SGLM_CHECK_ARG(smem <= 227 * 1024, SGLM_E_UNSUPPORTED,
                   "enet_cd: C=%d needs %zu bytes of shared memory per model (> 227 KB)", C, smem);
    cudaStream_t st = (cudaStream_t)stream;
#define CD_LAUNCH(N, RG, MB)                                                                               \
    do {                                                                                                   \
        SGLM_CUDA_OK(cudaFuncSetAttribute(enet_cd_gram_kernel<N, RG, MB>,                                  \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
        enet_cd_gram_kernel<N, RG, MB><<<n_models, (N + 1) * 32, smem, st>>>(                              \
            prob_Q, prob_q, prob_diag, prob_yy, ldq, C, prob_of_model, l1_reg, l2_reg, tol, max_iter,      \
            warm_start, do_screening, W, ldw, info);                                                       \
    } while (0)
    // (panel warps, rows per group, min CTAs/SM) tuned on B200: the variants without register
    // spills win; for C in (1024, 2048] one 9-warp CTA per SM (profiles/r1_cd_variants.txt)
    switch (nb) {
        case 1: CD_LAUNCH(1, 4, 4); break;
        case 2: CD_LAUNCH(2, 4, 4); break;
        case 4: CD_LAUNCH(4, 4, 2); break;
        default: CD_LAUNCH(8, 4, 1); break;
    }
#undef CD_LAUNCH
    SGLM_LAUNCH_OK("enet_cd_gram_kernel");
    return SGLM_OK;
}

extern "C" size_t sglm_ridge_workspace_bytes(int32_t C, int64_t ldq, int32_t n_alpha) {
    if (C <= 0 || ldq < C || n_alpha <= 0) return 0;
    return (size_t)n_alpha * (size_t)(C + 1) * (size_t)ldq * sizeof(double);
}

extern "C" int sglm_ridge_solve_f64(const double *Qc, int64_t ldq, const double *qc, int32_t C,
                                    const double *alpha, int32_t n_alpha, double *W, int64_t ldw,
                                    int32_t *status, void *work, size_t work_bytes, void *stream) {
    SGLM_CHECK_ARG(C > 0 && n_alpha >= 0 && ldq >= C && ldw >= C, SGLM_E_SHAPE, "ridge_solve: bad shape");
    if (n_alpha == 0) return SGLM_OK;
    SGLM_CHECK_ARG(Qc && qc && alpha && W && status && work, SGLM_E_INVALID_ARG, "ridge_solve: null pointer");
    SGLM_CHECK_ARG(work_bytes >= sglm_ridge_workspace_bytes(C, ldq, n_alpha), SGLM_E_WORKSPACE,
                   "ridge_solve: workspace too small");
    const size_t smem = (size_t)(RC_NB * 33 + 2 * 64 * 33 + C) * sizeof(double);
    SGLM_CHECK_ARG(smem <= 227 * 1024, SGLM_E_UNSUPPORTED, "ridge_solve: C=%d too large for shared memory", C);
    SGLM_CUDA_OK(cudaFuncSetAttribute(ridge_cholesky_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ridge_cholesky_kernel<<<n_alpha, RC_THREADS, smem, (cudaStream_t)stream>>>(Qc, ldq, qc, C, alpha, W, ldw, status,
                                                                              (double *)work);
    SGLM_LAUNCH_OK("ridge_cholesky_kernel");
    return SGLM_OK;
}

extern "C" int sglm_finalize_models_f64(const double *W, int64_t ldw, int32_t C, int32_t n_y,
                                        const int32_t *y_col_of_model, const double *const *xbar_of_model,
                                        const double *ybar_of_model, int32_t n_models, double *intercept,
                                        double *V, int64_t ldv, void *stream) {
    SGLM_CHECK_ARG(C > 0 && n_models >= 0 && ldw >= C, SGLM_E_SHAPE, "finalize_models: bad shape");
    if (n_models == 0) return SGLM_OK;
    SGLM_CHECK_ARG(W && intercept, SGLM_E_INVALID_ARG, "finalize_models: null pointer");
    SGLM_CHECK_ARG(!V || ldv >= C + n_y + 1, SGLM_E_SHAPE, "finalize_models: ldv too small");
    SGLM_CHECK_ARG(!xbar_of_model || ybar_of_model, SGLM_E_INVALID_ARG, "finalize_models: xbar without ybar");
    finalize_models_kernel<<<n_models, 128, 0, (cudaStream_t)stream>>>(W, ldw, C, n_y, y_col_of_model, xbar_of_model,
                                                                       ybar_of_model, intercept, V, ldv);
    SGLM_LAUNCH_OK("finalize_models_kernel");
    return SGLM_OK;
}

extern "C" int sglm_quadform_f64(const double *A, int64_t lda, int32_t n, const double *V, int64_t ldv,
                                 int32_t n_models, double *out, void *stream) {
    SGLM_CHECK_ARG(n > 0 && n_models >= 0 && lda >= n && ldv >= n, SGLM_E_SHAPE, "quadform: bad shape");
    if (n_models == 0) return SGLM_OK;
    SGLM_CHECK_ARG(A && V && out, SGLM_E_INVALID_ARG, "quadform: null pointer");
    const size_t smem = (size_t)QF_MB * n * sizeof(double);
    SGLM_CHECK_ARG(smem <= 227 * 1024, SGLM_E_UNSUPPORTED, "quadform: n=%d too large for shared memory", n);
    SGLM_CUDA_OK(cudaFuncSetAttribute(quadform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    quadform_kernel<<<ceil_div(n_models, QF_MB), QF_THREADS, smem, (cudaStream_t)stream>>>(A, lda, n, V, ldv,
                                                                                         n_models, out);
    SGLM_LAUNCH_OK("quadform_kernel");
    return SGLM_OK;
}
