// Gram coordinate descent, second generation: one thread-block CLUSTER per group of models.
//
// Replaces the same scikit-learn fits as solvers.cu's enet_cd_gram_kernel (sklearn
// linear_model/_cd_fast.pyx:1095-1290 reached from backend/sglm.py:241 for every
// (fold, alpha, l1_ratio) of backend/sglm_cv.py:106-170) and produces the same iterates;
// what changes is how one sweep is laid out on the machine.  Measured on the first
// generation (profiles/r1_cd_variants.txt, gpurun_out/cd_diag_v8.log): the grid is bound by the
// critical path of its heaviest models (482 ms alone vs 615 ms for all 1500), and per
// 32-coordinate block the sequential register phase takes 2.3 us while the panel update that
// streams the moved rows of Q takes 6 us on one SM.  Hence:
//
//   * column split over a cluster: CTA r of a K-CTA cluster owns a contiguous range of
//     coordinate blocks — their w, their slice of Qw, and their register phase.  A block's
//     deltas are published to every CTA of the cluster as a RECORD (st.async into a ring in
//     each CTA's shared memory, completion counted by an mbarrier there), and every CTA's
//     panel warps apply the record to their own columns: one model's row stream is pulled
//     through K SMs' L2 ports instead of one;
//   * look-ahead: the register warp applies its own record to the NEXT block's 32 columns in
//     registers (the off-diagonal 32x32 block is prefetched with cp.async next to the diagonal
//     one), so the next register phase starts at once and the panel update of record b runs
//     concurrently with the register phase of block b+1 — the panel leaves the critical path;
//   * M models of the same fold (same Q) share a cluster: M register warps run the block's
//     phase in parallel, the panel warps load each moved row ONCE and apply it to the M
//     models (rows of Q are the dominant traffic: bytes per model drop by up to M).
//
// The floating-point sequence per element is the sequential algorithm's: each Qw entry
// receives fma(delta_i, Q[i][k], .) in coordinate order (a zero delta contributes an exact
// +0), so iterates match the first-generation kernel bit for bit; only the duality-gap sums
// are reduced in a different order (per-CTA partials).
//
// Roofline: memory (rows of Q from L2/HBM); algorithmic bytes = 8*C per coordinate update that
// moved w, counted per model in info[6m+3] exactly as before.
#pragma once
#include <algorithm>

#include <cuda.h>

#include "common.cuh"

namespace sglm {
namespace cdc {

constexpr int RING = 16;   // records in flight per cluster (flow-controlled by completion counters)
constexpr int RS = 40;     // doubles per (record, model): 32 deltas + zeros that partial load groups read

struct Layout {
    int NBLK, BPC, CoP;
    size_t w, Qw, blk, qd, rdelta, rmask, full, bbar, xch, par, act, drp, pdone, ccnt, pmv, total;
};

__host__ __device__ inline Layout make_layout(int M, int K, int NB, int C) {
    Layout L;
    L.NBLK = (C + 31) >> 5;
    L.BPC = (L.NBLK + K - 1) / K;
    L.CoP = L.BPC * 32;
    size_t o = 0;
    L.w = o;      o += (size_t)M * L.CoP * 8;
    L.Qw = o;     o += (size_t)M * L.CoP * 8;
    L.blk = o;    o += (size_t)4 * 1024 * 8;          // [buf 2][diag, off][32][32]
    L.qd = o;     o += (size_t)2 * 64 * 8;            // [buf 2][q, diag][32]
    L.rdelta = o; o += (size_t)(RING * M + 1) * RS * 8;   // + one all-zero page (models that did not move)
    L.rmask = o;  o += (size_t)RING * M * 8;
    L.full = o;   o += (size_t)RING * 8;
    L.bbar = o;   o += (size_t)2 * 8;               // sub-block buffers landed
    L.xch = o;    o += (size_t)2 * K * M * 8 * 8;     // [parity][cta][model][8]
    L.par = o;    o += (size_t)M * 8 * 8;             // per model: l1, l2, d_w_tol, tol, max_iter, flags, gap, n_iter
    L.act = o;    o += (size_t)M * L.NBLK * 4;
    L.drp = o;    o += (size_t)M * L.NBLK * 4;
    L.pdone = o;  o += (size_t)NB * 4;
    L.ccnt = o;   o += (size_t)K * NB * 4;      // records completed per (CTA, panel warp) of the cluster
    L.pmv = o;    o += (size_t)M * L.NBLK;                // coordinates moved per (model, block) in the last sweep
    L.total = (o + 15) & ~(size_t)15;
    return L;
}

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ unsigned s_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned cluster_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned map_to(unsigned addr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(unsigned a, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_remote_arrive(unsigned ra) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ void mbar_remote_arrive_tx(unsigned ra, unsigned tx) {
    asm volatile("mbarrier.arrive.expect_tx.relaxed.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(ra), "r"(tx) : "memory");
}
__device__ __forceinline__ void st_async_b64(unsigned ra, unsigned long long v, unsigned rmbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(ra), "l"(v), "r"(rmbar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned a, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded waits: a protocol bug must surface as a trapped kernel (an error the host reports), never as a hung
// GPU.  No legitimate wait of this kernel spans more than one sweep of one model (milliseconds).
struct SpinGuard {
    unsigned n = 0;
    long long t0 = 0;
    __device__ __forceinline__ void tick() {
        if ((++n & 0xfffu) == 0u) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 40000000000LL) asm volatile("trap;");          // ~20 s
        }
    }
};
__device__ __forceinline__ void mbar_wait(unsigned a, unsigned parity) {
    SpinGuard g;
    while (!mbar_try_wait(a, parity)) g.tick();
}
__device__ __forceinline__ void st_remote_f64(unsigned ra, double v) {
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ra), "d"(v) : "memory");
}
__device__ __forceinline__ void st_remote_u32(unsigned ra, unsigned v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(ra), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(s_addr(p)) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_acquire_cluster_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.cluster.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(s_addr(p)) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed_cluster_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.cluster.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(s_addr(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_remote_u32(unsigned ra, unsigned v) {
    asm volatile("st.relaxed.cluster.shared::cluster.u32 [%0], %1;" ::"r"(ra), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_u32(unsigned *p, unsigned v) {
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(s_addr(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s_addr(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s_addr(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
template <int NTHREADS>
__device__ __forceinline__ void seq_bar() {
    if (NTHREADS == 32) __syncwarp();
    else asm volatile("bar.sync 1, %0;" ::"n"(NTHREADS) : "memory");
}

// duality gap from the reduced sums (sklearn _cd_fast.pyx:1006-1092 gap_enet_gram); s = {w'w, w'q,
// w'Qw, |w|_1, sum xta^2, max |xta|}
__device__ __forceinline__ double gap_from_sums(const double *s, double yy, double l1, double l2, double &dual_norm) {
    const double R2 = yy + s[2] - 2.0 * s[1];
    const double Ry = yy - s[1];
    const double w22 = (l2 > 0.0) ? s[0] : 0.0;
    if (l1 == 0.0) {
        dual_norm = s[4];
        if (l2 == 0.0) return s[4];
        return R2 + 0.5 * l2 * w22 - Ry + 1.0 / (2.0 * l2) * s[4];
    }
    dual_norm = s[5];
    const double primal = 0.5 * (R2 + l2 * w22) + l1 * s[3];
    const double scale = (dual_norm > l1) ? l1 / dual_norm : 1.0;
    const double dual = -0.5 * scale * scale * (R2 + l2 * w22) + scale * Ry;
    return primal - dual;
}

template <int M, int K, int NB, int CH, int RG, int MINB>
__global__ void __launch_bounds__((M + NB) * 32, MINB)
enet_cd_cluster_kernel(const double *const *__restrict__ prob_Q, const double *const *__restrict__ prob_q,
                       const double *const *__restrict__ prob_diag, const double *__restrict__ prob_yy,
                       long long ldq, int C, const int *__restrict__ prob_of_group,
                       const int *__restrict__ model_of_slot, const double *__restrict__ l1_reg,
                       const double *__restrict__ l2_reg, const double *__restrict__ tol_in,
                       const int *__restrict__ max_iter_in, int warm_start, int do_screening,
                       double *__restrict__ W, long long ldw, double *__restrict__ info,
                       const CUtensorMap *__restrict__ tmaps, double *__restrict__ group_stats) {
    constexpr int NT = (M + NB) * 32;
    constexpr int NTB = NB * 32;
    extern __shared__ __align__(128) unsigned char smraw[];
    const Layout L = make_layout(M, K, NB, C);
    double *w_s = reinterpret_cast<double *>(smraw + L.w);          // [M][CoP]
    double *Qw_s = reinterpret_cast<double *>(smraw + L.Qw);        // [M][CoP]
    double *blk = reinterpret_cast<double *>(smraw + L.blk);
    double *qd = reinterpret_cast<double *>(smraw + L.qd);
    double *rdelta = reinterpret_cast<double *>(smraw + L.rdelta);  // [RING][M][32]
    unsigned long long *rmask = reinterpret_cast<unsigned long long *>(smraw + L.rmask);   // [RING][M]
    unsigned long long *fullb = reinterpret_cast<unsigned long long *>(smraw + L.full);
    unsigned long long *bbar = reinterpret_cast<unsigned long long *>(smraw + L.bbar);
    double *xch = reinterpret_cast<double *>(smraw + L.xch);
    double *par = reinterpret_cast<double *>(smraw + L.par);
    unsigned *act = reinterpret_cast<unsigned *>(smraw + L.act);    // [M][NBLK]
    unsigned *drp = reinterpret_cast<unsigned *>(smraw + L.drp);    // [M][NBLK]
    unsigned *pdone = reinterpret_cast<unsigned *>(smraw + L.pdone);
    unsigned *ccnt = reinterpret_cast<unsigned *>(smraw + L.ccnt);
    unsigned char *pmv = smraw + L.pmv;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned rank = (K > 1) ? cluster_rank() : 0u;
    const int grp = blockIdx.x / K;
    // Warp roles.  The register warps sit on warp ids 0, 4, 8, ... — all on the same SM sub-partition — when the
    // CTA has enough warps: their dependent chains (one shuffle and three FP64 operations per coordinate) then
    // do not queue behind the panel warps' loads and FMAs in the same scheduler (measured: the register phase
    // took 2.5 us per block next to busy panel warps against 1.56 us alone, profiles/r1_cd_cluster.txt).
    constexpr bool SPREAD = 4 * (M - 1) < M + NB;
    const bool is_seq = SPREAD ? ((warp & 3) == 0 && (warp >> 2) < M) : (warp < M);
    const int m = SPREAD ? (warp >> 2) : warp;            // model slot of a register warp
    const int pw = SPREAD ? warp - min(M, (warp + 3) >> 2) : warp - M;
    const int bt = pw * 32 + lane;
    const int NBLK = L.NBLK, CoP = L.CoP;
    const int blo = min(NBLK, (int)rank * L.BPC), bhi = min(NBLK, blo + L.BPC);
    const int col0 = blo * 32;
    const int Co = max(0, min(C, bhi * 32) - col0);
    const int Co2 = (Co + 1) >> 1;
    const int pid = prob_of_group[grp];
    const double *__restrict__ Q = prob_Q[pid];
    const double *__restrict__ q = prob_q[pid];
    const double *__restrict__ dg = prob_diag[pid];
    const double yy = prob_yy[pid];
    const long long ld2 = ldq >> 1;

    // ---------------------------------------------------------------- setup
    if (tid < M) {
        const int md = model_of_slot[grp * M + tid];
        double *p = par + tid * 8;
        if (md >= 0) {
            const double l1 = l1_reg[md], tolr = tol_in[md];
            p[0] = l1; p[1] = l2_reg[md]; p[2] = tolr; p[3] = tolr * yy; p[4] = (double)max_iter_in[md];
            p[5] = (do_screening && l1 != 0.0) ? 3.0 : 1.0;      // bit 0: valid, bit 1: screening
        } else {
            p[0] = 1.0; p[1] = 1.0; p[2] = 1.0; p[3] = 0.0; p[4] = 0.0; p[5] = 0.0;
        }
        p[6] = 0.0; p[7] = 0.0;
    }
    for (int i = tid; i < M * CoP; i += NT) { w_s[i] = 0.0; Qw_s[i] = 0.0; }
    for (int i = tid; i < (RING * M + 1) * RS; i += NT) rdelta[i] = 0.0;      // entries 32.. of every page stay zero
    for (int i = tid; i < NB; i += NT) pdone[i] = 0u;
    for (int i = tid; i < K * NB; i += NT) ccnt[i] = 0u;
    if (tid == 0) {
        for (int s = 0; s < RING; ++s) mbar_init(s_addr(fullb + s), M);
        mbar_init(s_addr(bbar), 1);
        mbar_init(s_addr(bbar + 1), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned alive = 0u;
#pragma unroll
    for (int mm = 0; mm < M; ++mm) if (par[mm * 8 + 5] != 0.0) alive |= 1u << mm;
    for (int i = tid; i < M * NBLK; i += NT) {
        const int mm = i / NBLK, b = i - mm * NBLK;
        const int rem = C - (b << 5);
        act[i] = ((alive >> mm) & 1u) ? (rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u)) : 0u;
        drp[i] = 0u;
        pmv[i] = 0;
    }
    cluster_sync_all();           // barriers initialised cluster-wide before any remote arrive

    const int mdl = is_seq ? model_of_slot[grp * M + m] : -1;
    const bool valid = mdl >= 0;
    const double l1 = is_seq ? par[m * 8 + 0] : 0.0, l2 = is_seq ? par[m * 8 + 1] : 0.0;
    double my_gap = 0.0, my_dual = 0.0;
    double wmax_l = 0.0, dwmax_l = 0.0;
    long long n_upd = 0, n_blk = 0, t_p1 = 0, t_wait = 0, t_pub = 0, t_la = 0, t_top = 0, t_end = 0;
    const int dbg_sel = warm_start >> 8;      // diagnostics: which timer info[6m+5] reports
    warm_start &= 1;
    const long long t_begin = clock64();
    const unsigned long long gt_begin = global_timer_ns();
    unsigned rc = 0u;             // records published so far (identical in every thread of the cluster)
    long long rows_loaded = 0;    // panel warps: rows of Q the records asked for (union over the M models)
    unsigned xpar = 0u;
    int sweep = 0;

    auto group_active = [&](int b) -> bool {
        bool ga = false;
#pragma unroll
        for (int mm = 0; mm < M; ++mm) ga |= ((alive >> mm) & 1u) && act[mm * NBLK + b] != 0u;
        return ga;
    };
    auto next_ga = [&](int b) -> int {
        while (b < NBLK && !group_active(b)) ++b;
        return b;
    };
    // warp-local  Qw_own += a * Q[j, own columns]   (warm start, screening drops: rare)
    auto axpy_row_own = [&](int j, double a) {
        const double *row = Q + (long long)j * ldq + col0;
        double *dst = Qw_s + m * CoP;
        for (int k = lane; k < Co; k += 32) dst[k] += a * __ldg(row + k);
        __syncwarp();
    };

    if (is_seq && valid && warm_start) {
        for (int k = lane; k < Co; k += 32) w_s[m * CoP + k] = __ldcg(W + (long long)mdl * ldw + col0 + k);
        for (int j = 0; j < C; ++j) {
            const double wj = __ldcg(W + (long long)mdl * ldw + j);
            if (wj != 0.0) { axpy_row_own(j, wj); if (rank == 0) ++n_upd; }
        }
        __syncwarp();
    }

    // ---------------------------------------------------------------- end of a round: reductions over
    // the cluster, stopping rule, gap-safe screening (sklearn _cd_fast.pyx:1187-1208, :1246-1279)
    auto round_end = [&](bool initial) {
        if (is_seq) {
            if (!initial) {                                   // own panel warps have applied every record
                for (int i = 0; i < NB; ++i)
                    { SpinGuard g; while ((int)(ld_acquire_u32(pdone + i) - rc) < 0) { __nanosleep(256); g.tick(); } }
            }
            double v0 = 0, v1 = 0, v2 = 0, v3 = 0, v4 = 0, v5 = 0;
            if ((alive >> m) & 1u) {
                for (int k = lane; k < Co; k += 32) {
                    const double wj = w_s[m * CoP + k], Qwj = Qw_s[m * CoP + k], qj = __ldg(q + col0 + k);
                    v0 += wj * wj; v1 += wj * qj; v2 += wj * Qwj; v3 += fabs(wj);
                    const double xta = (l1 == 0.0) ? (qj - Qwj) : (qj - Qwj - l2 * wj);
                    v4 += xta * xta;
                    v5 = fmax(v5, fabs(xta));
                }
            }
            v0 = warp_sum(v0); v1 = warp_sum(v1); v2 = warp_sum(v2); v3 = warp_sum(v3); v4 = warp_sum(v4);
            v5 = warp_max(v5);
            const double v6 = warp_max(wmax_l), v7 = warp_max(dwmax_l);
            wmax_l = 0.0; dwmax_l = 0.0;
            if (lane < 8) {
                const double val = lane == 0 ? v0 : lane == 1 ? v1 : lane == 2 ? v2 : lane == 3 ? v3
                                 : lane == 4 ? v4 : lane == 5 ? v5 : lane == 6 ? v6 : v7;
                const unsigned la = s_addr(xch + ((xpar * K + rank) * M + m) * 8 + lane);
#pragma unroll
                for (int k = 0; k < K; ++k) st_remote_f64(map_to(la, k), val);
            }
        }
        cluster_sync_all();
        unsigned scr = 0u;
#pragma unroll
        for (int mm = 0; mm < M; ++mm) {
            if (!((alive >> mm) & 1u)) continue;
            double s[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) s[e] = 0.0;
            for (int k = 0; k < K; ++k) {
                const double *x = xch + ((xpar * K + k) * M + mm) * 8;
#pragma unroll
                for (int e = 0; e < 5; ++e) s[e] += x[e];
#pragma unroll
                for (int e = 5; e < 8; ++e) s[e] = fmax(s[e], x[e]);
            }
            const double *p = par + mm * 8;
            const double pl1 = p[0], pl2 = p[1], d_w_tol = p[2], tol = p[3];
            const int max_iter = (int)p[4];
            const bool screening = ((int)p[5] & 2) != 0;
            bool done = false, have_gap = false;
            double gap = 0.0, dual = 0.0;
            if (initial) {
                gap = gap_from_sums(s, yy, pl1, pl2, dual);
                have_gap = true;
                done = (gap >= 0.0 && gap <= tol) || max_iter <= 0;
                if (!done && screening) scr |= 1u << mm;
            } else {
                const double w_max = s[6], d_w_max = s[7];
                if (w_max == 0.0 || d_w_max / w_max <= d_w_tol || sweep == max_iter - 1) {
                    gap = gap_from_sums(s, yy, pl1, pl2, dual);
                    have_gap = true;
                    if (gap <= tol) done = true;
                    else if (screening) scr |= 1u << mm;
                }
                if (sweep + 1 >= max_iter) done = true;
            }
            if (have_gap && is_seq && mm == m) { my_gap = gap; my_dual = dual; }
            if (tid == 0) {
                if (have_gap) par[mm * 8 + 6] = gap;
                if (done) par[mm * 8 + 7] = (double)(initial ? 0 : sweep + 1);
            }
            if (done) alive &= ~(1u << mm);
        }
        if (!initial) ++sweep;
        if (scr) {
            if (is_seq && ((scr >> m) & 1u)) {
                const double radius = sqrt(2.0 * fabs(my_gap)) / l1;
                const double denom = fmax(l1, my_dual);
                for (int b = blo; b < bhi; ++b) {
                    const int j = (b << 5) + lane, k = j - col0;
                    const bool in = (j < C) && ((act[m * NBLK + b] >> lane) & 1u);
                    bool keep = false, drop_nz = false;
                    if (in) {
                        const double djj = __ldg(dg + j);
                        if (initial && djj == 0.0) {
                            w_s[m * CoP + k] = 0.0;
                        } else {
                            const double wj = w_s[m * CoP + k];
                            const double xta = __ldg(q + j) - Qw_s[m * CoP + k] - l2 * wj;
                            const double d_j = (1.0 - fabs(xta / denom)) / sqrt(djj + l2);
                            if (d_j <= radius) keep = true;
                            else if (wj != 0.0) {
                                drop_nz = true;
                                __stcg(W + (long long)mdl * ldw + j, wj);      // pre-drop value, read by every CTA
                            }
                        }
                    }
                    const unsigned km = __ballot_sync(0xffffffffu, keep), dm = __ballot_sync(0xffffffffu, drop_nz);
                    if (lane < K) {
                        st_remote_u32(map_to(s_addr(act + m * NBLK + b), lane), km);
                        st_remote_u32(map_to(s_addr(drp + m * NBLK + b), lane), dm);
                    }
                }
                __threadfence();
            }
            cluster_sync_all();
            if (is_seq && ((scr >> m) & 1u)) {
                for (int b = 0; b < NBLK; ++b) {
                    unsigned dm = drp[m * NBLK + b];
                    while (dm) {
                        const int j = (b << 5) + __ffs(dm) - 1;
                        dm &= dm - 1;
                        const double wj = __ldcg(W + (long long)mdl * ldw + j);
                        axpy_row_own(j, -wj);
                        if (lane == 0 && j >= col0 && j < col0 + Co) w_s[m * CoP + j - col0] = 0.0;
                        if (rank == 0) ++n_upd;
                        __syncwarp();
                    }
                }
            }
            __syncthreads();
        }
        xpar ^= 1u;
    };

    round_end(true);

    // asynchronous fetch of block b by ONE thread: two TMA boxes of 32x32 doubles — the diagonal sub-block
    // and the off-diagonal sub-block Q[b rows, nxt columns] for the look-ahead — completion counted on
    // bbar[buf]; rows / columns past C are zero-filled by the TMA unit.  (Per-row copies issued from the
    // register warp cost 1.9 us per block of its critical path: gpurun_out/cdc_full4.log.)
    const CUtensorMap *tmap = tmaps + pid;
    auto prefetch = [&](int b, int nxt, int buf) {
        if (lane != 0) return;
        const bool want_off = nxt < bhi;
        const unsigned bar = s_addr(bbar + buf);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(want_off ? 16384u : 8192u) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(s_addr(blk + (buf * 2 + 0) * 1024)), "l"((unsigned long long)tmap), "r"(bar), "r"(b << 5), "r"(b << 5) : "memory");
        if (want_off)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(s_addr(blk + (buf * 2 + 1) * 1024)), "l"((unsigned long long)tmap), "r"(bar), "r"(nxt << 5), "r"(b << 5) : "memory");
    };

    // ---------------------------------------------------------------- sweeps
    unsigned bph = 0u;                 // phase parity of bbar[0], bbar[1]
    while (alive) {
        if (is_seq) {
            int buf = 0;
            double q_nx = 0.0, d_nx = 0.0;
            {
                const int b1 = next_ga(blo);
                if (b1 < bhi) {
                    if (m == 0) prefetch(b1, next_ga(b1 + 1), 0);
                    const int j = (b1 << 5) + lane;
                    if (j < C) { q_nx = __ldg(q + j); d_nx = __ldg(dg + j); }
                }
            }
            bool carried = false;
            double Qw_l = 0.0;
            for (int b = 0; b < NBLK; ++b) {
                if (!group_active(b)) continue;
                const unsigned r = rc++;
                if (b < blo || b >= bhi) continue;
                const long long t_0 = clock64();
                const int nxt = next_ga(b + 1);
                const double q_l = q_nx, d_l = d_nx;
                if (M > 1) seq_bar<M * 32>();                        // every register warp is done with buffer buf^1
                if (nxt < bhi) {
                    if (m == 0) prefetch(nxt, next_ga(nxt + 1), buf ^ 1);
                    const int j = (nxt << 5) + lane;
                    q_nx = 0.0; d_nx = 0.0;
                    if (j < C) { q_nx = __ldg(q + j); d_nx = __ldg(dg + j); }
                }
                {   // block b's sub-blocks have landed
                    const unsigned bar = s_addr(bbar + buf), ph = (bph >> buf) & 1u;
                    unsigned ok;
                    SpinGuard guard;
                    do {
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                     : "=r"(ok) : "r"(bar), "r"(ph) : "memory");
                        if (!ok) guard.tick();
                    } while (!ok);
                    bph ^= 1u << buf;
                }
                t_top += clock64() - t_0;
                const bool live = (alive >> m) & 1u;
                const unsigned mask = live ? act[m * NBLK + b] : 0u;
                const int kl = ((b - blo) << 5) + lane;
                unsigned nz = 0u;
                double delta_l = 0.0;
                if (mask) {
                    if (!carried) {
                        const long long t_w = clock64();
                        const unsigned *pd = pdone + (((b - blo) << 4) % NTB >> 5);
                        { SpinGuard g; while ((int)(ld_acquire_u32(pd) - r) < 0) { __nanosleep(64); g.tick(); } }      // usually the hand-over between CTAs
                        Qw_l = Qw_s[m * CoP + kl];
                        t_wait += clock64() - t_w;
                    }
                    const long long t_b = clock64();
                    const double *S = blk + (buf * 2) * 1024;
                    const bool ok_l = ((mask >> lane) & 1u) && d_l != 0.0;
                    const double w_l = w_s[m * CoP + kl];
                    // soft threshold as one FMA per branch (see solvers.cu): with r = (q + w d) - Qw,
                    //   delta = r > l1 ? (r - l1)/den - w : r < -l1 ? (r + l1)/den - w : -w
                    // Lanes whose coordinate is screened out (or has a zero diagonal) get all-zero
                    // constants, so their delta is an exact 0 without a select on the dependent chain.
                    const double inv_l = ok_l ? 1.0 / (d_l + l2) : 0.0;
                    const double a_l = fma(w_l, d_l, q_l);
                    const double negw = ok_l ? -w_l : 0.0;
                    const double k_pos = ok_l ? fma(-l1, inv_l, -w_l) : 0.0, k_neg = ok_l ? fma(l1, inv_l, -w_l) : 0.0;
                    // Candidates of all 32 lanes from the current state; lanes whose candidate is zero need no
                    // turn (see solvers.cu): dense blocks run the straight 32-step chain, sparse ones jump from
                    // mover to mover.
                    // r = (q + w d) - Qw carried directly (see solvers.cu): one FMA per step on the dependent chain
                    double r_l = a_l - Qw_l;
                    auto candidate = [&]() -> double {
                        const double rr = r_l;
                        const double dpos = fma(rr, inv_l, k_pos), dneg = fma(rr, inv_l, k_neg);
                        return cd_soft_select(rr, l1, dpos, dneg, negw);
                    };
                    if (pmv[m * NBLK + b] >= 20) {
                        // dense block (20+ coordinates moved in the last sweep): straight-line, a screened-out
                        // coordinate contributes delta = 0 (exact no-op FMAs), no vote on the dependent path
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const double s_il = S[i * 32 + lane];
                            const double dc = candidate();
                            const double di = __shfl_sync(0xffffffffu, dc, i);
                            if (lane == i) delta_l = dc;
                            r_l = fma(-di, s_il, r_l);
                        }
                    } else {
                        double dc = candidate();
                        unsigned todo = mask;
                        while (true) {
                            const unsigned mv = __ballot_sync(0xffffffffu, dc != 0.0) & todo;
                            if (!mv) break;
                            const int i = __ffs(mv) - 1;
                            const double di = __shfl_sync(0xffffffffu, dc, i);
                            if (lane == i) delta_l = dc;
                            r_l = fma(-di, S[i * 32 + lane], r_l);
                            todo &= ~((2u << i) - 1u);                    // coordinates up to i have had their turn
                            if (!todo) break;
                            dc = candidate();
                        }
                    }
                    const double w_new_l = w_l + delta_l;
                    nz = __ballot_sync(0xffffffffu, delta_l != 0.0);
                    if (lane == 0) pmv[m * NBLK + b] = (unsigned char)__popc(nz);
                    if (ok_l) {
                        dwmax_l = fmax(dwmax_l, fabs(delta_l));
                        wmax_l = fmax(wmax_l, fabs(w_new_l));
                    }
                    if (delta_l != 0.0) w_s[m * CoP + kl] = w_new_l;
                    n_upd += __popc(nz);
                    ++n_blk;
                    t_p1 += clock64() - t_b;
                }
                // look-ahead source: the next block's Qw with every record before b applied
                const unsigned mask_n = (live && nxt < bhi) ? act[m * NBLK + nxt] : 0u;
                double Qn_l = 0.0;
                const long long t_1 = clock64();
                if (mask_n) {
                    const unsigned *pd = pdone + (((nxt - blo) << 4) % NTB >> 5);
                    { SpinGuard g; while ((int)(ld_acquire_u32(pd) - r) < 0) { __nanosleep(32); g.tick(); } }
                    Qn_l = Qw_s[m * CoP + ((nxt - blo) << 5) + lane];
                }
                const long long t_2 = clock64();
                t_wait += t_2 - t_1;
                // publish record r (deltas + moved mask of this model) to every CTA of the cluster
                {
                    const unsigned slot = r % RING, use = r / RING;
                    if (use > 0) {
                        // ring slot free: every panel warp of the cluster has completed record r - RING
                        // (monotonic counters: a register warp may be laps ahead of or behind the others,
                        // which a phase-parity wait could not tell apart)
                        const unsigned need = r - RING + 1u;
                        bool ok;
                        SpinGuard guard;
                        do {
                            ok = true;
                            for (int idx = lane; idx < K * NB; idx += 32)
                                ok &= (int)(ld_relaxed_cluster_u32(ccnt + idx) - need) >= 0;
                            if (!__all_sync(0xffffffffu, ok)) { ok = false; __nanosleep(64); guard.tick(); } else ok = true;
                        } while (!ok);
                    }
                    // groups of 8: the deltas of a record are stored row-major, [row][model], so that a panel thread fetches
                    // the 8 deltas of a row with four 16-byte loads; a model that did not move then has to send its zeros
                    // (there is no all-zero page to point at)
                    constexpr bool ROWMAJOR = M >= 8;
                    const bool send = ROWMAJOR || nz != 0u;
                    const unsigned la_d = ROWMAJOR ? s_addr(rdelta + ((size_t)slot * RS + lane) * M + m)
                                                   : s_addr(rdelta + (slot * M + m) * RS + lane);
                    const unsigned la_m = s_addr(rmask + slot * M + m);
                    const unsigned la_b = s_addr(fullb + slot);
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const unsigned rb = map_to(la_b, k);
                        if (lane == 0) mbar_remote_arrive_tx(rb, send ? 33u * 8u : 8u);
                        if (send) st_async_b64(map_to(la_d, k), (unsigned long long)__double_as_longlong(delta_l), rb);
                        if (lane == 0) st_async_b64(map_to(la_m, k), (unsigned long long)nz, rb);
                    }
                }
                const long long t_3 = clock64();
                t_pub += t_3 - t_2;
                if (mask_n) {
                    if (nz) {
                        // rows that did not move carry delta = 0: exact no-ops, no branches on the chain
                        const double *So = blk + (buf * 2 + 1) * 1024;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const double di = __shfl_sync(0xffffffffu, delta_l, i);
                            Qn_l = fma(di, So[i * 32 + lane], Qn_l);
                        }
                    }
                    Qw_l = Qn_l;
                    carried = true;
                } else {
                    carried = false;
                }
                t_la += clock64() - t_3;
                buf ^= 1;
            }
        } else {
            // ------------------------------------------------------------ panel warps: apply records in order
            double2 *Qw2 = reinterpret_cast<double2 *>(Qw_s);
            const int CoP2 = CoP >> 1;
            for (int b = 0; b < NBLK; ++b) {
                if (!group_active(b)) continue;
                const unsigned r = rc++;
                const unsigned slot = r % RING, use = r / RING;
                mbar_wait(s_addr(fullb + slot), use & 1u);
                unsigned mk[M], um = 0u;
#pragma unroll
                for (int mm = 0; mm < M; ++mm) { mk[mm] = (unsigned)rmask[slot * M + mm]; um |= mk[mm]; }
                rows_loaded += __popc(um);          // rows of Q this record makes the cluster load (once for the M models)
                if (um) {
                    const double2 *Q2 = reinterpret_cast<const double2 *>(Q + (long long)(b << 5) * ldq + col0);
                    // A model that moved sent all 32 deltas (exact zeros for the rows that stayed), a model that
                    // did not move reads the zero page: no per-row selects in the loop below.
                    const double *dlm[M];
#pragma unroll
                    for (int mm = 0; mm < M; ++mm) dlm[mm] = rdelta + (mk[mm] ? (slot * M + mm) : RING * M) * RS;
                    for (int c0 = bt; c0 < Co2; c0 += NTB * CH) {
                        double2 acc[M][CH];
                        bool okc[CH];
#pragma unroll
                        for (int u = 0; u < CH; ++u) {
                            okc[u] = c0 + u * NTB < Co2;
#pragma unroll
                            for (int mm = 0; mm < M; ++mm)
                                acc[mm][u] = okc[u] ? Qw2[mm * CoP2 + c0 + u * NTB] : make_double2(0.0, 0.0);
                        }
                        unsigned rem = um;
                        if constexpr (M >= 8) {
                            // wide groups: the deltas of a row ([row][model] in the record) are fetched with M / 2 16-byte
                            // loads right before the row's FMAs (M * RG of them in registers next to M * CH accumulators
                            // would not fit into the 128 registers of a 512-thread CTA)
                            const double2 *drec = reinterpret_cast<const double2 *>(rdelta + (size_t)slot * RS * M);
                            while (rem) {
                                double2 v[RG][CH];
                                int ii[RG];
#pragma unroll
                                for (int g = 0; g < RG; ++g) {
                                    const bool has = rem != 0u;
                                    const int i = has ? (__ffs(rem) - 1) : 32;        // 32: a row of zero deltas
                                    rem &= rem - 1;
                                    ii[g] = i;
                                    const double2 *row = Q2 + (long long)(i & 31) * ld2 + c0;
#pragma unroll
                                    for (int u = 0; u < CH; ++u)
                                        v[g][u] = okc[u] ? __ldg(row + u * NTB) : make_double2(0.0, 0.0);
                                }
#pragma unroll
                                for (int g = 0; g < RG; ++g) {
                                    const double2 *dp = drec + ii[g] * (M / 2);
                                    double2 dd[M / 2];
#pragma unroll
                                    for (int h = 0; h < M / 2; ++h) dd[h] = dp[h];
#pragma unroll
                                    for (int h = 0; h < M / 2; ++h)
#pragma unroll
                                        for (int u = 0; u < CH; ++u) {
                                            acc[2 * h][u].x = fma(dd[h].x, v[g][u].x, acc[2 * h][u].x);
                                            acc[2 * h][u].y = fma(dd[h].x, v[g][u].y, acc[2 * h][u].y);
                                            acc[2 * h + 1][u].x = fma(dd[h].y, v[g][u].x, acc[2 * h + 1][u].x);
                                            acc[2 * h + 1][u].y = fma(dd[h].y, v[g][u].y, acc[2 * h + 1][u].y);
                                        }
                                }
                            }
                        } else {
                        while (rem) {
                            double2 v[RG][CH];
                            double d[RG][M];
#pragma unroll
                            for (int g = 0; g < RG; ++g) {
                                const bool has = rem != 0u;
                                const int i = has ? (__ffs(rem) - 1) : 32;    // 32: zero delta, row 0 of the block
                                rem &= rem - 1;                               // (0 & -1) stays 0
#pragma unroll
                                for (int mm = 0; mm < M; ++mm) d[g][mm] = dlm[mm][i];
                                const double2 *row = Q2 + (long long)(i & 31) * ld2 + c0;
#pragma unroll
                                for (int u = 0; u < CH; ++u)
                                    v[g][u] = okc[u] ? __ldg(row + u * NTB) : make_double2(0.0, 0.0);
                            }
#pragma unroll
                            for (int g = 0; g < RG; ++g)
#pragma unroll
                                for (int u = 0; u < CH; ++u)
#pragma unroll
                                    for (int mm = 0; mm < M; ++mm) {
                                        acc[mm][u].x = fma(d[g][mm], v[g][u].x, acc[mm][u].x);
                                        acc[mm][u].y = fma(d[g][mm], v[g][u].y, acc[mm][u].y);
                                    }
                        }
                        }
#pragma unroll
                        for (int u = 0; u < CH; ++u)
                            if (okc[u]) {
#pragma unroll
                                for (int mm = 0; mm < M; ++mm) Qw2[mm * CoP2 + c0 + u * NTB] = acc[mm][u];
                            }
                    }
                }
                __syncwarp();
                if (lane == 0) st_release_u32(pdone + pw, r + 1u);
                // slot-reuse credit: the record's deltas were consumed by the FMAs above (their results are
                // already stored), so a relaxed store is enough and keeps a cluster-scope fence off this path
                if (lane < K) st_relaxed_remote_u32(map_to(s_addr(ccnt + rank * NB + pw), lane), r + 1u);
            }
        }
        { const long long t_e = clock64(); round_end(false); t_end += clock64() - t_e; }
    }

    // ---------------------------------------------------------------- results
    if (is_seq && valid) {
        for (int k = lane; k < Co; k += 32) W[(long long)mdl * ldw + col0 + k] = w_s[m * CoP + k];
        if (lane < 3) {
            const long long t_rep = dbg_sel == 1 ? t_wait : dbg_sel == 2 ? t_pub : dbg_sel == 3 ? t_la : dbg_sel == 4 ? t_top : dbg_sel == 5 ? t_end : t_p1;
            const double val = lane == 0 ? (double)n_upd : lane == 1 ? (double)n_blk : (double)t_rep;
            st_remote_f64(map_to(s_addr(xch + ((xpar * K + rank) * M + m) * 8 + lane), 0), val);
        }
    }
    cluster_sync_all();
    if (group_stats && rank == 0 && !is_seq && bt == 0) {
        group_stats[2 * grp + 0] = (double)rows_loaded;      // each row = 8*C bytes through the cluster's L2 ports
        group_stats[2 * grp + 1] = (double)rc;               // records (coordinate blocks) processed
    }
    if (rank == 0 && is_seq && valid && lane == 0) {
        double su = 0.0, sb = 0.0, st = 0.0;
        for (int k = 0; k < K; ++k) {
            const double *x = xch + ((xpar * K + k) * M + m) * 8;
            su += x[0]; sb += x[1]; st += x[2];
        }
        info[6 * mdl + 0] = par[m * 8 + 6];
        info[6 * mdl + 1] = par[m * 8 + 3];
        info[6 * mdl + 2] = par[m * 8 + 7];
        info[6 * mdl + 3] = su;
        info[6 * mdl + 4] = sb;
        info[6 * mdl + 5] = dbg_sel == 6 ? (double)gt_begin : dbg_sel == 7 ? (double)(global_timer_ns() - gt_begin)
                                         : st / (double)max(1LL, clock64() - t_begin);
    }
}

template <int M, int K, int NB, int CH, int RG, int MINB>
static int launch(const double *const *prob_Q, const double *const *prob_q, const double *const *prob_diag,
                  const double *prob_yy, long long ldq, int C, const int *prob_of_group, const int *model_of_slot,
                  const double *l1_reg, const double *l2_reg, const double *tol, const int *max_iter, int n_groups,
                  int warm_start, int do_screening, double *W, long long ldw, double *info, const CUtensorMap *tmaps,
                  double *group_stats, cudaStream_t st) {
    const Layout L = make_layout(M, K, NB, C);
    SGLM_CHECK_ARG(L.total <= 227 * 1024, SGLM_E_UNSUPPORTED,
                   "enet_cd_cluster: C=%d needs %zu bytes of shared memory per CTA (> 227 KB)", C, L.total);
    auto kern = enet_cd_cluster_kernel<M, K, NB, CH, RG, MINB>;
    SGLM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    if (K > 8) SGLM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)n_groups * K);
    cfg.blockDim = dim3((M + NB) * 32);
    cfg.dynamicSmemBytes = L.total;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = K;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    SGLM_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, prob_Q, prob_q, prob_diag, prob_yy, ldq, C, prob_of_group,
                                    model_of_slot, l1_reg, l2_reg, tol, max_iter, warm_start, do_screening, W, ldw,
                                    info, tmaps, group_stats));
    return SGLM_OK;
}


// Arguments of one launch (the C-ABI call, unpacked), so that the per-group-size translation
// units can be compiled in parallel.
struct Args {
    const double *const *prob_Q; const double *const *prob_q; const double *const *prob_diag;
    const double *prob_yy; long long ldq; int C; const int *prob_of_group; const int *model_of_slot;
    const double *l1_reg; const double *l2_reg; const double *tol; const int *max_iter; int n_groups;
    int warm_start; int do_screening; double *W; long long ldw; double *info; const CUtensorMap *tmaps; cudaStream_t st; int variant;
    double *group_stats;
};

template <int M, int K, int NB, int CH, int RG, int MINB>
static int launch_a(const Args &a) {
    return launch<M, K, NB, CH, RG, MINB>(a.prob_Q, a.prob_q, a.prob_diag, a.prob_yy, a.ldq, a.C, a.prob_of_group,
                                    a.model_of_slot, a.l1_reg, a.l2_reg, a.tol, a.max_iter, a.n_groups,
                                    a.warm_start, a.do_screening, a.W, a.ldw, a.info, a.tmaps, a.group_stats, a.st);
}

// Panel shapes (panel warps NB, 16-byte chunks per thread CH, rows per load group RG, min CTAs per SM).
// The chunk loop of the panel is generic, so every shape is valid for every C.  The panel is bound by
// latency x bytes in flight (RG*CH 16-byte loads per thread), so a CTA gets the register budget of a
// whole SM: 8 loads in flight per thread, one thread per 16-byte column chunk where the slice is narrow
// (measured: profiles/r1_cd_cluster.txt).  `variant` (tuning switch SGLM_CDC_VARIANT) forces one shape.
template <int M, int K>
static int launch_sized(const Args &a) {
    const int own2 = ((a.C + 31) / 32 + K - 1) / K * 16;
    switch (a.variant) {
        case 1: return launch_a<M, K, 8, 1, 4, 1>(a);
        case 2: return launch_a<M, K, 8, 2, 2, 2>(a);
        // 16-warp CTA: the M register warps share ONE SM sub-partition (warp ids 0, 4, 8, 12), 12 panel warps own the
        // other three — the register chains no longer queue behind the panel's FP64 FMAs in their scheduler
        case 6: return launch_a<M, K, 12, 1, 4, 1>(a);
        case 7: return launch_a<M, K, 12, 1, 8, 1>(a);
        default: break;
    }
    if (own2 <= 128) return launch_a<M, K, 4, 1, 8, 1>(a);
    if (own2 <= 256) return launch_a<M, K, 8, 1, 8, 1>(a);
    return launch_a<M, K, 8, 2, 4, 1>(a);
}

template <int M>
static int launch_group(const Args &a, int K) {
    switch (K) {
        case 1: return launch_sized<M, 1>(a);
        case 2: return launch_sized<M, 2>(a);
        case 4: return launch_sized<M, 4>(a);
        case 8: return launch_sized<M, 8>(a);
    }
    return fail(SGLM_E_UNSUPPORTED, "enet_cd_cluster: unsupported cluster size %d", K);
}

// Groups of 8 models: 8 register warps + 8 panel warps (512 threads, 128 registers each), one 16-byte column
// chunk per panel thread and pass (the chunk loop covers wider slices), 8 row loads in flight per thread.
template <int M>
static int launch_group_wide(const Args &a, int K) {
    switch (K) {
        case 2: return launch_a<M, 2, 8, 1, 8, 1>(a);
        case 4: return launch_a<M, 4, 8, 1, 8, 1>(a);
        case 8: return launch_a<M, 8, 4, 1, 8, 1>(a);
    }
    return fail(SGLM_E_UNSUPPORTED, "enet_cd_cluster: unsupported cluster size %d for groups of %d", K, M);
}

int launch_m1(const Args &a, int K);
int launch_m2(const Args &a, int K);
int launch_m4(const Args &a, int K);
int launch_m8(const Args &a, int K);

}  // namespace cdc
}  // namespace sglm

