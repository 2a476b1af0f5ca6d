// Sufficient statistics  G[s] = Z' diag(w_s) Z,  Z = [X | Y | 1]  (fp64 path).
//
// This is the one dense contraction of the sGLM hot path: every Gaussian fit of the CV
// grid (OLS / Ridge / Lasso / ElasticNet over folds x alpha x l1_ratio) and every scoring
// pass of the reference (sklearn fits at backend/sglm.py:241, fold copies at
// backend/sglm_cv.py:106-110, scores at backend/sglm.py:150-184) is a function of these
// matrices, and the Poisson IRLS step needs the weighted version X'WX.
//
// fp64 tensor path: mma.sync.m8n8k4.f64 (DMMA) with fp64 accumulators — the fp64-exact
// mode of the Gram builder (tcgen05 has no f64 kind).  Data movement: cp.async 16-byte
// copies into a 3-stage shared-memory ring (row stride padded to 132 doubles so that the
// 4 k-rows of a fragment hit disjoint bank groups), 128x128 output tile per CTA held in
// registers (64 doubles / thread), split-K over row chunks with a deterministic
// second-pass reduction (fixed chunk order, mirrored to the lower triangle).
// Fold handling: a set is a row-weight vector (multiplicity of each row in the fold);
// 16-row tiles whose weights are all zero are compacted away, so a 20 % test fold costs
// 20 % of a full pass and X is never copied per fold.
#include <algorithm>

#include "common.cuh"

namespace sglm {

constexpr int SS_BM = 128;
constexpr int SS_BK = 16;
constexpr int SS_STAGES = 3;
constexpr int SS_LDS = SS_BM + 4;
constexpr int SS_THREADS = 256;
constexpr int SS_MAX_SETS = 64;
constexpr int SS_PANEL = SS_BK * SS_LDS;  // doubles per panel per stage

struct SuffstatsParams {
    const double *X; long long ldx;
    const double *Y; long long ldy; int n_y;
    long long T; int C; int n_aug;
    const double *W; long long ldw;
    int n_sets;
    const int *tile_list;   // [n_sets][n_tiles] compacted non-zero-weight tiles (weighted only)
    const int *tile_count;  // [n_sets]
    long long n_tiles;
    int nt, n_pairs;
    double *partial;
    int ks[SS_MAX_SETS];
    int item_base[SS_MAX_SETS + 1];
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Load one 16 x 128 panel of Z (rows t0.., columns c0..) into smem.
template <int VEC>
__device__ __forceinline__ void load_panel(const SuffstatsParams &p, double *dst, long long t0, int c0) {
    const int tid = threadIdx.x;
    if (VEC == 2 && c0 + SS_BM <= p.C) {
        // interior panel: pure 16-byte async copies
#pragma unroll
        for (int r = 0; r < (SS_BK * SS_BM / 2) / SS_THREADS; ++r) {
            int chunk = tid + r * SS_THREADS;
            int row = chunk >> 6, cc = (chunk & 63) * 2;
            long long t = t0 + row;
            bool ok = t < p.T;
            const double *src = ok ? p.X + t * p.ldx + c0 + cc : p.X;
            cp_async16(dst + row * SS_LDS + cc, src, ok ? 16 : 0);
        }
    } else {
        // edge panel (or unaligned source): element-wise, synthesising [Y | 1 | 0...]
#pragma unroll 4
        for (int r = 0; r < (SS_BK * SS_BM) / SS_THREADS; ++r) {
            int e = tid + r * SS_THREADS;
            int row = e >> 7, cc = e & 127;
            long long t = t0 + row;
            int col = c0 + cc;
            double *d = dst + row * SS_LDS + cc;
            if (t < p.T && col < p.C) {
                cp_async8(d, p.X + t * p.ldx + col, 8);
            } else if (t < p.T && col < p.C + p.n_y) {
                cp_async8(d, p.Y + t * p.ldy + (col - p.C), 8);
            } else {
                *d = (t < p.T && col == p.n_aug - 1) ? 1.0 : 0.0;
            }
        }
    }
}

template <int VEC, bool WEIGHTED>
__global__ void __launch_bounds__(SS_THREADS, 1)
suffstats_dmma_kernel(const SuffstatsParams p) {
    extern __shared__ __align__(16) double smem[];
    double *sI = smem;                                   // [STAGES][PANEL]
    double *sJ = smem + SS_STAGES * SS_PANEL;            // [STAGES][PANEL]
    double *sW = smem + 2 * SS_STAGES * SS_PANEL;        // [STAGES][BK]

    // ---- decode work item -> (set, chunk, tile pair)
    int item = blockIdx.x, s = 0;
    while (s + 1 < p.n_sets && item >= p.item_base[s + 1]) ++s;
    int rem = item - p.item_base[s];
    const int chunk = rem / p.n_pairs;
    int pr = rem - chunk * p.n_pairs;
    int ti = 0;
    while (pr >= p.nt - ti) { pr -= p.nt - ti; ++ti; }
    const int tj = ti + pr;
    const bool diag = (ti == tj);

    const long long n_act = WEIGHTED ? (long long)p.tile_count[s] : p.n_tiles;
    const long long lo = (n_act * chunk) / p.ks[s];
    const long long hi = (n_act * (chunk + 1)) / p.ks[s];
    const int n_my = (int)(hi - lo);
    const int *my_list = WEIGHTED ? p.tile_list + (long long)s * p.n_tiles + lo : nullptr;
    const double *Ws = WEIGHTED ? p.W + (long long)s * p.ldw : nullptr;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int g = lane >> 2, kq = lane & 3;

    double acc[8][4][2];
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

    auto issue = [&](int it) {
        if (it < n_my) {
            const int stg = it % SS_STAGES;
            const long long tile = WEIGHTED ? (long long)my_list[it] : lo + it;
            const long long t0 = tile * SS_BK;
            load_panel<VEC>(p, sI + stg * SS_PANEL, t0, ti * SS_BM);
            if (!diag) load_panel<VEC>(p, sJ + stg * SS_PANEL, t0, tj * SS_BM);
            if (WEIGHTED && tid < SS_BK) {
                long long t = t0 + tid;
                bool ok = t < p.T;
                cp_async8(sW + stg * SS_BK + tid, ok ? Ws + t : Ws, ok ? 8 : 0);
            }
        }
        cp_async_commit();
    };

    issue(0);
    issue(1);
    for (int it = 0; it < n_my; ++it) {
        cp_async_wait<1>();
        __syncthreads();           // tile `it` visible to all; stage (it+2)%3 free (consumed at it-1)
        issue(it + 2);
        const int stg = it % SS_STAGES;
        const double *A = sI + stg * SS_PANEL + wm * 64 + g;
        const double *B = (diag ? sI : sJ) + stg * SS_PANEL + wn * 32 + g;
        const double *Wt = sW + stg * SS_BK;
#pragma unroll
        for (int kk = 0; kk < SS_BK / 4; ++kk) {
            const int k = kk * 4 + kq;
            double a[8], b[4];
            const double wv = WEIGHTED ? Wt[k] : 1.0;
#pragma unroll
            for (int mi = 0; mi < 8; ++mi) {
                a[mi] = A[k * SS_LDS + mi * 8];
                if (WEIGHTED) a[mi] *= wv;
            }
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) b[ni] = B[k * SS_LDS + ni * 8];
#pragma unroll
            for (int mi = 0; mi < 8; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma_m8n8k4(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
    }
    cp_async_wait<0>();

    // ---- epilogue: 128x128 partial tile (row-major) for the deterministic reduction
    double *out = p.partial + (long long)item * (SS_BM * SS_BM);
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            int row = wm * 64 + mi * 8 + g;
            int col = wn * 32 + ni * 8 + kq * 2;
            double2 v = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
            *reinterpret_cast<double2 *>(out + row * SS_BM + col) = v;
        }
}

// Sum the K-chunk partials in chunk order and write G (upper tile + mirrored lower tile).
__global__ void __launch_bounds__(256)
suffstats_reduce_kernel(const SuffstatsParams p, double *__restrict__ G, long long ldg) {
    const int per_set = p.n_pairs;
    int sp = blockIdx.y;                 // set * n_pairs + pair
    const int s = sp / per_set;
    int pr = sp - s * per_set;
    const int pair = pr;
    int ti = 0;
    while (pr >= p.nt - ti) { pr -= p.nt - ti; ++ti; }
    const int tj = ti + pr;
    double *Gs = G + (long long)s * p.n_aug * ldg;
    for (int e = blockIdx.x * 256 + threadIdx.x; e < SS_BM * SS_BM; e += gridDim.x * 256) {
        const int r = e >> 7, c = e & 127;
        const int gi = ti * SS_BM + r, gj = tj * SS_BM + c;
        if (gi >= p.n_aug || gj >= p.n_aug) continue;
        if (ti == tj && c < r) continue;            // diagonal tile: upper part only, mirrored below
        double sum = 0.0;
        for (int ch = 0; ch < p.ks[s]; ++ch)
            sum += p.partial[((long long)p.item_base[s] + (long long)ch * per_set + pair) * (SS_BM * SS_BM) + e];
        Gs[(long long)gi * ldg + gj] = sum;
        if (gi != gj) Gs[(long long)gj * ldg + gi] = sum;
    }
}

// flag[s][tile] -> compacted ascending list of tiles whose weights are not all zero.
__global__ void __launch_bounds__(1024)
suffstats_tile_list_kernel(const double *__restrict__ W, long long ldw, long long T, long long n_tiles,
                           int *__restrict__ tile_list, int *__restrict__ tile_count) {
    const int s = blockIdx.x;
    const double *Ws = W + (long long)s * ldw;
    int *list = tile_list + (long long)s * n_tiles;
    __shared__ int warp_tot[32];
    __shared__ int base_sh;
    if (threadIdx.x == 0) base_sh = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long b0 = 0; b0 < n_tiles; b0 += 1024) {
        long long tile = b0 + threadIdx.x;
        int flag = 0;
        if (tile < n_tiles) {
            long long t0 = tile * SS_BK;
#pragma unroll
            for (int k = 0; k < SS_BK; ++k)
                if (t0 + k < T && Ws[t0 + k] != 0.0) flag = 1;
        }
        unsigned bal = __ballot_sync(0xffffffffu, flag);
        int pre = __popc(bal & ((1u << lane) - 1));
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
        for (int w = 0; w < 32; ++w) {
            int v = warp_tot[w];
            if (w < warp) woff += v;
            tot += v;
        }
        const int base = base_sh;
        if (flag) list[base + woff + pre] = (int)tile;
        __syncthreads();
        if (threadIdx.x == 0) base_sh = base + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_count[s] = base_sh;
}

__global__ void __launch_bounds__(256)
index_counts_kernel(const long long *__restrict__ idx, long long n, double *__restrict__ counts, long long T) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        long long t = idx[i];
        if (t < 0) t += T;                      // numpy negative indexing
        if (t >= 0 && t < T) atomicAdd(counts + t, 1.0);
    }
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static int plan(long long T, int C, int n_y, int n_sets, const int *ksplit_host, SuffstatsParams &p,
                size_t &list_bytes, size_t &count_bytes, size_t &partial_bytes) {
    if (n_sets < 1 || n_sets > SS_MAX_SETS) return fail(SGLM_E_SHAPE, "suffstats: n_sets must be in [1,%d]", SS_MAX_SETS);
    if (T < 0 || C < 0 || n_y < 0) return fail(SGLM_E_SHAPE, "suffstats: negative shape");
    p.T = T; p.C = C; p.n_y = n_y; p.n_aug = C + n_y + 1; p.n_sets = n_sets;
    p.n_tiles = ceil_div<long long>(std::max<long long>(T, 1), SS_BK);
    if (p.n_tiles > 0x7fffffffLL) return fail(SGLM_E_SHAPE, "suffstats: too many row tiles");
    p.nt = ceil_div(p.n_aug, SS_BM);
    p.n_pairs = p.nt * (p.nt + 1) / 2;
    long long base = 0;
    const int sms = sm_count();
    for (int s = 0; s < n_sets; ++s) {
        int ks;
        if (ksplit_host) ks = ksplit_host[s];
        else {
            long long want = ceil_div<long long>(4LL * sms, (long long)n_sets * p.n_pairs);
            ks = (int)std::max<long long>(1, std::min<long long>(want, p.n_tiles / 8));
        }
        ks = (int)std::max<long long>(1, std::min<long long>(ks, p.n_tiles));
        p.ks[s] = ks;
        p.item_base[s] = (int)base;
        base += (long long)ks * p.n_pairs;
        if (base > 0x7fffffffLL) return fail(SGLM_E_SHAPE, "suffstats: too many work items");
    }
    p.item_base[n_sets] = (int)base;
    list_bytes = align256((size_t)n_sets * p.n_tiles * sizeof(int));
    count_bytes = align256((size_t)n_sets * sizeof(int));
    partial_bytes = (size_t)base * SS_BM * SS_BM * sizeof(double);
    return SGLM_OK;
}

}  // namespace sglm

using namespace sglm;

extern "C" size_t sglm_suffstats_workspace_bytes(int64_t T, int32_t C, int32_t n_y, int32_t n_sets,
                                                 const int32_t *ksplit_host) {
    SuffstatsParams p;
    size_t a, b, c;
    if (plan(T, C, n_y, n_sets, ksplit_host, p, a, b, c) != SGLM_OK) return 0;
    return a + b + c;
}

extern "C" int sglm_suffstats_f64(const double *X, int64_t ldx, const double *Y, int64_t ldy, int32_t n_y,
                                  int64_t T, int32_t C, const double *W, int64_t ldw, int32_t n_sets,
                                  const int32_t *ksplit_host, double *G, int64_t ldg, void *workspace,
                                  size_t workspace_bytes, void *stream) {
    SuffstatsParams p;
    size_t list_bytes, count_bytes, partial_bytes;
    int rc = plan(T, C, n_y, n_sets, ksplit_host, p, list_bytes, count_bytes, partial_bytes);
    if (rc != SGLM_OK) return rc;
    SGLM_CHECK_ARG(G && workspace, SGLM_E_INVALID_ARG, "suffstats: null output/workspace");
    SGLM_CHECK_ARG(C == 0 || X, SGLM_E_INVALID_ARG, "suffstats: null X");
    SGLM_CHECK_ARG(n_y == 0 || Y, SGLM_E_INVALID_ARG, "suffstats: null Y");
    SGLM_CHECK_ARG(ldx >= C && ldy >= n_y && ldg >= p.n_aug, SGLM_E_SHAPE, "suffstats: leading dimension too small");
    SGLM_CHECK_ARG(W == nullptr || ldw >= T, SGLM_E_SHAPE, "suffstats: ldw < T");
    SGLM_CHECK_ARG(workspace_bytes >= list_bytes + count_bytes + partial_bytes, SGLM_E_WORKSPACE,
                   "suffstats: workspace too small (%zu < %zu)", workspace_bytes,
                   list_bytes + count_bytes + partial_bytes);
    SGLM_CHECK_ARG(((uintptr_t)workspace & 255) == 0, SGLM_E_ALIGN, "suffstats: workspace must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    p.X = X; p.ldx = ldx; p.Y = Y; p.ldy = ldy; p.W = W; p.ldw = ldw;
    char *ws = (char *)workspace;
    p.tile_list = (const int *)ws;
    p.tile_count = (const int *)(ws + list_bytes);
    p.partial = (double *)(ws + list_bytes + count_bytes);

    if (W) {
        suffstats_tile_list_kernel<<<n_sets, 1024, 0, st>>>(W, ldw, T, p.n_tiles, (int *)p.tile_list,
                                                            (int *)p.tile_count);
        SGLM_LAUNCH_OK("suffstats_tile_list_kernel");
    }
    const bool vec2 = (ldx % 2 == 0) && (((uintptr_t)X & 15) == 0);
    const size_t smem = (size_t)(2 * SS_STAGES * SS_PANEL + SS_STAGES * SS_BK) * sizeof(double);
    const int grid = p.item_base[n_sets];
#define SS_LAUNCH(V, WT)                                                                              \
    do {                                                                                              \
        SGLM_CUDA_OK(cudaFuncSetAttribute(suffstats_dmma_kernel<V, WT>,                               \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
        suffstats_dmma_kernel<V, WT><<<grid, SS_THREADS, smem, st>>>(p);                              \
    } while (0)
    if (W) { if (vec2) SS_LAUNCH(2, true); else SS_LAUNCH(1, true); }
    else   { if (vec2) SS_LAUNCH(2, false); else SS_LAUNCH(1, false); }
#undef SS_LAUNCH
    SGLM_LAUNCH_OK("suffstats_dmma_kernel");
    dim3 rgrid(16, n_sets * p.n_pairs);
    suffstats_reduce_kernel<<<rgrid, 256, 0, st>>>(p, G, ldg);
    SGLM_LAUNCH_OK("suffstats_reduce_kernel");
    return SGLM_OK;
}

extern "C" int sglm_index_counts_f64(const int64_t *idx, int64_t n_idx, double *counts, int64_t T, void *stream) {
    SGLM_CHECK_ARG(n_idx >= 0 && T >= 0, SGLM_E_SHAPE, "index_counts: negative size");
    if (n_idx == 0) return SGLM_OK;
    SGLM_CHECK_ARG(idx && counts, SGLM_E_INVALID_ARG, "index_counts: null pointer");
    int grid = (int)std::min<long long>(ceil_div<long long>(n_idx, 256), (long long)sm_count() * 16);
    index_counts_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const long long *)idx, n_idx, counts, T);
    SGLM_LAUNCH_OK("index_counts_kernel");
    return SGLM_OK;
}
