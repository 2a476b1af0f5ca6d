// Lag / shift gather: builds the T x C design matrix from the T x P base signals in one
// launch.  Replaces sglm_pp.timeshift / shift / timeshift_multiple / concat_all_shifts
// (reference backend/sglm_pp.py:23-103, :298-357, :436-457), which make >= 4 full-size
// host copies per shift block.
//
// Roofline: HBM.  Algorithmic bytes = 8*T*P (read) + 8*T*C (write); the kernel is a pure
// streaming writer, so the design is: stage the source window (tile rows + lag halo) in
// shared memory with coalesced loads, pre-fill out-of-range rows with the fill pattern so
// the inner loop is branch-free, and emit 16-byte coalesced stores (one warp = 512
// contiguous bytes of an output row).
#include <stdarg.h>

#include "common.cuh"

namespace sglm {

static thread_local char g_err[512] = "";
char *last_error_buf() { return g_err; }
int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

constexpr int TS_THREADS = 256;

// Staged kernel.  smem: off[C] (int) then window[n_src_rows * P] (8-byte words).
// window row r holds source row (t0 - smax + r); rows outside [0,T) hold `fill`.
// out[t0+i, c] = window[(i + smax - shift_c) * P + src_c].
template <int VEC>
__global__ void __launch_bounds__(TS_THREADS)
timeshift_staged_kernel(const uint64_t *__restrict__ X, long long T, int P, long long ldx,
                        const int *__restrict__ col_src, const int *__restrict__ col_shift, int C,
                        uint64_t fill, uint64_t *__restrict__ out, long long ldo, int tile_rows,
                        int smin, int smax) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int *off = reinterpret_cast<int *>(smem_raw);
    const int off_words = (C + 1) / 2 * 2;  // keep the window 8-byte aligned
    uint64_t *win = reinterpret_cast<uint64_t *>(smem_raw) + off_words / 2;

    const long long t0 = (long long)blockIdx.x * tile_rows;
    const int rows = (int)min((long long)tile_rows, T - t0);
    const int n_src_rows = rows + (smax - smin);
    const long long src0 = t0 - smax;

    // column offsets; a shift with |a| >= T can never hit a valid row: point it at a
    // guaranteed fill row by clamping into the window's out-of-range part.
    for (int c = threadIdx.x; c < C; c += TS_THREADS)
        off[c] = (smax - col_shift[c]) * P + col_src[c];

    // stage the source window (coalesced along the row-major source)
    const int n_win = n_src_rows * P;
    if (ldx == P) {
        const long long base = src0 * (long long)P;
        const long long total = T * (long long)P;
        for (int i = threadIdx.x; i < n_win; i += TS_THREADS) {
            long long g = base + i;
            win[i] = (g >= 0 && g < total) ? X[g] : fill;
        }
    } else {
        for (int i = threadIdx.x; i < n_win; i += TS_THREADS) {
            int r = i / P, p = i - r * P;
            long long ts = src0 + r;
            win[i] = (ts >= 0 && ts < T) ? X[ts * ldx + p] : fill;
        }
    }
    __syncthreads();

    if (VEC == 2) {
        const int pairs = C >> 1;
        for (int pc = threadIdx.x; pc < pairs; pc += TS_THREADS) {
            const int o0 = off[2 * pc], o1 = off[2 * pc + 1];
            uint64_t *dst = out + t0 * ldo + 2 * pc;
#pragma unroll 4
            for (int i = 0; i < rows; ++i) {
                ulonglong2 v;
                v.x = win[i * P + o0];
                v.y = win[i * P + o1];
                __stcs(reinterpret_cast<ulonglong2 *>(dst + (long long)i * ldo), v);
            }
        }
    } else {
        for (int c = threadIdx.x; c < C; c += TS_THREADS) {
            const int o0 = off[c];
            uint64_t *dst = out + t0 * ldo + c;
#pragma unroll 4
            for (int i = 0; i < rows; ++i) dst[(long long)i * ldo] = win[i * P + o0];
        }
    }
}

// Direct kernel: general fallback when the source window does not fit shared memory
// (very wide sources or very long lags).  One thread per output element, coalesced
// stores, source reads served by L2.
__global__ void __launch_bounds__(TS_THREADS)
timeshift_direct_kernel(const uint64_t *__restrict__ X, long long T, long long ldx,
                        const int *__restrict__ col_src, const int *__restrict__ col_shift, int C,
                        uint64_t fill, uint64_t *__restrict__ out, long long ldo) {
    const long long total = T * (long long)C;
    for (long long e = (long long)blockIdx.x * TS_THREADS + threadIdx.x; e < total;
         e += (long long)gridDim.x * TS_THREADS) {
        long long t = e / C;
        int c = (int)(e - t * C);
        long long ts = t - col_shift[c];
        out[t * ldo + c] = (ts >= 0 && ts < T) ? X[ts * ldx + col_src[c]] : fill;
    }
}

__global__ void __launch_bounds__(256)
crop_rows_kernel(const double *__restrict__ X, long long ldx, long long row_begin, long long n_rows,
                 int n_cols, double *__restrict__ out, long long ldo) {
    const long long total = n_rows * (long long)n_cols;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total;
         e += (long long)gridDim.x * 256) {
        long long r = e / n_cols;
        int c = (int)(e - r * n_cols);
        out[r * ldo + c] = X[(row_begin + r) * ldx + c];
    }
}

// Gather restricted to listed rows (fused dropna compaction, er_refactored_from_scratch_cleanup.py:427-429:
// the rows pandas' dropna would remove are never built): out[r, c] = X[rows[r] - shift_c, src_c] or fill.
// Consecutive r mostly list consecutive rows, so a warp's source reads stay within a few cache lines (L2 / L1).
__global__ void __launch_bounds__(TS_THREADS)
timeshift_rows_kernel(const uint64_t *__restrict__ X, long long T, long long ldx,
                      const int *__restrict__ col_src, const int *__restrict__ col_shift, int C,
                      uint64_t fill, const long long *__restrict__ rows, long long n_rows,
                      uint64_t *__restrict__ out, long long ldo) {
    const long long total = n_rows * (long long)C;
    for (long long e = (long long)blockIdx.x * TS_THREADS + threadIdx.x; e < total;
         e += (long long)gridDim.x * TS_THREADS) {
        const long long r = e / C;
        const int c = (int)(e - r * C);
        const long long ts = rows[r] - col_shift[c];
        __stcs(out + r * ldo + c, (ts >= 0 && ts < T) ? X[ts * ldx + col_src[c]] : fill);
    }
}

// Rows of the lag design that hold no NaN, decided from the BASE signals and the column map (the design itself
// is not needed): row t is dropped when a source row t - shift_c falls outside [0, T) and the fill is NaN, or when
// X[t - shift_c, src_c] is NaN for some column c.  Pass 1 marks the edge rows, pass 2 scans the T x P base signals
// once and scatters every NaN it finds to the (few) design rows it reaches.
__global__ void __launch_bounds__(256)
lag_valid_init_kernel(unsigned char *__restrict__ valid, long long T, int smin, int smax, int fill_is_nan) {
    for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < T; t += (long long)gridDim.x * 256)
        valid[t] = (!fill_is_nan || (t >= (long long)smax && t < T + (long long)smin)) ? 1 : 0;
}
__global__ void __launch_bounds__(256)
lag_valid_scatter_kernel(const double *__restrict__ X, long long T, int P, long long ldx,
                         const int *__restrict__ col_src, const int *__restrict__ col_shift, int C,
                         unsigned char *__restrict__ valid) {
    const long long total = T * (long long)P;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const long long u = e / P;
        const int p = (int)(e - u * P);
        const double v = X[u * ldx + p];
        if (v == v) continue;
        for (int c = 0; c < C; ++c)
            if (col_src[c] == p) {
                const long long t = u + col_shift[c];
                if (t >= 0 && t < T) valid[t] = 0;
            }
    }
}
// summary of a 0/1 row mask: out3 = {count, first set row, last set row} (first = T, last = -1 when empty)
__global__ void __launch_bounds__(256)
mask_summary_kernel(const unsigned char *__restrict__ valid, long long T, long long *__restrict__ out3) {
    long long cnt = 0, first = T, last = -1;
    for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < T; t += (long long)gridDim.x * 256)
        if (valid[t]) { ++cnt; first = min(first, t); last = max(last, t); }
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
        last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd((unsigned long long *)out3, (unsigned long long)cnt);
        atomicMin(out3 + 1, first);
        atomicMax(out3 + 2, last);
    }
}
// Ordered compaction of the set rows of a mask: block b owns rows [b*CH, (b+1)*CH); pass A counts, a one-block
// exclusive scan of the block counts follows, pass B writes the row indices in ascending order.
constexpr int CMP_CH = 4096;
// match < 0: a row is selected when its byte is non-zero; match >= 0: when its byte equals `match` (cell ids)
__device__ __forceinline__ bool mask_hit(unsigned char v, int match) { return match < 0 ? (v != 0) : ((int)v == match); }

__global__ void __launch_bounds__(256)
mask_block_count_kernel(const unsigned char *__restrict__ valid, long long T, int match, long long *__restrict__ counts) {
    const long long t0 = (long long)blockIdx.x * CMP_CH;
    int c = 0;
    for (int i = threadIdx.x; i < CMP_CH; i += 256) c += (t0 + i < T && mask_hit(valid[t0 + i], match)) ? 1 : 0;
    __shared__ int sh[8];
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { int s = 0; for (int w = 0; w < 8; ++w) s += sh[w]; counts[blockIdx.x] = s; }
}
__global__ void __launch_bounds__(1024)
exclusive_scan_kernel(long long *__restrict__ counts, long long n) {
    // one block; n = number of 4096-row blocks (T = 2M: 489)
    __shared__ long long carry;
    __shared__ long long buf[1024];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (long long base = 0; base < n; base += 1024) {
        const long long i = base + threadIdx.x;
        const long long v = i < n ? counts[i] : 0;
        buf[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const long long add = threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
            __syncthreads();
            buf[threadIdx.x] += add;
            __syncthreads();
        }
        if (i < n) counts[i] = carry + buf[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += buf[1023];
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256)
mask_compact_kernel(const unsigned char *__restrict__ valid, long long T, int match, const long long *__restrict__ offs,
                    long long *__restrict__ rows) {
    // a warp owns 512 consecutive rows of the block's 4096 (ballot-ordered writes); warp prefix via shared memory
    const long long t0 = (long long)blockIdx.x * CMP_CH;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ int wcnt[8];
    unsigned m[16];
    int c = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const long long t = t0 + warp * 512 + k * 32 + lane;
        m[k] = __ballot_sync(0xffffffffu, t < T && mask_hit(valid[t], match));
        c += __popc(m[k]);
    }
    if (lane == 0) wcnt[warp] = c;
    __syncthreads();
    long long pos = offs[blockIdx.x];
    for (int w = 0; w < warp; ++w) pos += wcnt[w];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        if ((m[k] >> lane) & 1u) rows[pos + __popc(m[k] & ((1u << lane) - 1u))] = t0 + warp * 512 + k * 32 + lane;
        pos += __popc(m[k]);
    }
}

// Row mask of an index list (positions in [0, T)): mask[t] = 1 for listed rows; *dup is set when a row is listed
// twice.  Bytes are set through 32-bit atomicOr on the containing word (mask must be 4-byte aligned, zeroed).
__global__ void __launch_bounds__(256)
index_mask_kernel(const long long *__restrict__ idx, long long n, unsigned *__restrict__ mask_words, long long T,
                  int *__restrict__ dup) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const long long t = idx[i];
        if (t < 0 || t >= T) continue;
        const unsigned bit = 1u << ((t & 3) * 8);
        if (atomicOr(mask_words + (t >> 2), bit) & bit) *dup = 1;
    }
}
// Row signatures of overlapping row sets: sig[t] has bit `bit` set when row t is listed in set `bit` (<= 62 sets).
__global__ void __launch_bounds__(256)
rows_or_bit_kernel(const long long *__restrict__ idx, long long n, unsigned long long bitmask,
                   unsigned long long *__restrict__ sig, long long T) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const long long t = idx[i];
        if (t >= 0 && t < T) atomicOr(sig + t, bitmask);
    }
}
// Distinct signatures -> cells: a 256-slot open-addressing table (keys = signatures, ~0 = empty); cell_of[t] = slot of
// row t's signature, counts[slot] = rows of the cell.  *overflow is set when more than 200 distinct signatures exist.
__global__ void __launch_bounds__(256)
cells_from_signatures_kernel(const unsigned long long *__restrict__ sig, long long T, unsigned char *__restrict__ cell_of,
                             unsigned long long *__restrict__ keys, long long *__restrict__ counts, int *__restrict__ used,
                             int *__restrict__ overflow) {
    for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < T; t += (long long)gridDim.x * 256) {
        const unsigned long long k = sig[t];
        unsigned slot = (unsigned)((k * 0x9E3779B97F4A7C15ULL) >> 56);
        int found = -1;
        for (int probe = 0; probe < 256; ++probe, slot = (slot + 1) & 255u) {
            unsigned long long cur = keys[slot];
            if (cur == ~0ULL) {
                if (atomicAdd(used, 0) >= 200) break;
                cur = atomicCAS(keys + slot, ~0ULL, k);
                if (cur == ~0ULL) { atomicAdd(used, 1); cur = k; }
            }
            if (cur == k) { found = (int)slot; break; }
        }
        if (found < 0) { *overflow = 1; cell_of[t] = 255; continue; }
        cell_of[t] = (unsigned char)found;
        atomicAdd((unsigned long long *)(counts + found), 1ULL);
    }
}

// np.roll(y, shift): out[(i + shift) mod n] = y[i]   (backend/sglm_cv.py:95-96)
__global__ void __launch_bounds__(256)
roll_kernel(const double *__restrict__ y, long long n, long long shift, double *__restrict__ out) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        long long j = i + shift;
        if (j >= n) j -= n;
        out[j] = y[i];
    }
}

__global__ void minmax_shift_kernel(const int *__restrict__ col_shift, const int *__restrict__ col_src,
                                    int C, int n_cols_in, int *out4) {
    // out4 = {min shift, max shift, bad source flag, unused}
    int mn = INT_MAX, mx = INT_MIN, bad = 0;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        mn = min(mn, col_shift[c]);
        mx = max(mx, col_shift[c]);
        bad |= (col_src[c] < 0 || col_src[c] >= n_cols_in);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    __shared__ int s[3][32];
    if ((threadIdx.x & 31) == 0) {
        s[0][threadIdx.x >> 5] = mn; s[1][threadIdx.x >> 5] = mx; s[2][threadIdx.x >> 5] = bad;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
            mn = min(mn, s[0][w]); mx = max(mx, s[1][w]); bad |= s[2][w];
        }
        out4[0] = mn; out4[1] = mx; out4[2] = bad; out4[3] = 0;
    }
}

}  // namespace sglm

using namespace sglm;

extern "C" int sglm_version(void) { return 100; }
extern "C" const char *sglm_last_error(void) { return last_error_buf(); }

// Hot-path entry: the caller (which built the column map) passes the shift range, so
// nothing is read back from the device and the call never synchronises.
extern "C" int sglm_timeshift_f64_ranged(const double *X, int64_t T, int32_t n_cols_in, int64_t ldx,
                                         const int32_t *col_src, const int32_t *col_shift,
                                         int32_t n_cols_out, int32_t shift_min, int32_t shift_max,
                                         uint64_t fill_bits, double *out, int64_t ldo,
                                         void *stream) {
    SGLM_CHECK_ARG(T >= 0 && n_cols_in > 0 && n_cols_out >= 0, SGLM_E_SHAPE,
                   "timeshift: bad shape T=%lld P=%d C=%d", (long long)T, n_cols_in, n_cols_out);
    SGLM_CHECK_ARG(ldx >= n_cols_in && ldo >= n_cols_out, SGLM_E_SHAPE,
                   "timeshift: leading dimension smaller than row (ldx=%lld ldo=%lld)",
                   (long long)ldx, (long long)ldo);
    if (T == 0 || n_cols_out == 0) return SGLM_OK;
    SGLM_CHECK_ARG(X && col_src && col_shift && out, SGLM_E_INVALID_ARG, "timeshift: null pointer");
    SGLM_CHECK_ARG(shift_min <= shift_max, SGLM_E_INVALID_ARG, "timeshift: empty shift range");
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t *Xw = reinterpret_cast<const uint64_t *>(X);
    uint64_t *ow = reinterpret_cast<uint64_t *>(out);

    // clamp the halo: any |shift| >= T only ever produces fill, and the window formula
    // stays valid for a clamped range as long as those columns are routed to fill rows —
    // which would need a second map; use the direct kernel for that rare case instead.
    const long long span = (long long)shift_max - (long long)shift_min;
    bool oversize = (shift_max >= T || -(long long)shift_min >= T);
    const size_t smem_limit = 160 * 1024;
    const size_t off_bytes = (size_t)((n_cols_out + 1) / 2 * 2) * sizeof(int);
    int tile_rows = 0;
    if (!oversize) {
        for (int tr : {64, 32, 16, 8}) {
            size_t need = off_bytes + (size_t)(tr + span) * n_cols_in * 8;
            if (need <= smem_limit) { tile_rows = tr; break; }
        }
    }
    if (tile_rows == 0) {
        long long total = T * (long long)n_cols_out;
        int grid = (int)std::min<long long>(ceil_div<long long>(total, TS_THREADS), (long long)sm_count() * 32);
        timeshift_direct_kernel<<<grid, TS_THREADS, 0, st>>>(Xw, T, ldx, col_src, col_shift,
                                                            n_cols_out, fill_bits, ow, ldo);
        SGLM_LAUNCH_OK("timeshift_direct_kernel");
        return SGLM_OK;
    }
    const size_t smem = off_bytes + (size_t)(tile_rows + span) * n_cols_in * 8;
    const long long n_tiles = ceil_div<long long>(T, tile_rows);
    SGLM_CHECK_ARG(n_tiles <= 0x7fffffffLL, SGLM_E_SHAPE, "timeshift: too many row tiles");
    const bool vec2 = (n_cols_out % 2 == 0) && (ldo % 2 == 0) && ((uintptr_t)out % 16 == 0);
    if (vec2) {
        SGLM_CUDA_OK(cudaFuncSetAttribute(timeshift_staged_kernel<2>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_limit));
        timeshift_staged_kernel<2><<<(unsigned)n_tiles, TS_THREADS, smem, st>>>(
            Xw, T, n_cols_in, ldx, col_src, col_shift, n_cols_out, fill_bits, ow, ldo, tile_rows,
            shift_min, shift_max);
    } else {
        SGLM_CUDA_OK(cudaFuncSetAttribute(timeshift_staged_kernel<1>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_limit));
        timeshift_staged_kernel<1><<<(unsigned)n_tiles, TS_THREADS, smem, st>>>(
            Xw, T, n_cols_in, ldx, col_src, col_shift, n_cols_out, fill_bits, ow, ldo, tile_rows,
            shift_min, shift_max);
    }
    SGLM_LAUNCH_OK("timeshift_staged_kernel");
    return SGLM_OK;
}

extern "C" int sglm_timeshift_f64(const double *X, int64_t T, int32_t n_cols_in, int64_t ldx,
                                  const int32_t *col_src, const int32_t *col_shift,
                                  int32_t n_cols_out, uint64_t fill_bits, double *out, int64_t ldo,
                                  void *stream) {
    // Generic entry: recover the shift range from the device-resident map (one tiny
    // kernel + 16-byte copy; synchronises the stream).  The Python mirror, which built
    // the map, calls sglm_timeshift_f64_ranged directly and never synchronises.
    if (T == 0 || n_cols_out <= 0) return n_cols_out < 0 ? fail(SGLM_E_SHAPE, "timeshift: C<0") : SGLM_OK;
    SGLM_CHECK_ARG(col_src && col_shift, SGLM_E_INVALID_ARG, "timeshift: null column map");
    cudaStream_t st = (cudaStream_t)stream;
    int *d4 = nullptr;
    int h4[4];
    SGLM_CUDA_OK(cudaMallocAsync((void **)&d4, 4 * sizeof(int), st));
    minmax_shift_kernel<<<1, 256, 0, st>>>(col_shift, col_src, n_cols_out, n_cols_in, d4);
    SGLM_LAUNCH_OK("minmax_shift_kernel");
    SGLM_CUDA_OK(cudaMemcpyAsync(h4, d4, sizeof(h4), cudaMemcpyDeviceToHost, st));
    SGLM_CUDA_OK(cudaStreamSynchronize(st));
    SGLM_CUDA_OK(cudaFreeAsync(d4, st));
    SGLM_CHECK_ARG(h4[2] == 0, SGLM_E_INVALID_ARG, "timeshift: col_src out of range");
    return sglm_timeshift_f64_ranged(X, T, n_cols_in, ldx, col_src, col_shift, n_cols_out, h4[0],
                                     h4[1], fill_bits, out, ldo, stream);
}

extern "C" int sglm_timeshift_rows_f64(const double *X, int64_t T, int32_t n_cols_in, int64_t ldx,
                                      const int32_t *col_src, const int32_t *col_shift, int32_t n_cols_out,
                                      uint64_t fill_bits, const int64_t *rows, int64_t n_rows, double *out,
                                      int64_t ldo, void *stream) {
    SGLM_CHECK_ARG(T >= 0 && n_cols_in > 0 && n_cols_out >= 0 && n_rows >= 0 && ldx >= n_cols_in && ldo >= n_cols_out,
                   SGLM_E_SHAPE, "timeshift_rows: bad shape");
    if (n_rows == 0 || n_cols_out == 0) return SGLM_OK;
    SGLM_CHECK_ARG(X && col_src && col_shift && rows && out, SGLM_E_INVALID_ARG, "timeshift_rows: null pointer");
    const long long total = n_rows * (long long)n_cols_out;
    const int grid = (int)std::min<long long>(ceil_div<long long>(total, TS_THREADS), (long long)sm_count() * 32);
    timeshift_rows_kernel<<<grid, TS_THREADS, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint64_t *>(X), T, ldx, col_src, col_shift, n_cols_out, fill_bits,
        (const long long *)rows, n_rows, reinterpret_cast<uint64_t *>(out), ldo);
    SGLM_LAUNCH_OK("timeshift_rows_kernel");
    return SGLM_OK;
}

extern "C" int sglm_lag_valid_rows(const double *X, int64_t T, int32_t n_cols_in, int64_t ldx,
                                   const int32_t *col_src, const int32_t *col_shift, int32_t n_cols_out,
                                   int32_t shift_min, int32_t shift_max, int32_t fill_is_nan, uint8_t *valid,
                                   int64_t *summary3, void *stream) {
    SGLM_CHECK_ARG(T >= 0 && n_cols_in > 0 && n_cols_out >= 0 && ldx >= n_cols_in, SGLM_E_SHAPE, "lag_valid_rows: bad shape");
    SGLM_CHECK_ARG(valid && summary3 && (n_cols_out == 0 || (col_src && col_shift)), SGLM_E_INVALID_ARG,
                   "lag_valid_rows: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const long long init[3] = {0, (long long)T, -1};
    SGLM_CUDA_OK(cudaMemcpyAsync(summary3, init, sizeof(init), cudaMemcpyHostToDevice, st));
    SGLM_CUDA_OK(cudaStreamSynchronize(st));       // `init` lives on this frame
    if (T == 0) return SGLM_OK;
    const int grid = (int)std::min<long long>(ceil_div<long long>(T, 256), (long long)sm_count() * 16);
    if (n_cols_out == 0) { shift_min = 0; shift_max = 0; }
    lag_valid_init_kernel<<<grid, 256, 0, st>>>(valid, T, shift_min, shift_max, fill_is_nan);
    SGLM_LAUNCH_OK("lag_valid_init_kernel");
    if (n_cols_out > 0) {
        SGLM_CHECK_ARG(X != nullptr, SGLM_E_INVALID_ARG, "lag_valid_rows: null X");
        const long long total = T * (long long)n_cols_in;
        const int g2 = (int)std::min<long long>(ceil_div<long long>(total, 256), (long long)sm_count() * 32);
        lag_valid_scatter_kernel<<<g2, 256, 0, st>>>(X, T, n_cols_in, ldx, col_src, col_shift, n_cols_out, valid);
        SGLM_LAUNCH_OK("lag_valid_scatter_kernel");
    }
    mask_summary_kernel<<<grid, 256, 0, st>>>(valid, T, (long long *)summary3);
    SGLM_LAUNCH_OK("mask_summary_kernel");
    return SGLM_OK;
}

extern "C" int sglm_index_mask_u8(const int64_t *idx, int64_t n_idx, uint8_t *mask, int64_t T, int32_t *dup_flag,
                                  void *stream) {
    SGLM_CHECK_ARG(n_idx >= 0 && T >= 0, SGLM_E_SHAPE, "index_mask: negative size");
    SGLM_CHECK_ARG(mask && dup_flag && (n_idx == 0 || idx), SGLM_E_INVALID_ARG, "index_mask: null pointer");
    SGLM_CHECK_ARG(((uintptr_t)mask & 3) == 0, SGLM_E_ALIGN, "index_mask: mask must be 4-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    SGLM_CUDA_OK(cudaMemsetAsync(mask, 0, (size_t)((T + 3) / 4 * 4), st));     // mask holds T rounded up to 4 bytes
    SGLM_CUDA_OK(cudaMemsetAsync(dup_flag, 0, sizeof(int), st));
    if (n_idx == 0) return SGLM_OK;
    const int grid = (int)std::min<long long>(ceil_div<long long>(n_idx, 256), (long long)sm_count() * 16);
    index_mask_kernel<<<grid, 256, 0, st>>>((const long long *)idx, n_idx, (unsigned *)mask, T, dup_flag);
    SGLM_LAUNCH_OK("index_mask_kernel");
    return SGLM_OK;
}

extern "C" int sglm_rows_or_bit_u64(const int64_t *idx, int64_t n_idx, int32_t bit, uint64_t *sig, int64_t T, void *stream) {
    SGLM_CHECK_ARG(n_idx >= 0 && T >= 0 && bit >= 0 && bit < 63, SGLM_E_SHAPE, "rows_or_bit: bad argument");
    if (n_idx == 0) return SGLM_OK;
    SGLM_CHECK_ARG(idx && sig, SGLM_E_INVALID_ARG, "rows_or_bit: null pointer");
    const int grid = (int)std::min<long long>(ceil_div<long long>(n_idx, 256), (long long)sm_count() * 16);
    rows_or_bit_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const long long *)idx, n_idx, 1ULL << bit,
                                                              (unsigned long long *)sig, T);
    SGLM_LAUNCH_OK("rows_or_bit_kernel");
    return SGLM_OK;
}

// table: 256 keys (uint64) then 256 counts (int64) then {used, overflow} (2 int32) = 4104 bytes, device memory
extern "C" int sglm_cells_from_signatures(const uint64_t *sig, int64_t T, uint8_t *cell_of, void *table, void *stream) {
    SGLM_CHECK_ARG(T >= 0, SGLM_E_SHAPE, "cells_from_signatures: bad shape");
    SGLM_CHECK_ARG(sig && cell_of && table, SGLM_E_INVALID_ARG, "cells_from_signatures: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long *keys = (unsigned long long *)table;
    long long *counts = (long long *)(keys + 256);
    int *used = (int *)(counts + 256);
    SGLM_CUDA_OK(cudaMemsetAsync(keys, 0xff, 256 * sizeof(unsigned long long), st));
    SGLM_CUDA_OK(cudaMemsetAsync(counts, 0, 256 * sizeof(long long) + 2 * sizeof(int), st));
    if (T == 0) return SGLM_OK;
    const int grid = (int)std::min<long long>(ceil_div<long long>(T, 256), (long long)sm_count() * 16);
    cells_from_signatures_kernel<<<grid, 256, 0, st>>>((const unsigned long long *)sig, T, cell_of, keys, counts, used, used + 1);
    SGLM_LAUNCH_OK("cells_from_signatures_kernel");
    return SGLM_OK;
}

extern "C" int sglm_roll_f64(const double *y, int64_t n, int64_t shift, double *out, void *stream) {
    SGLM_CHECK_ARG(n >= 0, SGLM_E_SHAPE, "roll: negative size");
    if (n == 0) return SGLM_OK;
    SGLM_CHECK_ARG(y && out && y != out, SGLM_E_INVALID_ARG, "roll: null or aliased pointer");
    long long s = shift % n;
    if (s < 0) s += n;
    const int grid = (int)std::min<long long>(ceil_div<long long>(n, 256), (long long)sm_count() * 16);
    roll_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(y, n, s, out);
    SGLM_LAUNCH_OK("roll_kernel");
    return SGLM_OK;
}

extern "C" size_t sglm_mask_compact_workspace_bytes(int64_t T) {
    return (size_t)std::max<long long>(1, ceil_div<long long>(T, CMP_CH)) * sizeof(long long);
}

static int mask_compact(const uint8_t *valid, int64_t T, int match, int64_t *rows, void *workspace,
                        size_t workspace_bytes, void *stream);

extern "C" int sglm_mask_compact_rows(const uint8_t *valid, int64_t T, int64_t *rows, void *workspace,
                                      size_t workspace_bytes, void *stream) {
    return mask_compact(valid, T, -1, rows, workspace, workspace_bytes, stream);
}

// rows whose byte equals `match` (the cell ids written by sglm_cells_from_signatures), ascending
extern "C" int sglm_match_compact_rows(const uint8_t *ids, int64_t T, int32_t match, int64_t *rows, void *workspace,
                                       size_t workspace_bytes, void *stream) {
    SGLM_CHECK_ARG(match >= 0 && match < 256, SGLM_E_INVALID_ARG, "match_compact: match must be a byte value");
    return mask_compact(ids, T, match, rows, workspace, workspace_bytes, stream);
}

static int mask_compact(const uint8_t *valid, int64_t T, int match, int64_t *rows, void *workspace,
                        size_t workspace_bytes, void *stream) {
    SGLM_CHECK_ARG(T >= 0, SGLM_E_SHAPE, "mask_compact: bad shape");
    if (T == 0) return SGLM_OK;
    SGLM_CHECK_ARG(valid && rows && workspace, SGLM_E_INVALID_ARG, "mask_compact: null pointer");
    const long long nb = ceil_div<long long>(T, CMP_CH);
    SGLM_CHECK_ARG(workspace_bytes >= (size_t)nb * sizeof(long long), SGLM_E_WORKSPACE, "mask_compact: workspace too small");
    SGLM_CHECK_ARG(nb <= 0x7fffffffLL, SGLM_E_SHAPE, "mask_compact: too many rows");
    cudaStream_t st = (cudaStream_t)stream;
    long long *offs = (long long *)workspace;
    mask_block_count_kernel<<<(unsigned)nb, 256, 0, st>>>(valid, T, match, offs);
    SGLM_LAUNCH_OK("mask_block_count_kernel");
    exclusive_scan_kernel<<<1, 1024, 0, st>>>(offs, nb);
    SGLM_LAUNCH_OK("exclusive_scan_kernel");
    mask_compact_kernel<<<(unsigned)nb, 256, 0, st>>>(valid, T, match, offs, (long long *)rows);
    SGLM_LAUNCH_OK("mask_compact_kernel");
    return SGLM_OK;
}

extern "C" int sglm_crop_rows_f64(const double *X, int64_t ldx, int64_t row_begin, int64_t n_rows,
                                  int32_t n_cols, double *out, int64_t ldo, void *stream) {
    SGLM_CHECK_ARG(n_rows >= 0 && n_cols >= 0 && row_begin >= 0, SGLM_E_SHAPE, "crop_rows: bad shape");
    if (n_rows == 0 || n_cols == 0) return SGLM_OK;
    SGLM_CHECK_ARG(X && out, SGLM_E_INVALID_ARG, "crop_rows: null pointer");
    long long total = n_rows * (long long)n_cols;
    int grid = (int)std::min<long long>(ceil_div<long long>(total, 256), (long long)sm_count() * 32);
    crop_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(X, ldx, row_begin, n_rows, n_cols, out, ldo);
    SGLM_LAUNCH_OK("crop_rows_kernel");
    return SGLM_OK;
}
