// Measurement probes (bench.py only; not part of the hot path): the read bandwidth the SMs get from L2 for a
// buffer that stays resident there (the regime of the coordinate-descent panel: rows of a 32 MB Gram matrix
// streamed by every CTA), and from HBM for a buffer far larger than L2.  Plain 16-byte loads, 8 in flight per
// thread, every CTA walks the whole buffer with a grid stride — the access pattern of the panel warps.
#include "common.cuh"

namespace sglm {

__global__ void __launch_bounds__(256)
probe_read_kernel(const double2 *__restrict__ buf, long long n2, int repeats, double *__restrict__ sink) {
    double acc = 0.0;
    const long long stride = (long long)gridDim.x * 256;
    for (int r = 0; r < repeats; ++r) {
        long long i = (long long)blockIdx.x * 256 + threadIdx.x;
        for (; i + 7 * stride < n2; i += 8 * stride) {
            double2 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = __ldcg(buf + i + k * stride);      // L2 (no L1 allocation)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc += v[k].x + v[k].y;
        }
        for (; i < n2; i += stride) { const double2 v = __ldcg(buf + i); acc += v.x + v.y; }
    }
    if (acc == 1.2345678e300) sink[0] = acc;      // never true for finite data: keeps the loads alive
}

}  // namespace sglm

using namespace sglm;

// Reads `n_doubles` doubles `repeats` times (bytes moved = 8 * n_doubles * repeats); the caller times the call
// with CUDA events.  ctas_per_sm <= 0: 8.
extern "C" int sglm_probe_read_f64(const double *buf, int64_t n_doubles, int32_t repeats, int32_t ctas_per_sm,
                                   double *sink, void *stream) {
    SGLM_CHECK_ARG(buf && sink && n_doubles >= 2 && repeats >= 1, SGLM_E_INVALID_ARG, "probe_read: bad argument");
    SGLM_CHECK_ARG(((uintptr_t)buf & 15) == 0, SGLM_E_ALIGN, "probe_read: buffer must be 16-byte aligned");
    const int grid = sm_count() * (ctas_per_sm > 0 ? ctas_per_sm : 8);
    probe_read_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const double2 *)buf, n_doubles / 2, repeats, sink);
    SGLM_LAUNCH_OK("probe_read_kernel");
    return SGLM_OK;
}
