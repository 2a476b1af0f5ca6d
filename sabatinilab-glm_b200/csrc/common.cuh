// Shared helpers for libsglm_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/sglm_b200.h"

namespace sglm {

// Thread-local last error text (sglm_last_error()).
char *last_error_buf();
int fail(int code, const char *fmt, ...);

#define SGLM_CHECK_ARG(cond, code, ...)                    \
    do {                                                   \
        if (!(cond)) return ::sglm::fail((code), __VA_ARGS__); \
    } while (0)

#define SGLM_CUDA_OK(expr)                                                              \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess)                                                          \
            return ::sglm::fail(SGLM_E_CUDA, "%s failed: %s (%s:%d)", #expr,            \
                                cudaGetErrorString(_e), __FILE__, __LINE__);            \
    } while (0)

#define SGLM_LAUNCH_OK(name)                                                            \
    do {                                                                                \
        cudaError_t _e = cudaGetLastError();                                            \
        if (_e != cudaSuccess)                                                          \
            return ::sglm::fail(SGLM_E_CUDA, "launch of %s failed: %s", (name),         \
                                cudaGetErrorString(_e));                                \
    } while (0)

inline int sm_count() {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess)
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
}

template <typename T>
__host__ __device__ inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace sglm
