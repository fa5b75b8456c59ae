// Shared helpers for the edgegan_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/edgegan_b200.h"

// every kernel launch of the library is followed by this check; it also feeds eg_kernel_launches()
extern unsigned long long g_eg_kernel_launches;
#define EG_CHECK_LAUNCH()                                                       \
    do {                                                                        \
        ++g_eg_kernel_launches;                                                 \
        cudaError_t e__ = cudaGetLastError();                                   \
        if (e__ != cudaSuccess) return eg_fail(e__, __FILE__, __LINE__);        \
    } while (0)

#define EG_REQUIRE(cond)                                                        \
    do {                                                                        \
        if (!(cond)) return eg_fail_arg(#cond, __FILE__, __LINE__);             \
    } while (0)

int eg_fail(cudaError_t e, const char* file, int line);
int eg_fail_arg(const char* what, const char* file, int line);

// fused conv epilogue (include/edgegan_b200.h: EG_EPI_*): y = act(v) or y = v * act'(mask[same index])
struct EgEpi { int mode; int act; const float* mask; };

static inline int eg_ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of `v`; result valid in every thread.  `red` must hold >= 33 floats.
__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    if (wid == 0) {
        float t = lane < nw ? red[lane] : 0.f;
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

// Activation codes shared with the host side (include/edgegan_b200.h).
__device__ __forceinline__ float act_fwd(int act, float x) {
    switch (act) {
        case EG_ACT_RELU: return x > 0.f ? x : 0.f;
        case EG_ACT_LRELU_BLOCK: return x >= 0.f ? x : 0.2f * x;   // tf.maximum(x, 0.2x)
        case EG_ACT_TANH: return tanhf(x);
        case EG_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
        case EG_ACT_LRELU: return x > 0.f ? x : 0.2f * x;            // tf.maximum(0.2x, x)
        default: return x;
    }
}
// derivative w.r.t. the pre-activation value `x` (tie rules: SURVEY A8)
__device__ __forceinline__ float act_grad(int act, float x) {
    switch (act) {
        case EG_ACT_RELU: return x > 0.f ? 1.f : 0.f;
        case EG_ACT_LRELU_BLOCK: return x >= 0.f ? 1.f : 0.2f;
        case EG_ACT_TANH: { float t = tanhf(x); return 1.f - t * t; }
        case EG_ACT_SIGMOID: { float s = 1.f / (1.f + expf(-x)); return s * (1.f - s); }
        case EG_ACT_LRELU: return x > 0.f ? 1.f : 0.2f;              // tie at 0 -> first argument (0.2x)
        default: return 1.f;
    }
}
