// Shared helpers for libfar3d_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/far3d_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libfar3d_sm100 is written for sm_100a only"
#endif

namespace far3d {

extern thread_local char g_err[512];
extern std::atomic<int64_t> g_launches;

inline int fail(int code, const char* fmt, const char* a = "", long b = 0, long c = 0) {
    snprintf(g_err, sizeof(g_err), fmt, a, b, c);
    return code;
}

// call after every kernel launch: counts it and maps launch errors to FAR3D_E_CUDA
inline int launched(const char* name) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "%s: %s", name, cudaGetErrorString(e));
        return FAR3D_E_CUDA;
    }
    return FAR3D_OK;
}

#define FAR3D_REQUIRE(cond, msg)                                                        \
    do {                                                                                \
        if (!(cond)) return ::far3d::fail(FAR3D_E_INVALID, "%s: requirement failed: " msg, __func__); \
    } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// split an fp32 value into fp16 hi + fp16 lo: value ~= hi + lo to 2^-22 relative for |v| >= 2^-3 (hi: 11 significant bits,
// lo: up to 11 more; below that the lo plane enters fp16's subnormal range and the absolute error settles at 2^-25).
// Range: |v| must stay below 65504 (fp16 max) - beyond it the planes become inf / NaN, loudly, not a silently clipped value.
// (bf16 planes, the round-1a format, carried 2^-17: measured 4.9e-4 relative on feat_flatten after the 99 + 7 convolutions of
// cfg-2 against 4.6e-6 for exact fp32 - too close to the 1e-3 parity bar; fp16 planes cost the same MMAs.)
__device__ __forceinline__ void split_fp16(float v, __half& hi, __half& lo) {
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

}  // namespace far3d
