// Shared helpers for libfar3d_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
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

// split an fp32 value into bf16 hi + bf16 lo (value ~= hi + lo to 2^-17 relative)
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

}  // namespace far3d
