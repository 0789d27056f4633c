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

// ---------------------------------------------------------------------------------------------- e4m3 correction plane
// "fp16mx" operand format (round 2): value ~= hi + lo8 * 2^-(11+EA), with hi = fp16(v) as before and, instead of the fp16 lo
// plane, a CORRECTION plane of the same size that holds two e4m3 bytes per element:
//     lo8 = e4m3((v - hi) * 2^(11+EA))      the residual, 4 significant bits
//     hi8 = e4m3(hi * 2^EA)                 the value itself at 4 significant bits (multiplies the weights' residual)
// Layout inside a pixel row of C channels (2*C bytes, the fp16 lo plane's footprint): per 32-channel group g, 64 bytes =
// [lo8 of channels 32g..32g+31 | hi8 of the same channels].  A 64-"element" (128-byte) TMA box of the plane therefore is
// [lo8 x32 | hi8 x32 | lo8 x32 | hi8 x32] = four K = 32 blocks of an e4m3 x e4m3 (kind::f8f6f4) MMA, and the conv kernel issues, per 32
// channels, 2 fp16 MMAs (hi * w_hi) + 2 e4m3 MMAs (lo8 * w_hi8, hi8 * w_lo8) = 2 tensor-pipe passes per MAC instead of 3.
// Both correction products carry the power of two 2^(11+EA+w_exp); the fp16 WEIGHT plane is stored pre-multiplied by it and the
// conv epilogue divides it out (round 2's first form undid the pre-scales with uniform UE8M0 block scale factors in TMEM).
// `lo_fmt` arguments of the C ABI: 0 = fp16 residual plane, FAR3D_LO_MX(EA) = 64 + EA = this format.
__device__ __forceinline__ bool lo_is_mx(int lo_fmt) { return lo_fmt != 0; }
__device__ __forceinline__ int lo_mx_exp(int lo_fmt) { return lo_fmt - 64; }

// four floats -> four e4m3 bytes (round to nearest even, saturating at +-448), `a` in the lowest byte
__device__ __forceinline__ uint32_t pack_e4m3x4(float a, float b, float c, float d) {
    uint16_t lo, hi;
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(lo) : "f"(b), "f"(a));
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(hi) : "f"(d), "f"(c));
    return (uint32_t)lo | ((uint32_t)hi << 16);
}
// two e4m3 bytes (low 16 bits of x) -> two floats
__device__ __forceinline__ float2 unpack_e4m3x2(uint32_t x) {
    uint32_t h2;
    asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(h2) : "h"((uint16_t)x));
    return __half22float2(*reinterpret_cast<const __half2*>(&h2));
}
// one element: hi (fp16), and the two correction bytes
__device__ __forceinline__ void split_mx(float v, float lo_scale, float hi_scale, __half& hi, float& lo_s, float& hi_s) {
    hi = __float2half_rn(v);
    const float hf = __half2float(hi);
    lo_s = (v - hf) * lo_scale;
    hi_s = hf * hi_scale;
}
// byte address of the lo8 run that holds channel `ch` of a pixel row, given the fp16-plane ELEMENT pointer of that channel
// (callers index both planes alike): the group's 64 bytes start at element 32g, the channel's lo8 byte is at +(ch & 31)
__device__ __forceinline__ unsigned char* mx_lo8_ptr(const void* lo_elem_ptr, int ch) {
    return (unsigned char*)const_cast<void*>(lo_elem_ptr) - (ch & 31);
}

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

}  // namespace far3d
