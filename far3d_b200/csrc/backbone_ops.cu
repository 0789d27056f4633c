// Backbone / neck glue kernels (HBM-bound, NHWC): stem conv, max-pool, eSE, FPN top-down add, fp16 split/merge,
// and the fp32 SIMT implicit-GEMM convolution used as the exact-fp32 anchor for the tensor-core path.
// Reference: models/backbones/vovnet.py (stem :308-311, pooling :249, eSE :164-185), mmdet FPN.forward.
#include "common.cuh"

namespace far3d {

typedef __half fp16;

// `ch`: channel index of the element inside its pixel row (only the e4m3 correction-plane format needs it, common.cuh)
__device__ __forceinline__ void store_lo_scalar(float v, fp16 h, fp16* lo_elem, int lo_fmt, int ch) {
    if (!lo_is_mx(lo_fmt)) { *lo_elem = __float2half_rn(v - __half2float(h)); return; }
    const float hf = __half2float(h);
    const uint32_t b = pack_e4m3x4((v - hf) * exp2f((float)(11 + lo_mx_exp(lo_fmt))), hf * exp2f((float)lo_mx_exp(lo_fmt)), 0.f, 0.f);
    unsigned char* q = mx_lo8_ptr(lo_elem, ch);
    q[0] = (unsigned char)(b & 0xffu);
    q[32] = (unsigned char)((b >> 8) & 0xffu);
}
__device__ __forceinline__ float load_lo_scalar(const fp16* lo_elem, int lo_fmt, int ch) {
    if (!lo_is_mx(lo_fmt)) return __half2float(*lo_elem);
    return unpack_e4m3x2((uint32_t)*mx_lo8_ptr(lo_elem, ch)).x * exp2f(-(float)(11 + lo_mx_exp(lo_fmt)));
}
__device__ __forceinline__ void store_outputs(float v, size_t fidx, size_t bidx, float* y_f32, fp16* y_hi, fp16* y_lo,
                                              int lo_fmt = 0, int ch = 0) {
    if (y_f32) y_f32[fidx] = v;
    if (y_hi) {
        const fp16 h = __float2half_rn(v);
        y_hi[bidx] = h;
        if (y_lo) store_lo_scalar(v, h, y_lo + bidx, lo_fmt, ch);
    }
}

// ------------------------------------------------------------------------------------------ stem conv 1
// NCHW fp32 image -> NHWC, 3x3 s2 p1, Cin = 3.  One thread per (pixel, 4 output channels); weights in smem.
__global__ void __launch_bounds__(256)
stem_conv_kernel(const float* __restrict__ img, int N, int H, int W, const float* __restrict__ w,
                 const float* __restrict__ bias, int Cout, float* __restrict__ y_f32, fp16* __restrict__ y_hi,
                 fp16* __restrict__ y_lo, int lo_fmt) {
    extern __shared__ float sw[];     // [27][Cout] transposed + bias[Cout]
    for (int i = threadIdx.x; i < Cout * 27; i += blockDim.x) {
        int co = i / 27, t = i % 27;           // w layout (Cout, ky, kx, cin) -> t = (ky*3+kx)*3+ci
        sw[t * Cout + co] = w[i];
    }
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[27 * Cout + i] = bias ? bias[i] : 0.f;
    __syncthreads();
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    const int cq = Cout / 4;
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)N * Ho * Wo * cq) return;
    int c4 = (int)(idx % cq); long r = idx / cq;
    int ow = (int)(r % Wo); r /= Wo;
    int oh = (int)(r % Ho); int n = (int)(r / Ho);
    float acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = sw[27 * Cout + c4 * 4 + j];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        int ih = oh * 2 + ky - 1;
        if (ih < 0 || ih >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            int iw = ow * 2 + kx - 1;
            if (iw < 0 || iw >= W) continue;
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
                float xv = __ldg(img + (((size_t)n * 3 + ci) * H + ih) * W + iw);
                const float* wp = sw + ((ky * 3 + kx) * 3 + ci) * Cout + c4 * 4;
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[j] = fmaf(xv, wp[j], acc[j]);
            }
        }
    }
    size_t o = (((size_t)n * Ho + oh) * Wo + ow) * Cout + c4 * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) store_outputs(fmaxf(acc[j], 0.f), o + j, o + j, y_f32, y_hi, y_lo, lo_fmt, c4 * 4 + j);
}

// ------------------------------------------------------------------------------------------ max-pool 3x3 s2 ceil
template <bool F16>
__global__ void maxpool_kernel(const void* __restrict__ x_hi, const void* __restrict__ x_lo, int N, int H, int W, int C,
                               int x_cs, int x_co, void* __restrict__ y_hi, void* __restrict__ y_lo, int y_cs, int y_co,
                               int Ho, int Wo, int lo_fmt) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)N * Ho * Wo * C) return;
    int c = (int)(idx % C); long r = idx / C;
    int ow = (int)(r % Wo); r /= Wo;
    int oh = (int)(r % Ho); int n = (int)(r / Ho);
    float m = -INFINITY;
    for (int ky = 0; ky < 3; ++ky) {
        int ih = oh * 2 + ky;
        if (ih >= H) continue;
        for (int kx = 0; kx < 3; ++kx) {
            int iw = ow * 2 + kx;
            if (iw >= W) continue;
            size_t i = (((size_t)n * H + ih) * W + iw) * x_cs + x_co + c;
            float v;
            if (F16) {
                v = __half2float(((const fp16*)x_hi)[i]);
                if (x_lo) v += load_lo_scalar((const fp16*)x_lo + i, lo_fmt, x_co + c);
            } else v = ((const float*)x_hi)[i];
            m = fmaxf(m, v);
        }
    }
    size_t o = (((size_t)n * Ho + oh) * Wo + ow) * y_cs + y_co + c;
    if (F16) {
        const fp16 h = __float2half_rn(m);
        ((fp16*)y_hi)[o] = h;
        if (y_lo) store_lo_scalar(m, h, (fp16*)y_lo + o, lo_fmt, y_co + c);
    } else ((float*)y_hi)[o] = m;
}

// ------------------------------------------------------------------------------------------ eSE
// global average pool over HW of fp32 NHWC: grid (N, ceil(C/32)), block (32, 8): 8 row-strips reduced in smem.
__global__ void global_avgpool_kernel(const float* __restrict__ x, float* __restrict__ mean, int N, int HW, int C) {
    __shared__ float part[8][33];
    int n = blockIdx.x, c = blockIdx.y * 32 + threadIdx.x;
    float s = 0.f;
    if (c < C)
        for (int p = threadIdx.y; p < HW; p += 8) s += x[((size_t)n * HW + p) * C + c];
    part[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x];
        mean[(size_t)n * C + c] = t / (float)HW;
    }
}

// faster two-stage pooling: stage 1 partial sums over row chunks. grid (N, chunks, ceil(C/128)) block 128 (float per thread)
__global__ void avgpool_partial_kernel(const float* __restrict__ x, float* __restrict__ part, int N, int HW, int C,
                                       int chunks) {
    int n = blockIdx.x, ch = blockIdx.y, c = blockIdx.z * blockDim.x + threadIdx.x;
    if (c >= C) return;
    int per = (HW + chunks - 1) / chunks;
    int p0 = ch * per, p1 = min(HW, p0 + per);
    float s = 0.f;
    for (int p = p0; p < p1; ++p) s += x[((size_t)n * HW + p) * C + c];
    part[((size_t)n * chunks + ch) * C + c] = s;
}
__global__ void avgpool_final_kernel(const float* __restrict__ part, float* __restrict__ mean, int N, int HW, int C,
                                     int chunks) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * C) return;
    int n = idx / C, c = idx % C;
    float s = 0.f;
    for (int ch = 0; ch < chunks; ++ch) s += part[((size_t)n * chunks + ch) * C + c];
    mean[idx] = s / (float)HW;
}

// gate[n,c] = relu6(fc_w[c,:] . mean[n,:] + fc_b[c] + 3) / 6 ; warp per output
__global__ void ese_gate_kernel(const float* __restrict__ mean, const float* __restrict__ fc_w,
                                const float* __restrict__ fc_b, float* __restrict__ gate, int N, int C) {
    int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wid >= N * C) return;
    int n = wid / C, c = wid % C;
    float s = 0.f;
    for (int k = lane; k < C; k += 32) s = fmaf(fc_w[(size_t)c * C + k], mean[(size_t)n * C + k], s);
    s = warp_sum(s);
    if (lane == 0) {
        float t = s + fc_b[c] + 3.f;
        gate[wid] = fminf(fmaxf(t, 0.f), 6.f) / 6.f;
    }
}

__global__ void ese_apply_kernel(const float* __restrict__ xt, const float* __restrict__ gate,
                                 const float* __restrict__ id_f32, const fp16* __restrict__ id_hi,
                                 const fp16* __restrict__ id_lo, int id_cs, int id_co, int N, int HW, int C,
                                 float* __restrict__ y_f32, int yf_cs, int yf_co, fp16* __restrict__ y_hi,
                                 fp16* __restrict__ y_lo, int yb_cs, int yb_co, int lo_fmt) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)N * HW * C) return;
    int c = (int)(idx % C); long pix = idx / C;
    int n = (int)(pix / HW);
    float v = xt[idx] * gate[(size_t)n * C + c];
    size_t i = (size_t)pix * id_cs + id_co + c;
    if (id_f32) v += id_f32[i];
    else if (id_hi) {
        v += __half2float(id_hi[i]);
        if (id_lo) v += load_lo_scalar(id_lo + i, lo_fmt, id_co + c);
    }
    store_outputs(v, (size_t)pix * yf_cs + yf_co + c, (size_t)pix * yb_cs + yb_co + c, y_f32, y_hi, y_lo, lo_fmt, yb_co + c);
}


// ------------------------------------------------------------------------------------------ 8-channel vector helpers
struct F8 { float v[8]; };
__device__ __forceinline__ F8 ld_f8(const float* p) {
    F8 r; float4 a = ldg_f4(p), b = ldg_f4(p + 4);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void st_f8(float* p, const F8& r) {
    *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
__device__ __forceinline__ F8 ld_h8(const fp16* p) {          // 8 fp16 -> 8 floats
    F8 r; uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        r.v[2 * i] = f.x;
        r.v[2 * i + 1] = f.y;
    }
    return r;
}
// 8 channels starting at channel `ch` (a multiple of 8) of the pixel row: fp16 hi + the lo plane in either format
__device__ __forceinline__ void st_split8(fp16* hi, fp16* lo, const F8& r, int lo_fmt, int ch);
__device__ __forceinline__ F8 ld_lo8(const fp16* lo, int lo_fmt, int ch);
__device__ __forceinline__ void st_split8(fp16* hi, fp16* lo, const F8& r) {
    uint32_t ph[4], pl[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        fp16 h0, l0, h1, l1;
        split_fp16(r.v[2 * i], h0, l0);
        split_fp16(r.v[2 * i + 1], h1, l1);
        ph[i] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
        pl[i] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
    }
    *reinterpret_cast<uint4*>(hi) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    if (lo) *reinterpret_cast<uint4*>(lo) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
}

__device__ __forceinline__ void st_split8(fp16* hi, fp16* lo, const F8& r, int lo_fmt, int ch) {
    if (!lo_is_mx(lo_fmt) || !lo) { st_split8(hi, lo, r); return; }
    const float lo_scale = exp2f((float)(11 + lo_mx_exp(lo_fmt))), hi_scale = exp2f((float)lo_mx_exp(lo_fmt));
    uint32_t ph[4];
    float ls[8], hs[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        fp16 h0, h1;
        split_mx(r.v[2 * i], lo_scale, hi_scale, h0, ls[2 * i], hs[2 * i]);
        split_mx(r.v[2 * i + 1], lo_scale, hi_scale, h1, ls[2 * i + 1], hs[2 * i + 1]);
        ph[i] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
    }
    *reinterpret_cast<uint4*>(hi) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    unsigned char* q = mx_lo8_ptr(lo, ch);                       // 8-byte aligned: ch % 8 == 0 and rows are 64-byte aligned
    *reinterpret_cast<uint2*>(q) = make_uint2(pack_e4m3x4(ls[0], ls[1], ls[2], ls[3]), pack_e4m3x4(ls[4], ls[5], ls[6], ls[7]));
    *reinterpret_cast<uint2*>(q + 32) = make_uint2(pack_e4m3x4(hs[0], hs[1], hs[2], hs[3]), pack_e4m3x4(hs[4], hs[5], hs[6], hs[7]));
}
// the residual (value - hi) of 8 channels from the lo plane in either format
__device__ __forceinline__ F8 ld_lo8(const fp16* lo, int lo_fmt, int ch) {
    if (!lo_is_mx(lo_fmt)) return ld_h8(lo);
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(mx_lo8_ptr(lo, ch)));
    const float inv = exp2f(-(float)(11 + lo_mx_exp(lo_fmt)));
    F8 r;
    const uint32_t w[2] = {u.x, u.y};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = unpack_e4m3x2(w[i >> 1] >> ((i & 1) * 16));
        r.v[2 * i] = f.x * inv;
        r.v[2 * i + 1] = f.y * inv;
    }
    return r;
}

// eSE apply, 8 channels per thread (C % 8 == 0, all strides/offsets % 8 == 0)
__global__ void __launch_bounds__(256)
ese_apply_vec8_kernel(const float* __restrict__ xt, const float* __restrict__ gate, const float* __restrict__ id_f32,
                      const fp16* __restrict__ id_hi, const fp16* __restrict__ id_lo, int id_cs, int id_co, int N, int HW,
                      int C, float* __restrict__ y_f32, int yf_cs, int yf_co, fp16* __restrict__ y_hi,
                      fp16* __restrict__ y_lo, int yb_cs, int yb_co, int lo_fmt) {
    const int C8 = C >> 3;
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)N * HW * C8) return;
    const int c = (int)(idx % C8) << 3; const long pix = idx / C8;
    const int n = (int)(pix / HW);
    F8 x = ld_f8(xt + pix * C + c), g = ld_f8(gate + (size_t)n * C + c);
#pragma unroll
    for (int i = 0; i < 8; ++i) x.v[i] *= g.v[i];
    const size_t ii = (size_t)pix * id_cs + id_co + c;
    if (id_f32) {
        F8 t = ld_f8(id_f32 + ii);
#pragma unroll
        for (int i = 0; i < 8; ++i) x.v[i] += t.v[i];
    } else if (id_hi) {
        F8 t = ld_h8(id_hi + ii);
#pragma unroll
        for (int i = 0; i < 8; ++i) x.v[i] += t.v[i];
        if (id_lo) {
            F8 u = ld_lo8(id_lo + ii, lo_fmt, id_co + c);
#pragma unroll
            for (int i = 0; i < 8; ++i) x.v[i] += u.v[i];
        }
    }
    if (y_f32) st_f8(y_f32 + (size_t)pix * yf_cs + yf_co + c, x);
    if (y_hi) st_split8(y_hi + (size_t)pix * yb_cs + yb_co + c, y_lo ? y_lo + (size_t)pix * yb_cs + yb_co + c : nullptr, x, lo_fmt, yb_co + c);
}

// max-pool 3x3 s2 ceil on split-fp16 data, 8 channels per thread
__global__ void __launch_bounds__(256)
maxpool_fp16_vec8_kernel(const fp16* __restrict__ x_hi, const fp16* __restrict__ x_lo, int N, int H, int W, int C, int x_cs,
                         int x_co, fp16* __restrict__ y_hi, fp16* __restrict__ y_lo, int y_cs, int y_co, int Ho, int Wo, int lo_fmt) {
    const int C8 = C >> 3;
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)N * Ho * Wo * C8) return;
    const int c = (int)(idx % C8) << 3; long r = idx / C8;
    const int ow = (int)(r % Wo); r /= Wo;
    const int oh = (int)(r % Ho); const int n = (int)(r / Ho);
    F8 m;
#pragma unroll
    for (int i = 0; i < 8; ++i) m.v[i] = -INFINITY;
    for (int ky = 0; ky < 3; ++ky) {
        const int ih = oh * 2 + ky;
        if (ih >= H) continue;
        for (int kx = 0; kx < 3; ++kx) {
            const int iw = ow * 2 + kx;
            if (iw >= W) continue;
            const size_t i0 = (((size_t)n * H + ih) * W + iw) * x_cs + x_co + c;
            F8 v = ld_h8(x_hi + i0);
            if (x_lo) {
                F8 u = ld_lo8(x_lo + i0, lo_fmt, x_co + c);
#pragma unroll
                for (int i = 0; i < 8; ++i) v.v[i] += u.v[i];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) m.v[i] = fmaxf(m.v[i], v.v[i]);
        }
    }
    const size_t o = (((size_t)n * Ho + oh) * Wo + ow) * y_cs + y_co + c;
    st_split8(y_hi + o, y_lo ? y_lo + o : nullptr, m, lo_fmt, y_co + c);
}

// stem conv: one thread per (pixel, 16 output channels): 27 input loads feed 432 FMAs; weights broadcast from smem
__global__ void __launch_bounds__(256)
stem_conv16_kernel(const float* __restrict__ img, int N, int H, int W, const float* __restrict__ w,
                   const float* __restrict__ bias, int Cout, float* __restrict__ y_f32, fp16* __restrict__ y_hi,
                   fp16* __restrict__ y_lo, int lo_fmt) {
    extern __shared__ float sw[];     // [27][Cout] + bias[Cout]
    for (int i = threadIdx.x; i < Cout * 27; i += blockDim.x) sw[(i % 27) * Cout + i / 27] = w[i];
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[27 * Cout + i] = bias ? bias[i] : 0.f;
    __syncthreads();
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    const int cg = Cout / 16;
    // consecutive threads = consecutive pixels (coalesced image reads); channel group is the slow index
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long npix = (long)N * Ho * Wo;
    if (idx >= npix * cg) return;
    const int g = (int)(idx / npix); long r = idx - (long)g * npix;
    const int ow = (int)(r % Wo); r /= Wo;
    const int oh = (int)(r % Ho); const int n = (int)(r / Ho);
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = sw[27 * Cout + g * 16 + j];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int ih = oh * 2 + ky - 1;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int iw = ow * 2 + kx - 1;
            const bool ok = ih >= 0 && ih < H && iw >= 0 && iw < W;
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
                const float xv = ok ? __ldg(img + (((size_t)n * 3 + ci) * H + ih) * W + iw) : 0.f;
                const float4* wp = reinterpret_cast<const float4*>(sw + ((ky * 3 + kx) * 3 + ci) * Cout + g * 16);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 ww = wp[q];
                    acc[4 * q] = fmaf(xv, ww.x, acc[4 * q]); acc[4 * q + 1] = fmaf(xv, ww.y, acc[4 * q + 1]);
                    acc[4 * q + 2] = fmaf(xv, ww.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(xv, ww.w, acc[4 * q + 3]);
                }
            }
        }
    }
    const size_t o = (((size_t)n * Ho + oh) * Wo + ow) * Cout + g * 16;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        F8 v;
#pragma unroll
        for (int i = 0; i < 8; ++i) v.v[i] = fmaxf(acc[h * 8 + i], 0.f);
        if (y_f32) st_f8(y_f32 + o + h * 8, v);
        if (y_hi) st_split8(y_hi + o + h * 8, y_lo ? y_lo + o + h * 8 : nullptr, v, lo_fmt, g * 16 + h * 8);
    }
}

// Same conv, 4 consecutive output pixels x 16 channels per thread: every weight quad read from smem feeds 4 pixels (the
// one-pixel version was bound by the shared-memory return path: 432 LDS.128 per pixel) and each pixel's 16 channels
// are a full 32-byte sector of the hi and of the lo plane.  Needs Wo % 4 == 0.
__global__ void __launch_bounds__(256)
stem_conv16x4_kernel(const float* __restrict__ img, int N, int H, int W, const float* __restrict__ w,
                     const float* __restrict__ bias, int Cout, float* __restrict__ y_f32, fp16* __restrict__ y_hi,
                     fp16* __restrict__ y_lo, int lo_fmt) {
    extern __shared__ float sw[];     // [27][Cout] + bias[Cout]
    for (int i = threadIdx.x; i < Cout * 27; i += blockDim.x) sw[(i % 27) * Cout + i / 27] = w[i];
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[27 * Cout + i] = bias ? bias[i] : 0.f;
    __syncthreads();
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2, Wo4 = Wo / 4;
    const int cg = Cout / 16;
    const bool vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(img) & 15) == 0);
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long nquad = (long)N * Ho * Wo4;
    if (idx >= nquad * cg) return;
    // channel group fastest: the cg adjacent lanes that share a pixel quad broadcast-load the same inputs and together write
    // whole 128-byte pixel rows of each plane (was group-major: every row assembled from cg far-apart CTAs, 1.05 TB/s)
    const int g = (int)(idx % cg); long r = idx / cg;
    const int ow4 = (int)(r % Wo4); r /= Wo4;
    const int oh = (int)(r % Ho); const int n = (int)(r / Ho);
    float acc[4][16];
#pragma unroll
    for (int px = 0; px < 4; ++px)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[px][j] = sw[27 * Cout + g * 16 + j];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int ih = oh * 2 + ky - 1;
        const bool rowok = ih >= 0 && ih < H;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
            const float* row = img + (((size_t)n * 3 + ci) * H + (rowok ? ih : 0)) * W;
            float x[9];
            const int iw0 = ow4 * 8;
            if (vec && iw0 + 8 <= W) {                   // 3 loads instead of 9: the row segment [iw0, iw0+8) is 32-byte aligned
                if (rowok) {
                    const float4 a = ldg_f4(row + iw0), c4 = ldg_f4(row + iw0 + 4);
                    x[0] = iw0 > 0 ? __ldg(row + iw0 - 1) : 0.f;
                    x[1] = a.x; x[2] = a.y; x[3] = a.z; x[4] = a.w; x[5] = c4.x; x[6] = c4.y; x[7] = c4.z; x[8] = c4.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 9; ++j) x[j] = 0.f;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 9; ++j) {
                    const int iw = iw0 - 1 + j;
                    x[j] = (rowok && iw >= 0 && iw < W) ? __ldg(row + iw) : 0.f;
                }
            }
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float4* wp = reinterpret_cast<const float4*>(sw + ((ky * 3 + kx) * 3 + ci) * Cout + g * 16);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 ww = wp[q];
#pragma unroll
                    for (int px = 0; px < 4; ++px) {
                        const float xv = x[2 * px + kx];
                        acc[px][4 * q] = fmaf(xv, ww.x, acc[px][4 * q]); acc[px][4 * q + 1] = fmaf(xv, ww.y, acc[px][4 * q + 1]);
                        acc[px][4 * q + 2] = fmaf(xv, ww.z, acc[px][4 * q + 2]); acc[px][4 * q + 3] = fmaf(xv, ww.w, acc[px][4 * q + 3]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int px = 0; px < 4; ++px) {
        const size_t o = (((size_t)n * Ho + oh) * Wo + ow4 * 4 + px) * Cout + g * 16;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            F8 v;
#pragma unroll
            for (int i = 0; i < 8; ++i) v.v[i] = fmaxf(acc[px][h * 8 + i], 0.f);
            if (y_f32) st_f8(y_f32 + o + h * 8, v);
            if (y_hi) st_split8(y_hi + o + h * 8, y_lo ? y_lo + o + h * 8 : nullptr, v, lo_fmt, g * 16 + h * 8);
        }
    }
}

// GroupNorm in two coalesced passes.  Pass 1: per (n, pixel chunk) partial sum / sum-of-squares of every group
// (thread = 4 channels of one pixel; cpg must be a multiple of 4).  Pass 2: finalize + normalize + affine (+ReLU).
constexpr int GN_CHUNKS = 64;
__global__ void __launch_bounds__(256)
gn_partial_kernel(const float* __restrict__ x, float* __restrict__ part, int HW, int C, int groups) {
    extern __shared__ float sh[];                     // [2][groups]
    const int n = blockIdx.x, ch = blockIdx.y;
    for (int i = threadIdx.x; i < 2 * groups; i += blockDim.x) sh[i] = 0.f;
    __syncthreads();
    const int C4 = C >> 2, cpg = C / groups;
    const int per = (HW + GN_CHUNKS - 1) / GN_CHUNKS;
    const int p0 = ch * per, p1 = min(HW, p0 + per);
    const long total = (long)(p1 - p0) * C4;
    // a thread's channel quad is the same every iteration when blockDim % C4 == 0 (C = 256): accumulate in registers and
    // touch shared memory only when the group changes
    float s = 0.f, q = 0.f;
    int cur = -1;
    for (long i = threadIdx.x; i < total; i += blockDim.x) {
        const int c = (int)(i % C4) << 2; const int p = p0 + (int)(i / C4);
        const int g = c / cpg;
        if (g != cur) {
            if (cur >= 0) { atomicAdd(&sh[cur], s); atomicAdd(&sh[groups + cur], q); }
            cur = g; s = 0.f; q = 0.f;
        }
        const float4 v = ldg_f4(x + ((size_t)n * HW + p) * C + c);
        s += v.x + v.y + v.z + v.w;
        q += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    if (cur >= 0) { atomicAdd(&sh[cur], s); atomicAdd(&sh[groups + cur], q); }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * groups; i += blockDim.x)
        part[((size_t)n * GN_CHUNKS + ch) * 2 * groups + i] = sh[i];
}
__global__ void __launch_bounds__(256)
gn_apply_kernel(const float* __restrict__ x, const float* __restrict__ part, const float* __restrict__ gamma,
                const float* __restrict__ beta, int HW, int C, int groups, float eps, int relu, float* __restrict__ y_f32,
                fp16* __restrict__ y_hi, fp16* __restrict__ y_lo, int lo_fmt) {
    extern __shared__ float sh[];                     // mean[groups], rstd[groups]
    const int n = blockIdx.y;
    const int cpg = C / groups;
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        float s = 0.f, q = 0.f;
        for (int ch = 0; ch < GN_CHUNKS; ++ch) {
            s += part[((size_t)n * GN_CHUNKS + ch) * 2 * groups + g];
            q += part[((size_t)n * GN_CHUNKS + ch) * 2 * groups + groups + g];
        }
        const float cnt = (float)HW * cpg, mean = s / cnt;
        sh[g] = mean;
        sh[groups + g] = rsqrtf(fmaxf(q / cnt - mean * mean, 0.f) + eps);
    }
    __syncthreads();
    const int C8 = C >> 3;
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)HW * C8) return;
    const int c = (int)(idx % C8) << 3; const long p = idx / C8;
    const size_t o = ((size_t)n * HW + p) * C + c;
    F8 v = ld_f8(x + o), ga = ld_f8(gamma + c), be = ld_f8(beta + c);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int g = (c + i) / cpg;
        float t = (v.v[i] - sh[g]) * sh[groups + g] * ga.v[i] + be.v[i];
        v.v[i] = relu ? fmaxf(t, 0.f) : t;
    }
    if (y_f32) st_f8(y_f32 + o, v);
    if (y_hi) st_split8(y_hi + o, y_lo ? y_lo + o : nullptr, v, lo_fmt, c);
}

// ------------------------------------------------------------------------------------------ uint8 camera images
// NormalizeMultiviewImage + pad + HWC->CHW of the reference's test pipeline (datasets/pipelines/transform_3d.py:74-101,
// custom_pipeline.py:358-378 with pad_val 0; mmcv.imnormalize: (float32(x) - mean) * (1 / std), optional BGR->RGB swap)
// on the device, so a frame crosses PCIe as 1 byte per sample instead of 4.  One thread = 4 consecutive pixels of a row:
// 12 input bytes as three 32-bit loads, one float4 store per channel plane; the pad region is written as zeros.
__global__ void __launch_bounds__(256)
normalize_u8_vec4_kernel(const uint8_t* __restrict__ img, int N, int H, int W, int Hp, int Wp, float m0, float m1, float m2,
                         float s0, float s1, float s2, int swap_rb, float* __restrict__ out) {
    const int Wp4 = Wp >> 2;
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)N * Hp * Wp4) return;
    const int x0 = (int)(idx % Wp4) << 2; long r = idx / Wp4;
    const int y = (int)(r % Hp); const int n = (int)(r / Hp);
    float px[4][3];
#pragma unroll
    for (int i = 0; i < 4; ++i) px[i][0] = px[i][1] = px[i][2] = 0.f;
    if (y < H && x0 < W) {                             // W % 4 == 0: the 4 pixels are inside together
        const uint32_t* p = reinterpret_cast<const uint32_t*>(img + (((size_t)n * H + y) * W + x0) * 3);
        const uint32_t w[3] = {__ldg(p), __ldg(p + 1), __ldg(p + 2)};
#pragma unroll
        for (int b = 0; b < 12; ++b) {
            const float v = (float)((w[b >> 2] >> ((b & 3) * 8)) & 0xffu);
            const int c = b % 3;
            px[b / 3][c] = c == 0 ? (v - m0) * s0 : c == 1 ? (v - m1) * s1 : (v - m2) * s2;
        }
    }
    const size_t plane = (size_t)Hp * Wp;
    float* o = out + (size_t)n * 3 * plane + (size_t)y * Wp + x0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int cs = swap_rb ? 2 - c : c;
        *reinterpret_cast<float4*>(o + c * plane) = make_float4(px[0][cs], px[1][cs], px[2][cs], px[3][cs]);
    }
}
__global__ void normalize_u8_kernel(const uint8_t* __restrict__ img, int N, int H, int W, int Hp, int Wp, float m0, float m1,
                                    float m2, float s0, float s1, float s2, int swap_rb, float* __restrict__ out) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)N * 3 * Hp * Wp) return;
    const int x = (int)(idx % Wp); long r = idx / Wp;
    const int y = (int)(r % Hp); r /= Hp;
    const int c = (int)(r % 3); const int n = (int)(r / 3);
    float v = 0.f;
    if (y < H && x < W) {
        const int cs = swap_rb ? 2 - c : c;                  // output channel c reads input channel cs; mean / std follow the OUTPUT order
        const float raw = (float)img[(((size_t)n * H + y) * W + x) * 3 + cs];
        v = c == 0 ? (raw - m0) * s0 : c == 1 ? (raw - m1) * s1 : (raw - m2) * s2;
    }
    out[idx] = v;
}

// ------------------------------------------------------------------------------------------ FPN top-down
__global__ void upsample_add_kernel(float* __restrict__ dst, const float* __restrict__ src, int N, int Hd, int Wd, int Hs,
                                    int Ws, int C, fp16* __restrict__ d_hi, fp16* __restrict__ d_lo, int lo_fmt) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)N * Hd * Wd * C) return;
    int c = (int)(idx % C); long r = idx / C;
    int w = (int)(r % Wd); r /= Wd;
    int h = (int)(r % Hd); int n = (int)(r / Hd);
    // F.interpolate(mode='nearest'): src index = floor(dst * in/out) computed in float by torch for
    // non-integer ratios; our sizes are exact multiples so integer arithmetic is identical.
    int hs = min((int)(((long)h * Hs) / Hd), Hs - 1), ws = min((int)(((long)w * Ws) / Wd), Ws - 1);
    float v = dst[idx] + src[(((size_t)n * Hs + hs) * Ws + ws) * C + c];
    dst[idx] = v;
    if (d_hi) {
        const fp16 hh = __float2half_rn(v);
        d_hi[idx] = hh;
        if (d_lo) store_lo_scalar(v, hh, d_lo + idx, lo_fmt, c);
    }
}

// 8 channels per thread (C % 8 == 0): two 128-bit loads per operand, 128-bit stores of the fp32 map and of each fp16 plane;
// same arithmetic as the scalar kernel (one fp32 add, then the split)
__global__ void __launch_bounds__(256)
upsample_add_vec8_kernel(float* __restrict__ dst, const float* __restrict__ src, int N, int Hd, int Wd, int Hs, int Ws, int C,
                         fp16* __restrict__ d_hi, fp16* __restrict__ d_lo, int lo_fmt) {
    const int C8 = C >> 3;
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)N * Hd * Wd * C8) return;
    const int c = (int)(idx % C8) << 3; long r = idx / C8;
    const int w = (int)(r % Wd); r /= Wd;
    const int h = (int)(r % Hd); const int n = (int)(r / Hd);
    const int hs = min((int)(((long)h * Hs) / Hd), Hs - 1), ws = min((int)(((long)w * Ws) / Wd), Ws - 1);
    const size_t o = (((size_t)n * Hd + h) * Wd + w) * C + c;
    F8 a;
    {   // dst is read and written by this thread only: plain (non-.nc) loads
        const float4 a0 = *reinterpret_cast<const float4*>(dst + o), a1 = *reinterpret_cast<const float4*>(dst + o + 4);
        a.v[0] = a0.x; a.v[1] = a0.y; a.v[2] = a0.z; a.v[3] = a0.w; a.v[4] = a1.x; a.v[5] = a1.y; a.v[6] = a1.z; a.v[7] = a1.w;
    }
    const F8 b = ld_f8(src + (((size_t)n * Hs + hs) * Ws + ws) * C + c);
#pragma unroll
    for (int i = 0; i < 8; ++i) a.v[i] += b.v[i];
    st_f8(dst + o, a);
    if (d_hi) st_split8(d_hi + o, d_lo ? d_lo + o : nullptr, a, lo_fmt, c);
}

// ------------------------------------------------------------------------------------------ split / merge
__global__ void split_fp16_kernel(const float* __restrict__ x, const float* __restrict__ x_add, fp16* __restrict__ hi,
                                  fp16* __restrict__ lo, long n) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp16 h, l;
    split_fp16(x_add ? x[i] + x_add[i] : x[i], h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
}
// rows x C fp32 (dense) -> hi plane + lo plane in either format, 8 channels per thread (C % 8 == 0)
__global__ void __launch_bounds__(256)
split_planes_vec8_kernel(const float* __restrict__ x, fp16* __restrict__ hi, fp16* __restrict__ lo, long rows, int C, int lo_fmt) {
    const int C8 = C >> 3;
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * C8) return;
    const int c = (int)(idx % C8) << 3;
    const size_t o = (size_t)(idx / C8) * C + c;
    const F8 v = ld_f8(x + o);
    st_split8(hi + o, lo ? lo + o : nullptr, v, lo_fmt, c);
}
__global__ void merge_fp16_kernel(const fp16* __restrict__ hi, const fp16* __restrict__ lo, int cs, int co,
                                  float* __restrict__ y, long rows, int C, int lo_fmt) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * C) return;
    long r = i / C; int c = (int)(i % C);
    size_t s = (size_t)r * cs + co + c;
    float v = __half2float(hi[s]);
    if (lo) v += load_lo_scalar(lo + s, lo_fmt, co + c);
    y[i] = v;
}

// ------------------------------------------------------------------------------------------ fp32 implicit-GEMM conv
// y[pix, co] = relu(bias[co] + sum_{tap, ci} x[pix @ tap, ci] * w[co, tap, ci]).  Tile 128 pixels x 64 cout, BK 16.
constexpr int CV_M = 128, CV_N = 64, CV_K = 16;

__global__ void __launch_bounds__(256)
conv2d_f32_kernel(const float* __restrict__ x, int N, int H, int W, int x_cs, int x_co, int Cin,
                  const float* __restrict__ w, const float* __restrict__ bias, int Cout, int ks, int stride, int relu,
                  float* __restrict__ y, int y_cs, int y_co, int Ho, int Wo) {
    __shared__ float As[CV_K][CV_M + 4];
    __shared__ float Bs[CV_K][CV_N + 4];
    const int tid = threadIdx.x;
    const long M = (long)N * Ho * Wo;
    const long m0 = (long)blockIdx.y * CV_M;
    const int n0 = blockIdx.x * CV_N;
    const int ty = tid / 16, tx = tid % 16;
    const int taps = ks * ks, pad = ks / 2;
    const int K = taps * Cin;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int la_r = tid / 4, la_c = (tid % 4) * 4;
    // pre-decode the two pixel rows this thread loads
    int pn[2], poh[2], pow_[2]; bool pv[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        long gm = m0 + la_r + h * 64;
        pv[h] = gm < M;
        long t = pv[h] ? gm : 0;
        pow_[h] = (int)(t % Wo); t /= Wo;
        poh[h] = (int)(t % Ho); pn[h] = (int)(t / Ho);
    }
    for (int k0 = 0; k0 < K; k0 += CV_K) {
        // Cin % 4 == 0 is required, so a float4 never straddles taps
        int gk = k0 + la_c;
        int tap = gk / Cin, ci = gk - tap * Cin;
        int ky = tap / ks, kx = tap - ky * ks;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pv[h] && gk < K) {
                int ih = poh[h] * stride + ky - pad, iw = pow_[h] * stride + kx - pad;
                if (ih >= 0 && ih < H && iw >= 0 && iw < W)
                    v = *reinterpret_cast<const float4*>(x + (((size_t)pn[h] * H + ih) * W + iw) * x_cs + x_co + ci);
            }
            int r = la_r + h * 64;
            As[la_c + 0][r] = v.x; As[la_c + 1][r] = v.y; As[la_c + 2][r] = v.z; As[la_c + 3][r] = v.w;
        }
        {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            int gn = n0 + la_r;
            if (gn < Cout && gk < K) v = *reinterpret_cast<const float4*>(w + (size_t)gn * K + gk);
            Bs[la_c + 0][la_r] = v.x; Bs[la_c + 1][la_r] = v.y; Bs[la_c + 2][la_r] = v.z; Bs[la_c + 3][la_r] = v.w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < CV_K; ++k) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float b[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        long gm = m0 + ty * 8 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int gn = n0 + tx * 4 + j;
            if (gn >= Cout) continue;
            float v = acc[i][j] + (bias ? bias[gn] : 0.f);
            if (relu == 1) v = fmaxf(v, 0.f);
            else if (relu == 2) v = v / (1.f + __expf(-v));
            y[(size_t)gm * y_cs + y_co + gn] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------ GroupNorm (NHWC) + ReLU
// one block per (n, group): two passes over HW x cpg values (depth_predictor.py:44-46). Outputs fp32 and/or split fp16.
__global__ void __launch_bounds__(256)
groupnorm_nhwc_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                      int HW, int C, int groups, float eps, int relu, float* __restrict__ y_f32, fp16* __restrict__ y_hi,
                      fp16* __restrict__ y_lo, int lo_fmt) {
    __shared__ float red[2][8];
    const int n = blockIdx.x / groups, g = blockIdx.x % groups;
    const int cpg = C / groups;
    const float* xb = x + (size_t)n * HW * C + g * cpg;
    const int total = HW * cpg;
    float s = 0.f, q = 0.f;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        float v = xb[(size_t)(i / cpg) * C + (i % cpg)];
        s += v; q += v * v;
    }
    s = warp_sum(s); q = warp_sum(q);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = q; }
    __syncthreads();
    float ts = 0.f, tq = 0.f;
    for (int i = 0; i < 8; ++i) { ts += red[0][i]; tq += red[1][i]; }
    const float mean = ts / (float)total;
    const float var = fmaxf(tq / (float)total - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        int c = g * cpg + (i % cpg);
        size_t o = ((size_t)n * HW + (i / cpg)) * C + c;
        float v = (x[o] - mean) * rstd * gamma[c] + beta[c];
        if (relu) v = fmaxf(v, 0.f);
        store_outputs(v, o, o, y_f32, y_hi, y_lo, lo_fmt, c);
    }
}

}  // namespace far3d

using namespace far3d;

extern "C" int far3d_groupnorm_nhwc(const float* x, const float* gamma, const float* beta, float* workspace, int N, int HW,
                                    int C, int groups, float eps, int relu, float* y_f32, void* y_hi, void* y_lo,
                                    int lo_fmt, void* stream) {
    FAR3D_REQUIRE(x && gamma && beta && (y_f32 || y_hi), "null pointer");
    FAR3D_REQUIRE(N > 0 && HW > 0 && C > 0 && groups > 0 && C % groups == 0, "bad sizes");
    FAR3D_REQUIRE(lo_fmt == 0 || !y_lo || C % 32 == 0, "e4m3 correction plane needs C %% 32 == 0");
    if (workspace && C % 8 == 0 && (C / groups) % 4 == 0 && groups <= 256 && (uintptr_t)x % 16 == 0) {
        cudaStream_t st = (cudaStream_t)stream;
        gn_partial_kernel<<<dim3(N, GN_CHUNKS), 256, 2 * groups * sizeof(float), st>>>(x, workspace, HW, C, groups);
        int rc = launched("gn_partial_kernel");
        if (rc) return rc;
        gn_apply_kernel<<<dim3(cdiv((long)HW * (C / 8), 256), N), 256, 2 * groups * sizeof(float), st>>>(
            x, workspace, gamma, beta, HW, C, groups, eps, relu, y_f32, (fp16*)y_hi, (fp16*)y_lo, lo_fmt);
        return launched("gn_apply_kernel");
    }
    groupnorm_nhwc_kernel<<<N * groups, 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, HW, C, groups, eps, relu, y_f32,
                                                                       (fp16*)y_hi, (fp16*)y_lo, lo_fmt);
    return launched("groupnorm_nhwc_kernel");
}

extern "C" int far3d_stem_conv(const float* img_nchw, int N, int H, int W, const float* w, const float* bias, int Cout,
                               float* y_f32, void* y_hi, void* y_lo, int lo_fmt, void* stream) {
    FAR3D_REQUIRE(img_nchw && w && (y_f32 || y_hi), "null pointer");
    FAR3D_REQUIRE(N > 0 && H > 0 && W > 0 && Cout > 0 && Cout % 4 == 0 && Cout <= 256, "bad sizes");
    FAR3D_REQUIRE(lo_fmt == 0 || !y_lo || Cout % 32 == 0, "e4m3 correction plane needs Cout %% 32 == 0");
    int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    long total = (long)N * Ho * Wo * (Cout / 4);
    size_t smem = (size_t)(28 * Cout) * sizeof(float);
    if (Cout % 16 == 0 && Wo % 4 == 0) {
        long t = (long)N * Ho * (Wo / 4) * (Cout / 16);
        stem_conv16x4_kernel<<<cdiv(t, 256), 256, smem, (cudaStream_t)stream>>>(img_nchw, N, H, W, w, bias, Cout, y_f32,
                                                                              (fp16*)y_hi, (fp16*)y_lo, lo_fmt);
        return launched("stem_conv16x4_kernel");
    }
    if (Cout % 16 == 0) {
        long t16 = (long)N * Ho * Wo * (Cout / 16);
        stem_conv16_kernel<<<cdiv(t16, 256), 256, smem, (cudaStream_t)stream>>>(img_nchw, N, H, W, w, bias, Cout, y_f32,
                                                                              (fp16*)y_hi, (fp16*)y_lo, lo_fmt);
        return launched("stem_conv16_kernel");
    }
    stem_conv_kernel<<<cdiv(total, 256), 256, smem, (cudaStream_t)stream>>>(img_nchw, N, H, W, w, bias, Cout, y_f32,
                                                                           (fp16*)y_hi, (fp16*)y_lo, lo_fmt);
    return launched("stem_conv_kernel");
}

extern "C" int far3d_maxpool3x3s2(const void* x_hi, const void* x_lo, int dtype, int N, int H, int W, int C, int x_cs,
                                  int x_co, void* y_hi, void* y_lo, int y_cs, int y_co, int lo_fmt, void* stream) {
    FAR3D_REQUIRE(x_hi && y_hi, "null pointer");
    FAR3D_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0, "bad sizes");
    FAR3D_REQUIRE(lo_fmt == 0 || !x_lo || (C % 32 == 0 && x_cs % 32 == 0 && x_co % 32 == 0 && y_cs % 32 == 0 && y_co % 32 == 0),
                  "e4m3 correction planes need C, strides and offsets %% 32 == 0");
    // ceil_mode=True, no padding: out = ceil((H - 3) / 2) + 1, and the last window must start inside the input
    int Ho = (H - 3 + 1) / 2 + 1, Wo = (W - 3 + 1) / 2 + 1;
    if ((Ho - 1) * 2 >= H) --Ho;
    if ((Wo - 1) * 2 >= W) --Wo;
    long total = (long)N * Ho * Wo * C;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == 1 && C % 8 == 0 && x_cs % 8 == 0 && x_co % 8 == 0 && y_cs % 8 == 0 && y_co % 8 == 0) {
        maxpool_fp16_vec8_kernel<<<cdiv(total / 8, 256), 256, 0, st>>>((const fp16*)x_hi, (const fp16*)x_lo, N, H, W, C, x_cs,
                                                                       x_co, (fp16*)y_hi, (fp16*)y_lo, y_cs, y_co, Ho, Wo, lo_fmt);
        return launched("maxpool_fp16_vec8_kernel");
    }
    if (dtype == 1)
        maxpool_kernel<true><<<cdiv(total, 256), 256, 0, st>>>(x_hi, x_lo, N, H, W, C, x_cs, x_co, y_hi, y_lo, y_cs, y_co, Ho, Wo, lo_fmt);
    else
        maxpool_kernel<false><<<cdiv(total, 256), 256, 0, st>>>(x_hi, nullptr, N, H, W, C, x_cs, x_co, y_hi, nullptr, y_cs, y_co, Ho, Wo, 0);
    return launched("maxpool_kernel");
}

extern "C" int far3d_global_avgpool(const float* x, float* mean, float* workspace, int N, int HW, int C,
                                    void* stream) {
    FAR3D_REQUIRE(x && mean && N > 0 && HW > 0 && C > 0, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (workspace && HW >= 4 * FAR3D_AVGPOOL_CHUNKS) {   // two-stage, deterministic, fills the GPU
        dim3 grid(N, FAR3D_AVGPOOL_CHUNKS, cdiv(C, 128));
        avgpool_partial_kernel<<<grid, 128, 0, st>>>(x, workspace, N, HW, C, FAR3D_AVGPOOL_CHUNKS);
        int rc = launched("avgpool_partial_kernel");
        if (rc) return rc;
        avgpool_final_kernel<<<cdiv((long)N * C, 256), 256, 0, st>>>(workspace, mean, N, HW, C, FAR3D_AVGPOOL_CHUNKS);
        return launched("avgpool_final_kernel");
    }
    dim3 grid(N, cdiv(C, 32)), block(32, 8);
    global_avgpool_kernel<<<grid, block, 0, st>>>(x, mean, N, HW, C);
    return launched("global_avgpool_kernel");
}

extern "C" int far3d_ese_gate(const float* mean, const float* fc_w, const float* fc_b, float* gate, int N, int C,
                              void* stream) {
    FAR3D_REQUIRE(mean && fc_w && fc_b && gate && N > 0 && C > 0, "bad argument");
    ese_gate_kernel<<<cdiv((long)N * C * 32, 256), 256, 0, (cudaStream_t)stream>>>(mean, fc_w, fc_b, gate, N, C);
    return launched("ese_gate_kernel");
}

extern "C" int far3d_ese_apply(const float* xt, const float* gate, const float* id_f32, const void* id_hi,
                               const void* id_lo, int id_cs, int id_co, int N, int HW, int C, float* y_f32, int yf_cs,
                               int yf_co, void* y_hi, void* y_lo, int yb_cs, int yb_co, int lo_fmt, void* stream) {
    FAR3D_REQUIRE(xt && gate && (y_f32 || y_hi) && N > 0 && HW > 0 && C > 0, "bad argument");
    FAR3D_REQUIRE(lo_fmt == 0 || (!id_lo && !y_lo) ||
                      (C % 32 == 0 && id_cs % 32 == 0 && id_co % 32 == 0 && yb_cs % 32 == 0 && yb_co % 32 == 0),
                  "e4m3 correction planes need C, strides and offsets %% 32 == 0");
    long total = (long)N * HW * C;
    const bool vec = C % 8 == 0 && id_cs % 8 == 0 && id_co % 8 == 0 && yf_cs % 8 == 0 && yf_co % 8 == 0 && yb_cs % 8 == 0 &&
                     yb_co % 8 == 0 && (uintptr_t)xt % 16 == 0 && (uintptr_t)gate % 16 == 0;
    if (vec) {
        ese_apply_vec8_kernel<<<cdiv(total / 8, 256), 256, 0, (cudaStream_t)stream>>>(
            xt, gate, id_f32, (const fp16*)id_hi, (const fp16*)id_lo, id_cs, id_co, N, HW, C, y_f32, yf_cs, yf_co, (fp16*)y_hi,
            (fp16*)y_lo, yb_cs, yb_co, lo_fmt);
        return launched("ese_apply_vec8_kernel");
    }
    ese_apply_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(xt, gate, id_f32, (const fp16*)id_hi,
                                                                        (const fp16*)id_lo, id_cs, id_co, N, HW, C, y_f32,
                                                                        yf_cs, yf_co, (fp16*)y_hi, (fp16*)y_lo, yb_cs, yb_co, lo_fmt);
    return launched("ese_apply_kernel");
}

extern "C" int far3d_upsample_add(float* dst, const float* src, int N, int Hd, int Wd, int Hs, int Ws, int C, void* d_hi,
                                  void* d_lo, int lo_fmt, void* stream) {
    FAR3D_REQUIRE(dst && src && N > 0 && Hd > 0 && Wd > 0 && Hs > 0 && Ws > 0 && C > 0, "bad argument");
    FAR3D_REQUIRE(lo_fmt == 0 || !d_lo || C % 32 == 0, "e4m3 correction plane needs C %% 32 == 0");
    long total = (long)N * Hd * Wd * C;
    if (C % 8 == 0 && (uintptr_t)dst % 16 == 0 && (uintptr_t)src % 16 == 0 && (uintptr_t)d_hi % 16 == 0 &&
        (uintptr_t)d_lo % 16 == 0) {
        upsample_add_vec8_kernel<<<cdiv(total / 8, 256), 256, 0, (cudaStream_t)stream>>>(dst, src, N, Hd, Wd, Hs, Ws, C,
                                                                                        (fp16*)d_hi, (fp16*)d_lo, lo_fmt);
        return launched("upsample_add_vec8_kernel");
    }
    upsample_add_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(dst, src, N, Hd, Wd, Hs, Ws, C, (fp16*)d_hi,
                                                                           (fp16*)d_lo, lo_fmt);
    return launched("upsample_add_kernel");
}

extern "C" int far3d_normalize_u8(const uint8_t* img_nhwc, int N, int H, int W, int Hp, int Wp, const float* mean_host,
                                  const float* std_host, int to_rgb, float* out_nchw, void* stream) {
    FAR3D_REQUIRE(img_nhwc && mean_host && std_host && out_nchw, "null pointer");
    FAR3D_REQUIRE(N > 0 && H > 0 && W > 0 && Hp >= H && Wp >= W, "bad sizes (padded size must cover the image)");
    FAR3D_REQUIRE(std_host[0] != 0.f && std_host[1] != 0.f && std_host[2] != 0.f, "std must be non-zero");
    // mmcv.imnormalize: stdinv = 1 / float64(std), applied to the float32 image
    const float s0 = (float)(1.0 / (double)std_host[0]), s1 = (float)(1.0 / (double)std_host[1]), s2 = (float)(1.0 / (double)std_host[2]);
    cudaStream_t st = (cudaStream_t)stream;
    if (!to_rgb && W % 4 == 0 && Wp % 4 == 0 && (uintptr_t)img_nhwc % 4 == 0 && (uintptr_t)out_nchw % 16 == 0) {
        const long t = (long)N * Hp * (Wp / 4);
        normalize_u8_vec4_kernel<<<cdiv(t, 256), 256, 0, st>>>(img_nhwc, N, H, W, Hp, Wp, mean_host[0], mean_host[1], mean_host[2],
                                                               s0, s1, s2, 0, out_nchw);
        return launched("normalize_u8_vec4_kernel");
    }
    const long t = (long)N * 3 * Hp * Wp;
    normalize_u8_kernel<<<cdiv(t, 256), 256, 0, st>>>(img_nhwc, N, H, W, Hp, Wp, mean_host[0], mean_host[1], mean_host[2], s0, s1,
                                                      s2, to_rgb ? 1 : 0, out_nchw);
    return launched("normalize_u8_kernel");
}

extern "C" int far3d_split_fp16(const float* x, const float* x_add, void* hi, void* lo, int64_t n, void* stream) {
    FAR3D_REQUIRE(x && hi && n > 0, "bad argument");
    split_fp16_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(x, x_add, (fp16*)hi, (fp16*)lo, n);
    return launched("split_fp16_kernel");
}
extern "C" int far3d_merge_fp16(const void* hi, const void* lo, float* y, int64_t n, void* stream) {
    FAR3D_REQUIRE(hi && y && n > 0, "bad argument");
    merge_fp16_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>((const fp16*)hi, (const fp16*)lo, 1, 0, y, n, 1, 0);
    return launched("merge_fp16_kernel");
}
extern "C" int far3d_merge_fp16_strided(const void* hi, const void* lo, int lo_fmt, int cs, int co, float* y, int64_t rows,
                                        int C, void* stream) {
    FAR3D_REQUIRE(hi && y && rows > 0 && C > 0 && cs >= C, "bad argument");
    FAR3D_REQUIRE(lo_fmt == 0 || !lo || (cs % 32 == 0 && co % 32 == 0 && C % 32 == 0), "e4m3 correction plane needs cs, co, C %% 32 == 0");
    merge_fp16_kernel<<<cdiv(rows * C, 256), 256, 0, (cudaStream_t)stream>>>((const fp16*)hi, (const fp16*)lo, cs, co, y, rows, C,
                                                                            lo_fmt);
    return launched("merge_fp16_kernel");
}
extern "C" int far3d_split_planes(const float* x, void* hi, void* lo, int lo_fmt, int64_t rows, int C, void* stream) {
    FAR3D_REQUIRE(x && hi && rows > 0 && C > 0 && C % 8 == 0, "bad argument (C %% 8 == 0)");
    FAR3D_REQUIRE((uintptr_t)x % 16 == 0 && (uintptr_t)hi % 16 == 0 && (uintptr_t)lo % 16 == 0, "16-byte aligned pointers");
    FAR3D_REQUIRE(lo_fmt == 0 || !lo || C % 32 == 0, "e4m3 correction plane needs C %% 32 == 0");
    split_planes_vec8_kernel<<<cdiv(rows * (C / 8), 256), 256, 0, (cudaStream_t)stream>>>(x, (fp16*)hi, (fp16*)lo, rows, C, lo_fmt);
    return launched("split_planes_vec8_kernel");
}

extern "C" int far3d_conv2d_f32(const float* x, int N, int H, int W, int x_cs, int x_co, int Cin, const float* w,
                                const float* bias, int Cout, int ksize, int stride, int relu, float* y, int y_cs,
                                int y_co, void* stream) {
    FAR3D_REQUIRE(x && w && y, "null pointer");
    FAR3D_REQUIRE(N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "non-positive size");
    FAR3D_REQUIRE((ksize == 1 || ksize == 3) && (stride == 1 || stride == 2), "ksize in {1,3}, stride in {1,2}");
    FAR3D_REQUIRE(Cin % 4 == 0 && x_cs % 4 == 0 && x_co % 4 == 0 && (uintptr_t)x % 16 == 0 && (uintptr_t)w % 16 == 0,
                  "Cin, x_cs, x_co multiples of 4; 16-byte aligned pointers");
    int pad = ksize / 2;
    int Ho = (H + 2 * pad - ksize) / stride + 1, Wo = (W + 2 * pad - ksize) / stride + 1;
    long M = (long)N * Ho * Wo;
    dim3 grid(cdiv(Cout, CV_N), cdiv(M, CV_M));
    conv2d_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, N, H, W, x_cs, x_co, Cin, w, bias, Cout, ksize, stride,
                                                             relu, y, y_cs, y_co, Ho, Wo);
    return launched("conv2d_f32_kernel");
}
