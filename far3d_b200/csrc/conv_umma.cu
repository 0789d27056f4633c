// Implicit-GEMM convolution / GEMM on Blackwell tensor cores (tcgen05.mma, TMEM accumulators, TMA-fed smem).
//
// Replaces the cuDNN conv + BatchNorm(eval) + ReLU triples of models/backbones/vovnet.py:124-161 (80 conv3x3,
// 16 concat conv1x1, stem convs 2-3), the mmdet FPN lateral / output / extra convs (config far3d.py:50-57) and -
// with ksize 1 on a [rows, K] "image" - nn.Linear layers of the decoder.
//
// GEMM view: M = output pixels, N = Cout, K = taps * Cin.  One CTA computes a 128 x BN tile:
//   * A tile (128 pixels x 64 channels of ONE filter tap) is a plain tiled-TMA box {64, tw, th, 1} of the NHWC
//     activation tensor at spatially shifted coordinates (oh0 + ky - 1, ow0 + kx - 1); TMA zero-fills outside the
//     image, which is exactly the conv zero padding, and lands the 128 B rows in SWIZZLE_128B order = the canonical
//     K-major UMMA operand layout.  No im2col buffer, no descriptor-mode im2col.  Stride-2 convs view the
//     tensor as {2*C, W/2, 2, H/2, N} so that a tap is again a dense box.
//   * B tile (BN filters x 64 channels of the tap) is a box {64, 1, BN} of the [Cout, taps, Cin] weight tensor.
//   * warp 0: TMA producer, warp 1: single-thread tcgen05.mma issuer (M=128, N=BN, K=16, cta_group::1, fp32 accum in
//     TMEM), warps 2-5: epilogue (tcgen05.ld -> bias + ReLU -> fp32 / bf16 / split-bf16 stores with channel offset, so
//     OSA concat buffers are written in place and torch.cat of vovnet.py:230 disappears).
//   * NS-stage mbarrier ring (full/empty) between TMA and MMA, tcgen05.commit releases stages.
//
// "bf16x3" (split) mode: activations and weights are stored as bf16 hi + bf16 lo planes (value = hi + lo); each k-step
// issues hi*hi + lo*hi + hi*lo into the same fp32 accumulator, which reproduces fp32 convolution to ~2^-17 relative
// (the reference computes in fp32/TF32) while staying on the bf16 tensor pipe.
#include <cuda.h>
#include "common.cuh"

namespace far3d {

typedef __nv_bfloat16 bf16;

struct ConvParams {
    int N, H, W, Ho, Wo;          // input / output spatial dims
    int Cin, Cout, ks, stride;
    int tw, th, tiles_w, tiles_h; // M tile = th x tw output pixels (tw*th == 128)
    int bn;                       // N tile
    int kchunks;                  // ceil(Cin / 64)
    int x_cs, x_co;               // used for stride-2 channel coordinate
    int num_stages;
    int relu;
    const float* bias;
    float* y_f32; int yf_cs, yf_co; long long yf_ns;
    bf16* y_hi; bf16* y_lo; int yb_cs, yb_co;
};

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded spin: a pipeline bug traps (launch error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();     // ~2 s at 1.9 GHz
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major), bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                          // descriptor version 1 (Blackwell), bits [46,48)
    d |= (uint64_t)2 << 61;                          // layout type SWIZZLE_128B, bits [61,64)
    return d;
}
// instruction descriptor for kind::f16: D=f32, A=B=bf16, both K-major, M=128, N=bn (cute::UMMA::InstrDescriptor)
__device__ __forceinline__ uint32_t umma_idesc_bf16(int bn) {
    uint32_t d = 0;
    d |= 1u << 4;                    // c_format = F32
    d |= 1u << 7;                    // a_format = BF16
    d |= 1u << 10;                   // b_format = BF16
    d |= (uint32_t)(bn >> 3) << 17;  // n_dim
    d |= (uint32_t)(128 >> 4) << 24; // m_dim
    return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

constexpr int UM_THREADS = 192;   // warp 0 TMA, warp 1 MMA (+TMEM alloc), warps 2..5 epilogue
constexpr int UM_BM = 128, UM_BK = 64;
constexpr int UM_A_BYTES = UM_BM * UM_BK * 2;   // 16 KB per plane

template <bool SPLIT>
__global__ void __launch_bounds__(UM_THREADS, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                 const ConvParams p) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ uint64_t bar_full[8], bar_empty[8], bar_acc;
    __shared__ uint32_t s_tmem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NS = p.num_stages;
    const uint32_t b_bytes = (uint32_t)p.bn * UM_BK * 2;
    const uint32_t stage_bytes = (SPLIT ? 2u : 1u) * (UM_A_BYTES + b_bytes);
    unsigned char* tiles = (unsigned char*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);

    // tile coordinates
    int mt = blockIdx.x;
    const int txi = mt % p.tiles_w; mt /= p.tiles_w;
    const int tyi = mt % p.tiles_h; const int img = mt / p.tiles_h;
    const int ow0 = txi * p.tw, oh0 = tyi * p.th;
    const int n0 = blockIdx.y * p.bn;
    const int taps = p.ks * p.ks, pad = p.ks / 2;
    const int KT = taps * p.kchunks;

    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < p.bn) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
        mbar_init(&bar_acc, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            for (int it = 0; it < KT; ++it) {
                const int s = it % NS;
                const uint32_t ph = (uint32_t)(it / NS) & 1u;
                mbar_wait(&bar_empty[s], ph ^ 1u);
                mbar_expect_tx(&bar_full[s], stage_bytes);
                const int tap = it / p.kchunks, kc = it - tap * p.kchunks;
                const int ky = tap / p.ks, kx = tap - ky * p.ks;
                unsigned char* sa = tiles + (size_t)s * stage_bytes;
                unsigned char* sb = sa + (SPLIT ? 2 : 1) * UM_A_BYTES;
                const int c0 = kc * UM_BK;
                if (p.stride == 1) {
                    const int cw = ow0 + kx - pad, chh = oh0 + ky - pad;
                    tma_load_4d(sa, &tmA_hi, &bar_full[s], c0, cw, chh, img);
                    if (SPLIT) tma_load_4d(sa + UM_A_BYTES, &tmA_lo, &bar_full[s], c0, cw, chh, img);
                } else {
                    const int dy = ky - pad, dx = kx - pad;
                    const int hpar = dy & 1, wpar = dx & 1;
                    const int hoff = (dy - hpar) / 2, woff = (dx - wpar) / 2;
                    const int cc = wpar * p.x_cs + p.x_co + c0;
                    tma_load_5d(sa, &tmA_hi, &bar_full[s], cc, ow0 + woff, hpar, oh0 + hoff, img);
                    if (SPLIT) tma_load_5d(sa + UM_A_BYTES, &tmA_lo, &bar_full[s], cc, ow0 + woff, hpar, oh0 + hoff, img);
                }
                tma_load_3d(sb, &tmB_hi, &bar_full[s], c0, tap, n0);
                if (SPLIT) tma_load_3d(sb + b_bytes, &tmB_lo, &bar_full[s], c0, tap, n0);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        const uint32_t idesc = umma_idesc_bf16(p.bn);
        for (int it = 0; it < KT; ++it) {
            const int s = it % NS;
            const uint32_t ph = (uint32_t)(it / NS) & 1u;
            mbar_wait(&bar_full[s], ph);
            tc_fence_after();
            if (lane == 0) {
                const int kc = it % p.kchunks;
                const int kvalid = min(UM_BK, p.Cin - kc * UM_BK);
                const int ksteps = (kvalid + 15) / 16;
                const uint32_t sa = smem_u32(tiles + (size_t)s * stage_bytes);
                const uint32_t sb = sa + (SPLIT ? 2 : 1) * UM_A_BYTES;
                const uint64_t a_hi = umma_desc_sw128(sa), b_hi = umma_desc_sw128(sb);
                for (int k = 0; k < ksteps; ++k) {
                    const uint64_t koff = (uint64_t)(k * 32 >> 4);        // 16 bf16 = 32 B along K inside the swizzle atom
                    const uint32_t first = (it > 0 || k > 0) ? 1u : 0u;
                    if (SPLIT) {
                        const uint64_t a_lo = umma_desc_sw128(sa + UM_A_BYTES), b_lo = umma_desc_sw128(sb + b_bytes);
                        umma_bf16(tmem_base, a_lo + koff, b_hi + koff, idesc, first);
                        umma_bf16(tmem_base, a_hi + koff, b_lo + koff, idesc, 1u);
                        umma_bf16(tmem_base, a_hi + koff, b_hi + koff, idesc, 1u);
                    } else {
                        umma_bf16(tmem_base, a_hi + koff, b_hi + koff, idesc, first);
                    }
                }
                umma_commit(&bar_empty[s]);                 // frees the smem stage when these MMAs retire
                if (it == KT - 1) umma_commit(&bar_acc);    // accumulator complete
            }
            __syncwarp();
        }
    } else {
        // ================= epilogue: TMEM -> registers -> global =================
        const int quad = warp & 3;                          // TMEM lane quadrant this warp may access
        const int m = quad * 32 + lane;                     // row of the tile = pixel
        const int hh = m / p.tw, ww = m - hh * p.tw;
        const int oh = oh0 + hh, ow = ow0 + ww;
        const bool pix_ok = (oh < p.Ho) && (ow < p.Wo);
        const size_t pix = ((size_t)img * p.Ho + oh) * p.Wo + ow;
        float* yf = p.y_f32 ? p.y_f32 + (size_t)img * p.yf_ns + ((size_t)oh * p.Wo + ow) * p.yf_cs + p.yf_co : nullptr;
        bf16* yh = p.y_hi ? p.y_hi + pix * p.yb_cs + p.yb_co : nullptr;
        bf16* yl = p.y_lo ? p.y_lo + pix * p.yb_cs + p.yb_co : nullptr;
        mbar_wait(&bar_acc, 0);
        tc_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
        for (int c = 0; c < p.bn; c += 16) {
            uint32_t r[16];
            tmem_ld16(trow + (uint32_t)c, r);
            tmem_ld_wait();
            const int col0 = n0 + c;
            if (!pix_ok || col0 >= p.Cout) continue;
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float t = __uint_as_float(r[j]);
                if (p.bias && col0 + j < p.Cout) t += __ldg(p.bias + col0 + j);
                if (p.relu == 1) t = fmaxf(t, 0.f);
                else if (p.relu == 2) t = t / (1.f + __expf(-t));      // Swish (YOLOX towers)
                v[j] = t;
            }
            if (col0 + 16 <= p.Cout) {
                if (yf) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        *reinterpret_cast<float4*>(yf + col0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
                if (yh) {
                    uint32_t ph[8], pl[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        bf16 h0, l0, h1, l1;
                        split_bf16(v[2 * j], h0, l0);
                        split_bf16(v[2 * j + 1], h1, l1);
                        ph[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                        pl[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                    }
                    uint4* dh = reinterpret_cast<uint4*>(yh + col0);
                    dh[0] = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                    dh[1] = make_uint4(ph[4], ph[5], ph[6], ph[7]);
                    if (yl) {
                        uint4* dl = reinterpret_cast<uint4*>(yl + col0);
                        dl[0] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                        dl[1] = make_uint4(pl[4], pl[5], pl[6], pl[7]);
                    }
                }
            } else {
                for (int j = 0; j < 16 && col0 + j < p.Cout; ++j) {
                    if (yf) yf[col0 + j] = v[j];
                    if (yh) {
                        bf16 h, l;
                        split_bf16(v[j], h, l);
                        yh[col0 + j] = h;
                        if (yl) yl[col0 + j] = l;
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)f;
    }
    return fn;
}

static int encode(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                  const cuuint32_t* box) {
    EncodeTiledFn fn = get_encode();
    if (!fn) return fail(FAR3D_E_CUDA, "%scuTensorMapEncodeTiled unavailable (no driver)", "");
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(FAR3D_E_CUDA, "%scuTensorMapEncodeTiled failed (CUresult %ld, rank %ld)", "", (long)r, rank);
    return FAR3D_OK;
}

static int g_force_bn = 0, g_force_stages = 0;

}  // namespace far3d

using namespace far3d;

// tuning hooks for experiments (not part of the reference-facing ABI): force N-tile / stage count (0 = heuristic)
extern "C" void far3d_conv_umma_tune(int bn, int stages) { g_force_bn = bn; g_force_stages = stages; }

extern "C" int far3d_conv2d_umma(const void* x_hi, const void* x_lo, int N, int H, int W, int x_cs, int x_co, int Cin,
                                 const void* w_hi, const void* w_lo, const float* bias, int Cout, int ksize, int stride,
                                 int relu, float* y_f32, int yf_cs, int yf_co, int64_t yf_ns, void* y_hi, void* y_lo,
                                 int yb_cs, int yb_co, void* stream) {
    FAR3D_REQUIRE(x_hi && w_hi && (y_f32 || y_hi), "null pointer");
    FAR3D_REQUIRE((x_lo == nullptr) == (w_lo == nullptr), "x_lo and w_lo must both be given (split mode) or both NULL");
    FAR3D_REQUIRE(N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "non-positive size");
    FAR3D_REQUIRE((ksize == 1 || ksize == 3) && (stride == 1 || (stride == 2 && ksize == 3)), "ksize/stride unsupported");
    FAR3D_REQUIRE(Cin % 16 == 0 && x_cs % 8 == 0 && x_co % 8 == 0, "Cin %% 16, x_cs %% 8, x_co %% 8");
    FAR3D_REQUIRE((uintptr_t)x_hi % 16 == 0 && (uintptr_t)w_hi % 16 == 0, "16-byte aligned operands");
    FAR3D_REQUIRE(!y_f32 || (yf_cs % 4 == 0 && yf_co % 4 == 0 && (uintptr_t)y_f32 % 16 == 0), "fp32 output alignment");
    FAR3D_REQUIRE(!y_hi || (yb_cs % 8 == 0 && yb_co % 8 == 0 && (uintptr_t)y_hi % 16 == 0), "bf16 output alignment");
    if (stride == 2) FAR3D_REQUIRE(H % 2 == 0 && W % 2 == 0 && Cin % 64 == 0, "stride 2 needs even H, W and Cin %% 64 == 0");
    const bool split = x_lo != nullptr;
    const int pad = ksize / 2;
    ConvParams p;
    p.N = N; p.H = H; p.W = W;
    p.Ho = (H + 2 * pad - ksize) / stride + 1; p.Wo = (W + 2 * pad - ksize) / stride + 1;
    p.Cin = Cin; p.Cout = Cout; p.ks = ksize; p.stride = stride;
    p.x_cs = x_cs; p.x_co = x_co; p.relu = relu; p.bias = bias;
    p.y_f32 = y_f32; p.yf_cs = yf_cs; p.yf_co = yf_co;
    p.yf_ns = yf_ns > 0 ? yf_ns : (long long)p.Ho * p.Wo * yf_cs;
    p.y_hi = (bf16*)y_hi; p.y_lo = (bf16*)y_lo; p.yb_cs = yb_cs; p.yb_co = yb_co;
    p.kchunks = (Cin + UM_BK - 1) / UM_BK;

    // ---- M tile shape: th x tw = 128 with the fewest tiles
    int best_tw = 128; long best_tiles = -1;
    for (int tw = 8; tw <= 128; tw <<= 1) {
        int th = 128 / tw;
        long t = (long)((p.Wo + tw - 1) / tw) * ((p.Ho + th - 1) / th);
        if (best_tiles < 0 || t < best_tiles || (t == best_tiles && tw > best_tw)) { best_tiles = t; best_tw = tw; }
    }
    p.tw = best_tw; p.th = 128 / best_tw;
    p.tiles_w = (p.Wo + p.tw - 1) / p.tw; p.tiles_h = (p.Ho + p.th - 1) / p.th;
    const long m_tiles = (long)N * p.tiles_w * p.tiles_h;

    // ---- N tile: minimise waves * (bn + overhead)
    int bn = 0;
    if (g_force_bn > 0) bn = g_force_bn;
    else {
        const int cand[] = {256, 224, 192, 160, 128, 112, 96, 80, 64, 48, 32, 16};
        double best = 1e30;
        for (int c : cand) {
            if (c > ((Cout + 15) / 16) * 16) continue;
            long ctas = m_tiles * ((Cout + c - 1) / c);
            long waves = (ctas + 147) / 148;
            double cost = (double)waves * (c + 48) * (1.0 + 0.02 * (((Cout + c - 1) / c) * c - Cout));
            if (cost < best) { best = cost; bn = c; }
        }
    }
    FAR3D_REQUIRE(bn >= 16 && bn <= 256 && bn % 16 == 0, "bad N tile");
    p.bn = bn;
    const int n_tiles = (Cout + bn - 1) / bn;

    // ---- stages
    const size_t stage_bytes = (size_t)(split ? 2 : 1) * (UM_A_BYTES + (size_t)bn * UM_BK * 2);
    int ns = g_force_stages > 0 ? g_force_stages : (int)((200 * 1024) / stage_bytes);
    if (ns > 8) ns = 8;
    if (ns < 2) ns = 2;
    const int KT = ksize * ksize * p.kchunks;
    if (ns > KT) ns = KT < 2 ? 2 : KT;
    p.num_stages = ns;
    const size_t smem = ns * stage_bytes + 1024;
    if (smem > 227 * 1024) return fail(FAR3D_E_UNSUPPORTED, "%sconv_umma smem %ld exceeds 227 KB", "", (long)smem);

    // ---- tensor maps
    CUtensorMap tmA_hi, tmA_lo, tmB_hi, tmB_lo;
    int rc;
    auto mapA = [&](CUtensorMap* tm, const void* base) -> int {
        if (stride == 1) {
            cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
            cuuint64_t str[3] = {(cuuint64_t)x_cs * 2, (cuuint64_t)W * x_cs * 2, (cuuint64_t)H * W * x_cs * 2};
            cuuint32_t box[4] = {(cuuint32_t)UM_BK, (cuuint32_t)p.tw, (cuuint32_t)p.th, 1};
            return encode(tm, (const bf16*)base + x_co, 4, dims, str, box);
        }
        cuuint64_t dims[5] = {(cuuint64_t)2 * x_cs, (cuuint64_t)W / 2, 2, (cuuint64_t)H / 2, (cuuint64_t)N};
        cuuint64_t str[4] = {(cuuint64_t)2 * x_cs * 2, (cuuint64_t)W * x_cs * 2, (cuuint64_t)2 * W * x_cs * 2,
                             (cuuint64_t)H * W * x_cs * 2};
        cuuint32_t box[5] = {(cuuint32_t)UM_BK, (cuuint32_t)p.tw, 1, (cuuint32_t)p.th, 1};
        return encode(tm, base, 5, dims, str, box);
    };
    auto mapB = [&](CUtensorMap* tm, const void* base) -> int {
        cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)(ksize * ksize), (cuuint64_t)Cout};
        cuuint64_t str[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)ksize * ksize * Cin * 2};
        cuuint32_t box[3] = {(cuuint32_t)UM_BK, 1, (cuuint32_t)bn};
        return encode(tm, base, 3, dims, str, box);
    };
    if ((rc = mapA(&tmA_hi, x_hi))) return rc;
    if ((rc = mapB(&tmB_hi, w_hi))) return rc;
    if (split) {
        if ((rc = mapA(&tmA_lo, x_lo))) return rc;
        if ((rc = mapB(&tmB_lo, w_lo))) return rc;
    } else { tmA_lo = tmA_hi; tmB_lo = tmB_hi; }

    if (m_tiles > 0x7fffffffL || n_tiles > 65535) return fail(FAR3D_E_UNSUPPORTED, "%sgrid too large", "");
    dim3 grid((unsigned)m_tiles, (unsigned)n_tiles);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    if (split) {
        e = cudaFuncSetAttribute(conv_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(FAR3D_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        conv_umma_kernel<true><<<grid, UM_THREADS, smem, st>>>(tmA_hi, tmA_lo, tmB_hi, tmB_lo, p);
    } else {
        e = cudaFuncSetAttribute(conv_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(FAR3D_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        conv_umma_kernel<false><<<grid, UM_THREADS, smem, st>>>(tmA_hi, tmA_lo, tmB_hi, tmB_lo, p);
    }
    return launched("conv_umma_kernel");
}
