// nn.Linear for the decoder's token matrices (~1000 rows, K = 256) on warp-level tensor-core MMAs, fp32-grade.
//
//   y[M, N] = act((x (+ x_add))[M, K] @ w[N, K]^T + bias) (+ residual[M, N])
//
// Replaces the cuBLAS GEMMs inside mmcv MultiheadAttention / FFN and the head MLPs (models/utils/detr3d_transformer.py:311-422,
// 503-512; dense_heads/farhead.py:228-282).  far3d_linear_umma runs these on the persistent tcgen05 conv kernel: 10-15 us per
// GEMM for ~1 us of MMAs (a split kernel for the activations, then TMEM allocation, barrier set-up, cluster syncs, a TMA round trip
// and a TMEM epilogue per launch), 61 times per frame - and a CTA of that kernel (192 threads x 255 registers, ~200 KB of shared
// memory) can never sit next to a backbone conv CTA of the other frame in flight.  This kernel splits the fp32 activations into
// fp16 hi + lo on the fly (no split launch), takes the cached fp16 hi / lo weight planes, issues three mma.sync.m16n8k16 per product
// (hi.hi + lo.hi + hi.lo, fp32 accumulate), and a CTA is 128 threads, ~100 registers and 20 KB of shared memory.
//
// CTA = (16 * WARPS) rows x 64 columns; warp = 16 rows x 64 columns (8 n-tiles).  K runs in chunks of 32: the chunk's weight rows
// (64 x 32 halves x 2 planes) arrive by cp.async in double-buffered shared memory (row stride 40 halves: the 8 rows a fragment load
// touches fall in distinct banks); the A fragments are loaded straight from global memory into registers one chunk ahead.
#include "common.cuh"

namespace far3d {

constexpr int LM_BN = 64, LM_BK = 32, LM_LD = 40;
constexpr int LM_PLANE = LM_BN * LM_LD;                    // halves per plane and buffer

__device__ __forceinline__ void lm_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void lm_split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(x0, x1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void lm_cp_async16(void* dst, const void* src, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    const int n = valid ? 16 : 0;                                                   // src-size 0 -> zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}

template <int WARPS, bool ADD>
__global__ void __launch_bounds__(WARPS * 32)
linear_mma_kernel(const float* __restrict__ x, const float* __restrict__ x_add, int ldx, const __half* __restrict__ w_hi,
                  const __half* __restrict__ w_lo, const float* __restrict__ bias, const float* __restrict__ residual, int ldr,
                  float* __restrict__ y, int ldy, int M, int N, int K, int act) {
    __shared__ __align__(16) __half sb[2][2][LM_PLANE];          // [buffer][hi | lo][n][LM_LD]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int m0 = blockIdx.y * (16 * WARPS), n0 = blockIdx.x * LM_BN;
    const int row0 = m0 + warp * 16 + g, row1 = row0 + 8;
    const bool ok0 = row0 < M, ok1 = row1 < M;
    const float* x0p = x + (size_t)(ok0 ? row0 : 0) * ldx + 2 * t;
    const float* x1p = x + (size_t)(ok1 ? row1 : 0) * ldx + 2 * t;
    const float* a0p = ADD ? x_add + (size_t)(ok0 ? row0 : 0) * ldx + 2 * t : nullptr;
    const float* a1p = ADD ? x_add + (size_t)(ok1 ? row1 : 0) * ldx + 2 * t : nullptr;

    // weight chunk loader: 64 rows x 4 pieces of 8 halves x 2 planes
    auto load_b = [&](int chunk, int buf) {
        for (int pc = tid; pc < LM_BN * 4 * 2; pc += WARPS * 32) {
            const int plane = pc / (LM_BN * 4), r = (pc >> 2) % LM_BN, piece = pc & 3;
            const int n = n0 + r;
            const __half* src = (plane ? w_lo : w_hi) + (size_t)(n < N ? n : 0) * K + chunk * LM_BK + piece * 8;
            lm_cp_async16(&sb[buf][plane][r * LM_LD + piece * 8], src, n < N);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // A chunk (rows row0 / row1, 32 k): 2 k-steps x {k lo, k hi} x {row0, row1} float2
    float2 ax[2][2][2];
    auto load_a = [&](int chunk) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const int kk = chunk * LM_BK + 16 * ks + 8 * hf;
                float2 v0 = make_float2(0.f, 0.f), v1 = v0;
                if (ok0) { v0 = *reinterpret_cast<const float2*>(x0p + kk); if (ADD) { const float2 a = *reinterpret_cast<const float2*>(a0p + kk); v0.x += a.x; v0.y += a.y; } }
                if (ok1) { v1 = *reinterpret_cast<const float2*>(x1p + kk); if (ADD) { const float2 a = *reinterpret_cast<const float2*>(a1p + kk); v1.x += a.x; v1.y += a.y; } }
                ax[ks][hf][0] = v0; ax[ks][hf][1] = v1;
            }
    };

    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;

    const int nchunks = K / LM_BK;
    load_b(0, 0);
    load_a(0);
    for (int c = 0; c < nchunks; ++c) {
        const int buf = c & 1;
        // this chunk's A fragments (split now, so the registers can take the next chunk's loads)
        uint32_t ah[2][4], al[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                lm_split2(ax[ks][hf][0].x, ax[ks][hf][0].y, ah[ks][2 * hf], al[ks][2 * hf]);              // row g
                lm_split2(ax[ks][hf][1].x, ax[ks][hf][1].y, ah[ks][2 * hf + 1], al[ks][2 * hf + 1]);      // row g + 8
            }
        if (c + 1 < nchunks) {
            load_b(c + 1, buf ^ 1);                      // (the other buffer: its last readers passed the barrier at the end of chunk c - 1)
            load_a(c + 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();                                 // chunk c's weights are in shared memory for every warp
        const __half *bh = sb[buf][0], *bl = sb[buf][1];
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const int off = (8 * j + g) * LM_LD + 16 * ks + 2 * t;
                const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(bh + off), bh1 = *reinterpret_cast<const uint32_t*>(bh + off + 8);
                const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(bl + off), bl1 = *reinterpret_cast<const uint32_t*>(bl + off + 8);
                lm_mma(acc[j], al[ks], bh0, bh1);
                lm_mma(acc[j], ah[ks], bl0, bl1);
                lm_mma(acc[j], ah[ks], bh0, bh1);
            }
        __syncthreads();                                 // every warp is done with buffer `buf` before chunk c + 2 overwrites it
    }

    // ---- epilogue: element i of n-tile j = row (i < 2 ? row0 : row1), column n0 + 8 j + 2 t + (i & 1)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int col = n0 + 8 * j + 2 * t;
        if (col >= N) continue;
        const bool two = col + 1 < N;
        const float b0 = bias ? __ldg(bias + col) : 0.f, b1 = (bias && two) ? __ldg(bias + col + 1) : 0.f;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int row = half ? row1 : row0;
            if (row >= M) continue;
            float v0 = acc[j][2 * half] + b0, v1 = acc[j][2 * half + 1] + b1;
            if (act == 1) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
            if (residual) {
                v0 += residual[(size_t)row * ldr + col];
                if (two) v1 += residual[(size_t)row * ldr + col + 1];
            }
            float* dst = y + (size_t)row * ldy + col;
            if (two && ((reinterpret_cast<uintptr_t>(dst) & 7) == 0)) *reinterpret_cast<float2*>(dst) = make_float2(v0, v1);
            else { dst[0] = v0; if (two) dst[1] = v1; }
        }
    }
}

}  // namespace far3d

using namespace far3d;

extern "C" int far3d_linear_mma(const float* x, const float* x_add, int ldx, const void* w_hi, const void* w_lo, const float* bias,
                                const float* residual, int ldr, float* y, int ldy, int M, int N, int K, int act, void* stream) {
    FAR3D_REQUIRE(x && w_hi && w_lo && y, "null pointer");
    FAR3D_REQUIRE(M > 0 && N > 0 && K > 0, "non-positive size");
    FAR3D_REQUIRE(K % LM_BK == 0, "K must be a multiple of 32");
    FAR3D_REQUIRE(ldx >= K && ldy >= N && ldx % 2 == 0, "bad row strides (ldx even, >= K; ldy >= N)");
    FAR3D_REQUIRE(((uintptr_t)x % 8 == 0) && (!x_add || (uintptr_t)x_add % 8 == 0), "x / x_add must be 8-byte aligned");
    FAR3D_REQUIRE(((uintptr_t)w_hi % 16 == 0) && ((uintptr_t)w_lo % 16 == 0), "weight planes must be 16-byte aligned");
    FAR3D_REQUIRE(!residual || ldr >= N, "residual row stride smaller than row");
    cudaStream_t st = (cudaStream_t)stream;
    const __half *wh = (const __half*)w_hi, *wl = (const __half*)w_lo;
    // 64-row CTAs when that already gives every SM a CTA, 32-row CTAs otherwise
    const long ctas64 = (long)cdiv(M, 64) * cdiv(N, LM_BN);
    if (ctas64 >= 120) {
        dim3 grid(cdiv(N, LM_BN), cdiv(M, 64));
        if (x_add) linear_mma_kernel<4, true><<<grid, 128, 0, st>>>(x, x_add, ldx, wh, wl, bias, residual, ldr, y, ldy, M, N, K, act);
        else linear_mma_kernel<4, false><<<grid, 128, 0, st>>>(x, x_add, ldx, wh, wl, bias, residual, ldr, y, ldy, M, N, K, act);
    } else {
        dim3 grid(cdiv(N, LM_BN), cdiv(M, 32));
        if (x_add) linear_mma_kernel<2, true><<<grid, 64, 0, st>>>(x, x_add, ldx, wh, wl, bias, residual, ldr, y, ldy, M, N, K, act);
        else linear_mma_kernel<2, false><<<grid, 64, 0, st>>>(x, x_add, ldx, wh, wl, bias, residual, ldr, y, ldy, M, N, K, act);
    }
    return launched("linear_mma_kernel");
}
