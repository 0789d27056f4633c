// Multi-head attention core softmax(Q K^T / sqrt(32)) V on the tensor cores, head dim 32, fp32-grade.
//
// Replaces the torch MHA core inside mmcv MultiheadAttention (models/utils/detr3d_transformer.py:311-422 -> torch
// nn.MultiheadAttention): ~1047 queries x ~1924 keys x 8 heads per decoder layer.  The SIMT kernel (mha_d32_kernel, decoder_ops.cu)
// is bound by the FMA pipe at 62-66 us per layer; this one runs both products as warp-level MMAs (mma.sync m16n8k16, fp16 operands,
// fp32 accumulate) with every operand split in two fp16 planes (x = hi + lo to 2^-22) and three MMAs per product
// (hi.hi + lo.hi + hi.lo) - the split-operand scheme of the convolutions.
//
// CTA = 64 queries of one (batch, head): KG key groups x 4 warps x 16 query rows.  A key group owns every KG-th key tile, its own
// double-buffered shared memory and its own named barrier, so KG warps per scheduler hide each other's MMA / shared-memory / MUFU
// latencies (with one key group the kernel ran at 0.23 IPC and was no faster than the SIMT one: 69.5 us); the groups' partial
// (max, sum, O) states are merged through shared memory at the end.  Keys stream in tiles of 64 through shared memory:
// the next tile's K / V rows are fetched into registers (fp32 from the projection GEMM) while the current tile is computed, then
// split and stored - K as [key][dim], V transposed [dim][key] - so that every B fragment is one 32-bit shared-memory load.
// Per tile and warp: S = Q K^T (8 n-tiles x 2 k-steps x 3 MMAs), online softmax on the accumulator fragments (rows g and g + 8
// of the thread, quad shuffles for the row maximum), P = exp(S - m) split into the A fragments of the second product directly
// from the accumulator registers (the C layout of two adjacent n-tiles is the A layout of one k-step), O += P V (4 x 4 x 3 MMAs).
#include "common.cuh"

namespace far3d {

constexpr int MM_QT = 64, MM_KT = 64, MM_WARPS = 4;
constexpr int MM_KLD = 40;      // halves per K row   (32 + 8 pad: the 8 rows a fragment load touches fall in distinct banks)
constexpr int MM_VLD = 72;      // halves per V^T row (64 + 8 pad)
constexpr int MM_K_PLANE = MM_KT * MM_KLD, MM_V_PLANE = 32 * MM_VLD;                 // halves per plane and buffer
constexpr int MM_GROUP_HALVES = 2 * 2 * (MM_K_PLANE + MM_V_PLANE);                    // per key group: 2 buffers x (hi, lo) x (K, V^T)
constexpr int MM_GROUP_BYTES = MM_GROUP_HALVES * 2;                                   // 38912

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// (x0, x1) -> packed fp16 pairs of the hi and lo planes (packed converts: 6 instructions per pair)
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(x0, x1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void group_barrier(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

template <int KG>
__global__ void __launch_bounds__(KG * MM_WARPS * 32)
mha_mma_d32_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k, int ldk, const float* __restrict__ v, int ldv,
                   float* __restrict__ o, int ldo, int B, int Nq, int Nk, int H, const int* __restrict__ key_skip) {
    extern __shared__ __align__(16) unsigned char mm_smem[];
    const int kgroup = threadIdx.x >> 7;                                   // key group of this warp
    __half* sm = reinterpret_cast<__half*>(mm_smem) + (size_t)kgroup * MM_GROUP_HALVES;
    // buffer b: K hi | K lo | V^T hi | V^T lo
    auto k_hi = [&](int b) { return sm + b * 2 * (MM_K_PLANE + MM_V_PLANE); };
    auto k_lo = [&](int b) { return k_hi(b) + MM_K_PLANE; };
    auto v_hi = [&](int b) { return k_hi(b) + 2 * MM_K_PLANE; };
    auto v_lo = [&](int b) { return v_hi(b) + MM_V_PLANE; };
    const int skip0 = key_skip ? __ldg(key_skip) : 0, skip1 = key_skip ? skip0 + __ldg(key_skip + 1) : 0;
    const int tid = threadIdx.x & 127, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;   // (within the key group)
    const int qtiles = (Nq + MM_QT - 1) / MM_QT;
    int bid = blockIdx.x;
    const int qt = bid % qtiles; bid /= qtiles;
    const int h = bid % H; const int b = bid / H;
    const float scale = 0.17677669529663687f;   // 1/sqrt(32)

    // ---- Q fragments (A of the first product): rows r0 = g, r1 = g + 8 of the warp's 16; k-step ks covers dims 16 ks .. 16 ks + 15
    const int row0 = qt * MM_QT + warp * 16 + g, row1 = row0 + 8;
    uint32_t qh[2][4], ql[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int half = 0; half < 2; ++half) {           // dims 16 ks + 2 t (+8)
            const int d = 16 * ks + 2 * t + 8 * half;
            float2 x0 = make_float2(0.f, 0.f), x1 = make_float2(0.f, 0.f);
            if (row0 < Nq) x0 = *reinterpret_cast<const float2*>(q + ((size_t)b * Nq + row0) * ldq + h * 32 + d);
            if (row1 < Nq) x1 = *reinterpret_cast<const float2*>(q + ((size_t)b * Nq + row1) * ldq + h * 32 + d);
            split2(x0.x * scale, x0.y * scale, qh[ks][2 * half], ql[ks][2 * half]);              // a0/a4: row g
            split2(x1.x * scale, x1.y * scale, qh[ks][2 * half + 1], ql[ks][2 * half + 1]);      // a2/a6 -> registers 1 / 3: row g + 8
        }
    // register order of the A fragment is {a01 (row g, k lo), a23 (row g+8, k lo), a45 (row g, k hi), a67 (row g+8, k hi)}: the loop
    // above filled [0] = (g, lo), [1] = (g+8, lo), [2] = (g, hi), [3] = (g+8, hi) - already that order

    float oacc[4][4];
#pragma unroll
    for (int nd = 0; nd < 4; ++nd)
#pragma unroll
        for (int i = 0; i < 4; ++i) oacc[nd][i] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;      // running maxima (rows g, g + 8) and this thread's partial sums

    // ---- tile loader: 64 keys x 32 dims of K and of V = 512 float4 each; thread handles pieces tid + 128 r
    float4 pk[4], pv[4];
    auto fetch = [&](int tile) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int pc = tid + 128 * r, key = tile * MM_KT + (pc >> 3), c4 = (pc & 7) * 4;
            if (key < Nk) {
                const size_t row = (size_t)b * Nk + key;
                pk[r] = __ldg(reinterpret_cast<const float4*>(k + row * ldk + h * 32 + c4));
                pv[r] = __ldg(reinterpret_cast<const float4*>(v + row * ldv + h * 32 + c4));
            } else {
                pk[r] = make_float4(0.f, 0.f, 0.f, 0.f); pv[r] = pk[r];
            }
        }
    };
    auto stash = [&](int buf) {
        __half *kh = k_hi(buf), *kl = k_lo(buf), *vh = v_hi(buf), *vl = v_lo(buf);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int pc = tid + 128 * r, key = pc >> 3, c4 = (pc & 7) * 4;
            uint32_t h01, l01, h23, l23;
            split2(pk[r].x, pk[r].y, h01, l01); split2(pk[r].z, pk[r].w, h23, l23);
            *reinterpret_cast<uint2*>(kh + key * MM_KLD + c4) = make_uint2(h01, h23);
            *reinterpret_cast<uint2*>(kl + key * MM_KLD + c4) = make_uint2(l01, l23);
            split2(pv[r].x, pv[r].y, h01, l01); split2(pv[r].z, pv[r].w, h23, l23);
            unsigned short* vhs = reinterpret_cast<unsigned short*>(vh);
            unsigned short* vls = reinterpret_cast<unsigned short*>(vl);
            vhs[(c4 + 0) * MM_VLD + key] = (unsigned short)(h01 & 0xFFFFu); vhs[(c4 + 1) * MM_VLD + key] = (unsigned short)(h01 >> 16);
            vhs[(c4 + 2) * MM_VLD + key] = (unsigned short)(h23 & 0xFFFFu); vhs[(c4 + 3) * MM_VLD + key] = (unsigned short)(h23 >> 16);
            vls[(c4 + 0) * MM_VLD + key] = (unsigned short)(l01 & 0xFFFFu); vls[(c4 + 1) * MM_VLD + key] = (unsigned short)(l01 >> 16);
            vls[(c4 + 2) * MM_VLD + key] = (unsigned short)(l23 & 0xFFFFu); vls[(c4 + 3) * MM_VLD + key] = (unsigned short)(l23 >> 16);
        }
    };

    const int ntiles = (Nk + MM_KT - 1) / MM_KT;
    if (kgroup < ntiles) { fetch(kgroup); stash(0); }
    group_barrier(1 + kgroup);
    int buf = 0;
    for (int tile = kgroup; tile < ntiles; tile += KG) {
        if (tile + KG < ntiles) fetch(tile + KG);        // global loads in flight while this tile is computed
        const int key0 = tile * MM_KT;
        const bool all_masked = key0 >= skip0 && key0 + MM_KT <= skip1;     // (group-uniform)
        const bool edge = key0 + MM_KT > Nk || (key0 < skip1 && key0 + MM_KT > skip0);   // tile holds keys that do not exist
        if (!all_masked) {
            const __half *kh = k_hi(buf), *kl = k_lo(buf), *vh = v_hi(buf), *vl = v_lo(buf);
            // ---- S = Q K^T: n-tile j = keys 8 j .. 8 j + 7; B fragment b0 = K[key 8j+g][dims 16ks+2t,+1], b1 = dims +8
            float s[8][4];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const int off = (8 * j + g) * MM_KLD + 16 * ks + 2 * t;
                    const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(kh + off), bh1 = *reinterpret_cast<const uint32_t*>(kh + off + 8);
                    const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(kl + off), bl1 = *reinterpret_cast<const uint32_t*>(kl + off + 8);
                    mma16816(s[j], ql[ks], bh0, bh1);
                    mma16816(s[j], qh[ks], bl0, bl1);
                    mma16816(s[j], qh[ks], bh0, bh1);
                }
            }
            // ---- mask + online softmax; accumulator element i of n-tile j: row (i < 2 ? g : g + 8), key 8 j + 2 t + (i & 1)
            float mx0 = -INFINITY, mx1 = -INFINITY;
            if (edge) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int key = key0 + 8 * j + 2 * t + (i & 1);
                        const bool ok = key < Nk && !(key >= skip0 && key < skip1);
                        if (!ok) s[j][i] = -INFINITY;
                    }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
                mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
            }
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
            const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
            // a row whose every key so far is masked keeps m = -inf: exponentials of -inf - (-inf) would be NaN, use 0 instead
            const float c0 = mn0 == -INFINITY ? 1.f : __expf(m0 - mn0), c1 = mn1 == -INFINITY ? 1.f : __expf(m1 - mn1);
            const float b0s = mn0 == -INFINITY ? 0.f : mn0, b1s = mn1 == -INFINITY ? 0.f : mn1;
            m0 = mn0; m1 = mn1;
            l0 *= c0; l1 *= c1;
#pragma unroll
            for (int nd = 0; nd < 4; ++nd) { oacc[nd][0] *= c0; oacc[nd][1] *= c0; oacc[nd][2] *= c1; oacc[nd][3] *= c1; }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                s[j][0] = __expf(s[j][0] - b0s); s[j][1] = __expf(s[j][1] - b0s);
                s[j][2] = __expf(s[j][2] - b1s); s[j][3] = __expf(s[j][3] - b1s);
                l0 += s[j][0] + s[j][1]; l1 += s[j][2] + s[j][3];
            }
            // ---- O += P V: k-step kk = keys 16 kk .. 16 kk + 15 = n-tiles 2 kk, 2 kk + 1 of S; n-tile nd = dims 8 nd .. 8 nd + 7
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                uint32_t ph[4], pl[4];
                split2(s[2 * kk][0], s[2 * kk][1], ph[0], pl[0]);               // row g,     keys 16 kk + 2 t, + 1
                split2(s[2 * kk][2], s[2 * kk][3], ph[1], pl[1]);               // row g + 8
                split2(s[2 * kk + 1][0], s[2 * kk + 1][1], ph[2], pl[2]);       // row g,     keys 16 kk + 8 + 2 t, + 1
                split2(s[2 * kk + 1][2], s[2 * kk + 1][3], ph[3], pl[3]);       // row g + 8
#pragma unroll
                for (int nd = 0; nd < 4; ++nd) {
                    const int off = (8 * nd + g) * MM_VLD + 16 * kk + 2 * t;    // V^T[dim 8 nd + g][keys 16 kk + 2 t, + 1]
                    const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(vh + off), bh1 = *reinterpret_cast<const uint32_t*>(vh + off + 8);
                    const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(vl + off), bl1 = *reinterpret_cast<const uint32_t*>(vl + off + 8);
                    mma16816(oacc[nd], pl, bh0, bh1);
                    mma16816(oacc[nd], ph, bl0, bl1);
                    mma16816(oacc[nd], ph, bh0, bh1);
                }
            }
        }
        if (tile + KG < ntiles) stash(buf ^ 1);          // the other buffer: every warp of the group finished reading it before the last barrier
        group_barrier(1 + kgroup);
        buf ^= 1;
    }
    // ---- row sums across the quad, normalise, store (row g: elements 0, 1; row g + 8: elements 2, 3; dims 8 nd + 2 t, + 1)
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    if (KG > 1) {
        // ---- merge the key groups' partial states (flash-decoding): groups 1.. publish (m, l, O) per thread, group 0 folds them in
        __syncthreads();                                 // every group is done with its tile buffers
        float* mg = reinterpret_cast<float*>(mm_smem);   // [KG - 1][128 threads][20]
        if (kgroup > 0) {
            float* dst = mg + ((size_t)(kgroup - 1) * 128 + tid) * 20;
            dst[0] = m0; dst[1] = m1; dst[2] = l0; dst[3] = l1;
#pragma unroll
            for (int nd = 0; nd < 4; ++nd)
#pragma unroll
                for (int i = 0; i < 4; ++i) dst[4 + 4 * nd + i] = oacc[nd][i];
        }
        __syncthreads();
        if (kgroup > 0) return;
#pragma unroll
        for (int kg = 1; kg < KG; ++kg) {
            const float* src = mg + ((size_t)(kg - 1) * 128 + tid) * 20;
            const float pm0 = src[0], pm1 = src[1];
            const float M0 = fmaxf(m0, pm0), M1 = fmaxf(m1, pm1);
            // a side that saw no key at all has m = -inf and l = 0, O = 0: its factor is irrelevant, keep it finite
            const float a0 = m0 == -INFINITY ? 0.f : __expf(m0 - M0), b0 = pm0 == -INFINITY ? 0.f : __expf(pm0 - M0);
            const float a1 = m1 == -INFINITY ? 0.f : __expf(m1 - M1), b1 = pm1 == -INFINITY ? 0.f : __expf(pm1 - M1);
            l0 = l0 * a0 + src[2] * b0; l1 = l1 * a1 + src[3] * b1;
#pragma unroll
            for (int nd = 0; nd < 4; ++nd) {
                oacc[nd][0] = oacc[nd][0] * a0 + src[4 + 4 * nd + 0] * b0; oacc[nd][1] = oacc[nd][1] * a0 + src[4 + 4 * nd + 1] * b0;
                oacc[nd][2] = oacc[nd][2] * a1 + src[4 + 4 * nd + 2] * b1; oacc[nd][3] = oacc[nd][3] * a1 + src[4 + 4 * nd + 3] * b1;
            }
            m0 = M0; m1 = M1;
        }
    }
    const float i0 = 1.f / l0, i1 = 1.f / l1;
#pragma unroll
    for (int nd = 0; nd < 4; ++nd) {
        const int d = h * 32 + 8 * nd + 2 * t;
        if (row0 < Nq) *reinterpret_cast<float2*>(o + ((size_t)b * Nq + row0) * ldo + d) = make_float2(oacc[nd][0] * i0, oacc[nd][1] * i0);
        if (row1 < Nq) *reinterpret_cast<float2*>(o + ((size_t)b * Nq + row1) * ldo + d) = make_float2(oacc[nd][2] * i1, oacc[nd][3] * i1);
    }
}

}  // namespace far3d

using namespace far3d;

static int g_mha_kg = 3;
// tools: key groups per CTA (1..4) of the tensor-core form
void far3d_mha_mma_set_key_groups(int kg) { g_mha_kg = kg < 1 ? 1 : kg > 4 ? 4 : kg; }

// called by mha_impl (decoder_ops.cu) when the tensor-core form is selected
int far3d_mha_mma_launch(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* o, int ldo, int B, int Nq,
                         int Nk, int H, const int* key_skip, void* stream) {
    const int qtiles = cdiv(Nq, MM_QT);
    const int kg = g_mha_kg;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(mha_mma_d32_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * MM_GROUP_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(mha_mma_d32_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * MM_GROUP_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(mha_mma_d32_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * MM_GROUP_BYTES) != cudaSuccess)
            return fail(FAR3D_E_CUDA, "%smha: cannot opt in to %ld bytes of shared memory", "", (long)(4 * MM_GROUP_BYTES));
        attr_set = true;
    }
#define FAR3D_MHA_LAUNCH(KG)                                                                                              \
    mha_mma_d32_kernel<KG><<<B * H * qtiles, KG * MM_WARPS * 32, KG * MM_GROUP_BYTES, (cudaStream_t)stream>>>(            \
        q, ldq, k, ldk, v, ldv, o, ldo, B, Nq, Nk, H, key_skip)
    if (kg == 1) FAR3D_MHA_LAUNCH(1);
    else if (kg == 2) FAR3D_MHA_LAUNCH(2);
    else if (kg == 4) FAR3D_MHA_LAUNCH(4);
    else FAR3D_MHA_LAUNCH(3);
#undef FAR3D_MHA_LAUNCH
    return launched("mha_mma_d32_kernel");
}
