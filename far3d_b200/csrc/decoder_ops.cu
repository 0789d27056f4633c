// Decoder-side kernels: fp32 SIMT GEMM (nn.Linear), LayerNorm, multi-head attention core, position encodings,
// MLN.  Reference call sites are listed in include/far3d_b200.h next to each entry point.
#include "common.cuh"

namespace far3d {

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

// ------------------------------------------------------------------------------------------ SGEMM  y = x w^T
// 128x64 tile, BK 16, 256 threads, 8x4 micro-tile per thread. x [M,K] (ldx), w [N,K] (K contiguous).
constexpr int GB_M = 128, GB_N = 64, GB_K = 16;

template <int TM>       // rows per CTA tile: 128 (8 per thread) or 32 (2 per thread, for grids that would not fill the GPU)
__global__ void __launch_bounds__(256)
linear_f32_kernel(const float* __restrict__ x, const float* __restrict__ x_add, int ldx, const float* __restrict__ w,
                  const float* __restrict__ bias,
                  const float* __restrict__ residual, int ldr, float* __restrict__ y, int ldy, int M, int N, int K,
                  int act, int vec_ok) {
    constexpr int RPT = TM / 16;
    __shared__ __align__(16) float As[GB_K][TM + 4];
    __shared__ __align__(16) float Bs[GB_K][GB_N + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * GB_N;
    const int ty = tid / 16, tx = tid % 16;          // 16 x 16 threads -> rows ty*RPT.., cols tx*4..
    float acc[RPT][4];
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    // loaders: A tile 128x16 = 512 float4 -> 2 per thread; B tile 64x16 = 256 float4 -> 1 per thread
    const int la_r = tid / 4, la_c = (tid % 4) * 4;
    for (int k0 = 0; k0 < K; k0 += GB_K) {
#pragma unroll
        for (int h = 0; h < (TM + 63) / 64; ++h) {
            int r = la_r + h * 64;
            if (r >= TM) break;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            int gm = m0 + r, gk = k0 + la_c;
            if (gm < M) {
                if (vec_ok && gk + 3 < K) {
                    v = *reinterpret_cast<const float4*>(x + (size_t)gm * ldx + gk);
                    if (x_add) {
                        float4 a = *reinterpret_cast<const float4*>(x_add + (size_t)gm * ldx + gk);
                        v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
                    }
                } else {
                    float t[4] = {0.f, 0.f, 0.f, 0.f};
                    for (int i = 0; i < 4; ++i)
                        if (gk + i < K) t[i] = x[(size_t)gm * ldx + gk + i] + (x_add ? x_add[(size_t)gm * ldx + gk + i] : 0.f);
                    v = make_float4(t[0], t[1], t[2], t[3]);
                }
            }
            As[la_c + 0][r] = v.x; As[la_c + 1][r] = v.y; As[la_c + 2][r] = v.z; As[la_c + 3][r] = v.w;
        }
        {
            int r = la_r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            int gn = n0 + r, gk = k0 + la_c;
            if (gn < N) {
                if (vec_ok && gk + 3 < K) v = *reinterpret_cast<const float4*>(w + (size_t)gn * K + gk);
                else {
                    float t[4] = {0.f, 0.f, 0.f, 0.f};
                    for (int i = 0; i < 4; ++i) if (gk + i < K) t[i] = w[(size_t)gn * K + gk + i];
                    v = make_float4(t[0], t[1], t[2], t[3]);
                }
            }
            Bs[la_c + 0][r] = v.x; Bs[la_c + 1][r] = v.y; Bs[la_c + 2][r] = v.z; Bs[la_c + 3][r] = v.w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GB_K; ++k) {
            float a[RPT], b[4];
#pragma unroll
            for (int i = 0; i < RPT; ++i) a[i] = As[k][ty * RPT + i];
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
#pragma unroll
            for (int i = 0; i < RPT; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        int gm = m0 + ty * RPT + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            float v = acc[i][j] + (bias ? bias[gn] : 0.f);
            if (act == 1) v = fmaxf(v, 0.f);
            if (residual) v += residual[(size_t)gm * ldr + gn];
            y[(size_t)gm * ldy + gn] = v;
        }
    }
}

// N <= 64 columns and K <= 512 (the 256 -> 39 key-point offsets, the 256 -> 10 / 26 branch heads): one warp per output ROW, the
// row (+ x_add) held in registers, the weight rows streamed through L1 with coalesced float4 reads, one shuffle reduction per
// column.  The tiled kernel below runs such shapes as <= 29 CTAs with a barrier per 16-wide k step: 17 us for 900 x 256 -> 39.
constexpr int LRW_MAXK4 = 4;                         // float4 per lane: K <= 512
__global__ void __launch_bounds__(256)
linear_f32_rowwarp_kernel(const float* __restrict__ x, const float* __restrict__ x_add, int ldx, const float* __restrict__ w,
                          const float* __restrict__ bias, const float* __restrict__ residual, int ldr,
                          float* __restrict__ y, int ldy, int M, int N, int K, int act) {
    const int m = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (m >= M) return;
    float4 xv[LRW_MAXK4];
#pragma unroll
    for (int j = 0; j < LRW_MAXK4; ++j) {
        const int k = lane * 4 + 128 * j;
        xv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < K) {
            xv[j] = *reinterpret_cast<const float4*>(x + (size_t)m * ldx + k);
            if (x_add) {
                const float4 a = *reinterpret_cast<const float4*>(x_add + (size_t)m * ldx + k);
                xv[j].x += a.x; xv[j].y += a.y; xv[j].z += a.z; xv[j].w += a.w;
            }
        }
    }
    float mine = 0.f;                                // lane n keeps column n (n < 32), lane n - 32 column n of the second half
    for (int n0 = 0; n0 < N; n0 += 32) {
        for (int n = n0; n < min(N, n0 + 32); ++n) {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < LRW_MAXK4; ++j) {
                const int k = lane * 4 + 128 * j;
                if (k < K) {
                    const float4 wv = __ldg(reinterpret_cast<const float4*>(w + (size_t)n * K + k));
                    acc = fmaf(xv[j].x, wv.x, fmaf(xv[j].y, wv.y, fmaf(xv[j].z, wv.z, fmaf(xv[j].w, wv.w, acc))));
                }
            }
            acc = warp_sum(acc);
            if (lane == n - n0) mine = acc;
        }
        const int n = n0 + lane;
        if (n < N) {
            float v = mine + (bias ? __ldg(bias + n) : 0.f);
            if (act == 1) v = fmaxf(v, 0.f);
            if (residual) v += residual[(size_t)m * ldr + n];
            y[(size_t)m * ldy + n] = v;
        }
    }
}

// M <= 8 rows: one warp per output column n, lanes split K in float4s (coalesced weight-row reads), shuffle reduction.
constexpr int LSK_MAXM = 8;
__global__ void __launch_bounds__(256)
linear_f32_skinny_kernel(const float* __restrict__ x, const float* __restrict__ x_add, int ldx, const float* __restrict__ w,
                         const float* __restrict__ bias, const float* __restrict__ residual, int ldr,
                         float* __restrict__ y, int ldy, int M, int N, int K, int act) {
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= N) return;
    float acc[LSK_MAXM];
#pragma unroll
    for (int m = 0; m < LSK_MAXM; ++m) acc[m] = 0.f;
    for (int k = lane * 4; k < K; k += 128) {
        const float4 wv = *reinterpret_cast<const float4*>(w + (size_t)n * K + k);
#pragma unroll
        for (int m = 0; m < LSK_MAXM; ++m) {
            if (m < M) {
                float4 xv = *reinterpret_cast<const float4*>(x + (size_t)m * ldx + k);
                if (x_add) {
                    const float4 a = *reinterpret_cast<const float4*>(x_add + (size_t)m * ldx + k);
                    xv.x += a.x; xv.y += a.y; xv.z += a.z; xv.w += a.w;
                }
                acc[m] = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, fmaf(xv.z, wv.z, fmaf(xv.w, wv.w, acc[m]))));
            }
        }
    }
#pragma unroll
    for (int m = 0; m < LSK_MAXM; ++m) acc[m] = warp_sum(acc[m]);
    if (lane == 0) {
        const float bv = bias ? bias[n] : 0.f;
#pragma unroll
        for (int m = 0; m < LSK_MAXM; ++m) {
            if (m < M) {
                float v = acc[m] + bv;
                if (act == 1) v = fmaxf(v, 0.f);
                if (residual) v += residual[(size_t)m * ldr + n];
                y[(size_t)m * ldy + n] = v;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ LayerNorm
// warp per row, C <= 1024
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ add, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float* __restrict__ y, int M, int C, float eps, int relu_before,
                 int relu_after) {
    int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= M) return;
    int lane = threadIdx.x & 31;
    float v[32];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        int c = lane + i * 32;
        float t = 0.f;
        if (c < C) {
            t = x[(size_t)row * C + c];
            if (add) t += add[(size_t)row * C + c];
            if (relu_before) t = fmaxf(t, 0.f);
        }
        v[i] = t; s += t;
    }
    float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        int c = lane + i * 32;
        if (c < C) { float d = v[i] - mean; q += d * d; }
    }
    float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        int c = lane + i * 32;
        if (c < C) {
            float t = (v[i] - mean) * rstd;
            if (gamma) t = t * gamma[c] + (beta ? beta[c] : 0.f);
            if (relu_after) t = fmaxf(t, 0.f);
            y[(size_t)row * C + c] = t;
        }
    }
}

// ------------------------------------------------------------------------------------------ MHA core, Dh = 32
// One CTA = 64 queries of one (batch, head): every lane owns QPL = 2 queries (q rows, output accumulators and online-softmax
// states in registers), the 8 warps split the KEYS (16-key tiles, round-robin).  A warp streams its tiles through a private
// double-buffered smem slot with cp.async; every lane then reads the same K / V row (broadcast LDS.128), and each row read
// feeds both queries: the lane-per-key layout of the first kernel was shared-memory bound, the one-query-per-lane version
// LSU-return bound.  The 8 partial (m, l, acc) states per query are merged through smem at the end (flash-decoding style).
constexpr int MHA_WARPS = 8, MHA_TK = 16, MHA_QPL = 2, MHA_QT = 32 * MHA_QPL;
constexpr int MHA_SLOT = 2 * MHA_TK * 32;                     // floats per buffer: K tile then V tile
constexpr int MHA_MERGE = MHA_WARPS * 32 * 33 + 2 * MHA_WARPS * 32;            // floats: partial acc (stride 33) + m + l
constexpr int MHA_SMEM = (MHA_WARPS * 2 * MHA_SLOT > MHA_MERGE ? MHA_WARPS * 2 * MHA_SLOT : MHA_MERGE) * 4;   // 64 KB

__device__ __forceinline__ void cp_async16(float* dst, const float* src, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    const int n = valid ? 16 : 0;                                                   // src-size 0 -> zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}

__global__ void __launch_bounds__(MHA_WARPS * 32, 1)
mha_d32_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k, int ldk, const float* __restrict__ v,
               int ldv, float* __restrict__ o, int ldo, int B, int Nq, int Nk, int H, const int* __restrict__ key_skip) {
    extern __shared__ __align__(16) float mha_smem[];
    // keys [skip0, skip1) do not exist (padding rows of a bucketed query count; the range is data dependent, so it lives in
    // device memory and a captured graph replays with whatever the proposal kernels wrote there)
    const int skip0 = key_skip ? __ldg(key_skip) : 0, skip1 = key_skip ? skip0 + __ldg(key_skip + 1) : 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int qtiles = (Nq + MHA_QT - 1) / MHA_QT;
    int bid = blockIdx.x;
    const int qt = bid % qtiles; bid /= qtiles;
    const int h = bid % H; const int b = bid / H;
    const float scale = 0.17677669529663687f;   // 1/sqrt(32)
    // packed fp32 pairs (FFMA2, sm_100): (d, d+1) of q / k / v / acc ride in one 64-bit register pair, so the FMA count
    // of the two inner products halves; rounding per element is the ordinary fma.rn
    float2 qr[MHA_QPL][16], acc[MHA_QPL][16];
    float m[MHA_QPL], l[MHA_QPL];
#pragma unroll
    for (int u = 0; u < MHA_QPL; ++u) {
        const int qi = qt * MHA_QT + u * 32 + lane;
        const bool active = qi < Nq;
        const float4* qp = reinterpret_cast<const float4*>(q + ((size_t)b * Nq + (active ? qi : 0)) * ldq + h * 32);
#pragma unroll
        for (int d4 = 0; d4 < 8; ++d4) {
            float4 t = active ? qp[d4] : make_float4(0.f, 0.f, 0.f, 0.f);
            qr[u][2 * d4] = make_float2(t.x * scale, t.y * scale); qr[u][2 * d4 + 1] = make_float2(t.z * scale, t.w * scale);
        }
        m[u] = -INFINITY; l[u] = 0.f;
#pragma unroll
        for (int d = 0; d < 16; ++d) acc[u][d] = make_float2(0.f, 0.f);
    }

    float* slot = mha_smem + warp * 2 * MHA_SLOT;
    const int ntiles = (Nk + MHA_TK - 1) / MHA_TK;
    auto issue = [&](int t, int buf) {                       // 16 keys x 128 B of K and of V: 4 + 4 16-byte pieces per lane
        float* dst = slot + buf * MHA_SLOT;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int pc = lane + 32 * j, r = pc >> 3, c4 = (pc & 7) * 4;
            const int key = t * MHA_TK + r;
            const bool ok = key < Nk;
            const size_t row = (size_t)b * Nk + (ok ? key : 0);
            cp_async16(dst + r * 32 + c4, k + row * ldk + h * 32 + c4, ok);
            cp_async16(dst + MHA_TK * 32 + r * 32 + c4, v + row * ldv + h * 32 + c4, ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int buf = 0;
    if (warp < ntiles) issue(warp, 0);
    for (int t = warp; t < ntiles; t += MHA_WARPS) {
        if (t + MHA_WARPS < ntiles) {
            issue(t + MHA_WARPS, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();
        const float* kt = slot + buf * MHA_SLOT;
        const float* vt = kt + MHA_TK * 32;
#pragma unroll
        for (int c = 0; c < MHA_TK / 8; ++c) {
            if (t * MHA_TK + c * 8 >= Nk) break;             // chunk entirely past Nk (warp-uniform)
            if (t * MHA_TK + c * 8 >= skip0 && t * MHA_TK + c * 8 + 8 <= skip1) continue;   // chunk entirely masked (warp-uniform)
            float s[MHA_QPL][8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4* kr = reinterpret_cast<const float4*>(kt + (c * 8 + j) * 32);
                float2 s0[MHA_QPL], s1[MHA_QPL];       // four partial sums per query: (d, d+1) lanes of two independent chains
#pragma unroll
                for (int u = 0; u < MHA_QPL; ++u) { s0[u] = make_float2(0.f, 0.f); s1[u] = make_float2(0.f, 0.f); }
#pragma unroll
                for (int d4 = 0; d4 < 8; ++d4) {
                    const float4 kk = kr[d4];
                    const float2 k01 = make_float2(kk.x, kk.y), k23 = make_float2(kk.z, kk.w);
#pragma unroll
                    for (int u = 0; u < MHA_QPL; ++u) {
                        s0[u] = __ffma2_rn(qr[u][2 * d4], k01, s0[u]);
                        s1[u] = __ffma2_rn(qr[u][2 * d4 + 1], k23, s1[u]);
                    }
                }
                const int key = t * MHA_TK + c * 8 + j;
                const bool ok = key < Nk && !(key >= skip0 && key < skip1);
#pragma unroll
                for (int u = 0; u < MHA_QPL; ++u) s[u][j] = ok ? (s0[u].x + s0[u].y) + (s1[u].x + s1[u].y) : -INFINITY;
            }
            float corr[MHA_QPL];
#pragma unroll
            for (int u = 0; u < MHA_QPL; ++u) {
                float mx = s[u][0];
#pragma unroll
                for (int j = 1; j < 8; ++j) mx = fmaxf(mx, s[u][j]);
                const float mn = fmaxf(m[u], mx);            // finite: at least one key of this chunk is valid
                corr[u] = __expf(m[u] - mn);
                m[u] = mn;
                l[u] *= corr[u];
                const float2 c2 = make_float2(corr[u], corr[u]);
#pragma unroll
                for (int d = 0; d < 16; ++d) acc[u][d] = __fmul2_rn(acc[u][d], c2);
#pragma unroll
                for (int j = 0; j < 8; ++j) { s[u][j] = __expf(s[u][j] - mn); l[u] += s[u][j]; }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4* vr = reinterpret_cast<const float4*>(vt + (c * 8 + j) * 32);
#pragma unroll
                for (int d4 = 0; d4 < 8; ++d4) {
                    const float4 vv = vr[d4];
                    const float2 v01 = make_float2(vv.x, vv.y), v23 = make_float2(vv.z, vv.w);
#pragma unroll
                    for (int u = 0; u < MHA_QPL; ++u) {
                        const float2 p2 = make_float2(s[u][j], s[u][j]);
                        acc[u][2 * d4] = __ffma2_rn(p2, v01, acc[u][2 * d4]);
                        acc[u][2 * d4 + 1] = __ffma2_rn(p2, v23, acc[u][2 * d4 + 1]);
                    }
                }
            }
        }
        __syncwarp();
        buf ^= 1;
    }
    // merge the 8 per-warp states of each query through smem, one query group (32 queries) at a time:
    // part[w][lane][33] (stride 33: conflict-free), pm / pl [w][lane]
    float* part = mha_smem;
    float* pm = mha_smem + MHA_WARPS * 32 * 33;
    float* pl = pm + MHA_WARPS * 32;
#pragma unroll
    for (int u = 0; u < MHA_QPL; ++u) {
        __syncthreads();                                     // tile buffers / the previous group's partials are no longer read
#pragma unroll
        for (int d = 0; d < 16; ++d) {
            part[(warp * 32 + lane) * 33 + 2 * d] = acc[u][d].x;
            part[(warp * 32 + lane) * 33 + 2 * d + 1] = acc[u][d].y;
        }
        pm[warp * 32 + lane] = m[u]; pl[warp * 32 + lane] = l[u];
        __syncthreads();
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < MHA_WARPS; ++w) M = fmaxf(M, pm[w * 32 + lane]);
        float lt = 0.f, r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
#pragma unroll
        for (int w = 0; w < MHA_WARPS; ++w) {
            const float mw = pm[w * 32 + lane];
            const float f = (mw == -INFINITY) ? 0.f : __expf(mw - M);
            lt = fmaf(pl[w * 32 + lane], f, lt);
            const float* pr = part + (w * 32 + lane) * 33 + warp * 4;
            r0 = fmaf(pr[0], f, r0); r1 = fmaf(pr[1], f, r1); r2 = fmaf(pr[2], f, r2); r3 = fmaf(pr[3], f, r3);
        }
        const int qi = qt * MHA_QT + u * 32 + lane;
        if (qi < Nq) {
            const float inv = 1.f / lt;
            *reinterpret_cast<float4*>(o + ((size_t)b * Nq + qi) * ldo + h * 32 + warp * 4) =
                make_float4(r0 * inv, r1 * inv, r2 * inv, r3 * inv);
        }
    }
}

// ------------------------------------------------------------------------------------------ position encodings
__global__ void pos2posemb3d_kernel(const float* __restrict__ pos, float* __restrict__ emb, int M, int F) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)M * 3 * F) return;
    int c = (int)(idx % (3 * F)); int m = (int)(idx / (3 * F));
    int blk = c / F, i = c % F;
    int axis = blk == 0 ? 1 : (blk == 1 ? 0 : 2);                 // (y, x, z), positional_encoding.py:24
    float p = pos[(size_t)m * 3 + axis] * 6.283185307179586f;
    float dim_t = powf(10000.f, (float)(2 * (i / 2)) / (float)F);
    float e = p / dim_t;
    emb[idx] = (i & 1) ? cosf(e) : sinf(e);
}

__global__ void pos2posemb1d_kernel(const float* __restrict__ pos, int ldp, float* __restrict__ emb, int M, int F) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)M * F) return;
    int i = (int)(idx % F); int m = (int)(idx / F);
    float p = pos[(size_t)m * ldp] * 6.283185307179586f;
    float dim_t = powf(10000.f, (float)(2 * (i / 2)) / (float)F);
    float e = p / dim_t;
    emb[idx] = (i & 1) ? cosf(e) : sinf(e);
}

__global__ void nerf_posenc_kernel(const float* __restrict__ x, float* __restrict__ emb, int M, int Cin, int nfreq) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    int W = Cin * 2 * nfreq;
    if (idx >= (long)M * W) return;
    int c = (int)(idx % W); int m = (int)(idx / W);
    int f = c / (2 * Cin), r = c % (2 * Cin);
    int is_cos = r / Cin, ch = r % Cin;
    float t = x[(size_t)m * Cin + ch] * exp2f((float)f);
    emb[idx] = is_cos ? cosf(t) : sinf(t);
}

// ------------------------------------------------------------------------------------------ MLN
// channels-last: pure elementwise (float4). out[bn, start+hw, c] = gamma[bn,c]*x[bn,hw,c] + beta[bn,c]
__global__ void mln_flatten_nhwc_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                        const float* __restrict__ beta, float* __restrict__ out, int BN, int HW, int C,
                                        int S, int start) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;      // float4 index
    int C4 = C / 4;
    if (idx >= (long)BN * HW * C4) return;
    int c4 = (int)(idx % C4); long r = idx / C4;
    int hw = (int)(r % HW); int bn = (int)(r / HW);
    float4 v = ldg_f4(x + ((size_t)bn * HW + hw) * C + c4 * 4);
    float4 g = ldg_f4(gamma + (size_t)bn * C + c4 * 4), b = ldg_f4(beta + (size_t)bn * C + c4 * 4);
    float4 o = make_float4(g.x * v.x + b.x, g.y * v.y + b.y, g.z * v.z + b.z, g.w * v.w + b.w);
    *reinterpret_cast<float4*>(out + ((size_t)bn * S + start + hw) * C + c4 * 4) = o;
}

// NCHW input: 32x32 smem transpose tile. grid (ceil(HW/32), ceil(C/32), BN), block (32, 8)
__global__ void mln_flatten_nchw_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                        const float* __restrict__ beta, float* __restrict__ out, int BN, int HW, int C,
                                        int S, int start) {
    __shared__ float tile[32][33];
    int bn = blockIdx.z, hw0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        int c = c0 + i, hw = hw0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && hw < HW) ? x[((size_t)bn * C + c) * HW + hw] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        int hw = hw0 + i, c = c0 + threadIdx.x;
        if (hw < HW && c < C)
            out[((size_t)bn * S + start + hw) * C + c] = gamma[(size_t)bn * C + c] * tile[threadIdx.x][i] + beta[(size_t)bn * C + c];
    }
}

// tokens: warp per row; optional parameter-free LayerNorm (eps 1e-5) of x first (misc.py:171-190)
__global__ void __launch_bounds__(256)
mln_tokens_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                  float* __restrict__ out, int M, int C, int use_ln) {
    int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= M) return;
    int lane = threadIdx.x & 31;
    float v[32]; float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) { int c = lane + i * 32; v[i] = c < C ? x[(size_t)row * C + c] : 0.f; s += v[i]; }
    float mean = 0.f, rstd = 1.f;
    if (use_ln) {
        mean = warp_sum(s) / (float)C;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) { int c = lane + i * 32; if (c < C) { float d = v[i] - mean; q += d * d; } }
        rstd = rsqrtf(warp_sum(q) / (float)C + 1e-5f);
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        int c = lane + i * 32;
        if (c < C) out[(size_t)row * C + c] = gamma[(size_t)row * C + c] * ((v[i] - mean) * rstd) + beta[(size_t)row * C + c];
    }
}


// ---- camera-embedding logits of ALL decoder layers in one launch.
// detr3d_transformer.py:530-540: cam_embed(lidar2img[..., :3, :].flatten(-2)) = LN(relu(W1 relu(W0 x + b0) + b1)), and - weights_fc
// being linear - its contribution to the aggregation logits is wc = W_fc cam (no bias; the query side carries it).  The input is
// the frame's lidar2img, the same for every layer, so the six layers' 7-row MLPs (5 launches each, the 12-wide GEMM alone 20 us
// on one SM) are one grid (camera row, layer) before the decoder loop.
struct CamLayer { const float *w0, *b0, *w1, *b1, *g, *beta, *wfc; };
struct CamParams { CamLayer l[FAR3D_MAX_CAM_LAYERS]; int layers, rows, E, H, J; float eps; };

__global__ void __launch_bounds__(256)
cam_logits_kernel(CamParams p, const float* __restrict__ lidar2img, float* __restrict__ out) {
    extern __shared__ float cl_smem[];                 // h1[H] | h2[E]
    __shared__ float s_x[12], s_red[8], s_stat[2];
    float* h1 = cl_smem;
    float* h2 = cl_smem + p.H;
    const int row = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const CamLayer L = p.l[blockIdx.y];
    if (tid < 12) s_x[tid] = lidar2img[(size_t)row * 16 + tid];
    __syncthreads();
    for (int j = tid; j < p.H; j += 256) {
        float a = __ldg(L.b0 + j);
#pragma unroll
        for (int k = 0; k < 12; ++k) a = fmaf(__ldg(L.w0 + j * 12 + k), s_x[k], a);
        h1[j] = fmaxf(a, 0.f);
    }
    __syncthreads();
    for (int j = warp; j < p.E; j += 8) {              // warp per output row, lanes across K (coalesced weight reads)
        float a = 0.f;
        for (int k = lane; k < p.H; k += 32) a = fmaf(__ldg(L.w1 + (size_t)j * p.H + k), h1[k], a);
        a = warp_sum(a);
        if (lane == 0) h2[j] = fmaxf(a + __ldg(L.b1 + j), 0.f);
    }
    __syncthreads();
    // LayerNorm over E (two passes: mean, then the variance of the centred values, as torch does)
    float s = 0.f;
    for (int j = tid; j < p.E; j += 256) s += h2[j];
    s = warp_sum(s);
    if (lane == 0) s_red[warp] = s;
    __syncthreads();
    if (tid == 0) { float t = 0.f; for (int w = 0; w < 8; ++w) t += s_red[w]; s_stat[0] = t / (float)p.E; }
    __syncthreads();
    const float mean = s_stat[0];
    float v = 0.f;
    for (int j = tid; j < p.E; j += 256) { const float d = h2[j] - mean; v = fmaf(d, d, v); }
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    if (tid == 0) { float t = 0.f; for (int w = 0; w < 8; ++w) t += s_red[w]; s_stat[1] = rsqrtf(t / (float)p.E + p.eps); }
    __syncthreads();
    const float rstd = s_stat[1];
    for (int j = tid; j < p.E; j += 256) h2[j] = (h2[j] - mean) * rstd * __ldg(L.g + j) + __ldg(L.beta + j);
    __syncthreads();
    float* o = out + ((size_t)blockIdx.y * p.rows + row) * p.J;
    for (int j = warp; j < p.J; j += 8) {
        float a = 0.f;
        for (int k = lane; k < p.E; k += 32) a = fmaf(__ldg(L.wfc + (size_t)j * p.E + k), h2[k], a);
        a = warp_sum(a);
        if (lane == 0) o[j] = a;
    }
}

}  // namespace far3d

using namespace far3d;

extern "C" const char* far3d_last_error(void) { return g_err; }
extern "C" int far3d_abi_version(void) { return 1; }
extern "C" int64_t far3d_launch_count(void) { return g_launches.load(); }
// kernels launched on our behalf by a CUDA-graph replay (the graph was captured from n of our launches)
extern "C" void far3d_add_launches(int64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" int far3d_linear_f32(const float* x, const float* x_add, int ldx, const float* w, const float* bias,
                                const float* residual, int ldr, float* y, int ldy, int M, int N, int K, int act,
                                void* stream) {
    FAR3D_REQUIRE(x && w && y, "null pointer");
    FAR3D_REQUIRE(M > 0 && N > 0 && K > 0, "non-positive size");
    FAR3D_REQUIRE(ldx >= K && ldy >= N, "row stride smaller than row");
    // 128-bit operand loads when everything is 16-byte aligned, scalar loads otherwise (e.g. the 14-wide MLN input)
    const int vec_ok = ((uintptr_t)x % 16 == 0) && ((uintptr_t)w % 16 == 0) && ldx % 4 == 0 && K % 4 == 0 &&
                       (!x_add || (uintptr_t)x_add % 16 == 0);
    if (M <= LSK_MAXM && vec_ok) {       // a handful of rows (camera embeddings): warp per output column, lanes split K
        linear_f32_skinny_kernel<<<cdiv(N, 8), 256, 0, (cudaStream_t)stream>>>(x, x_add, ldx, w, bias, residual, ldr, y, ldy,
                                                                              M, N, K, act);
        return launched("linear_f32_skinny_kernel");
    }
    if (N <= 64 && K <= 128 * LRW_MAXK4 && vec_ok && M >= 64) {     // few columns: warp per row
        linear_f32_rowwarp_kernel<<<cdiv(M, 8), 256, 0, (cudaStream_t)stream>>>(x, x_add, ldx, w, bias, residual, ldr, y, ldy,
                                                                               M, N, K, act);
        return launched("linear_f32_rowwarp_kernel");
    }
    if ((long)cdiv(N, GB_N) * cdiv(M, GB_M) < 64) {
        dim3 grid(cdiv(N, GB_N), cdiv(M, 32));
        linear_f32_kernel<32><<<grid, 256, 0, (cudaStream_t)stream>>>(x, x_add, ldx, w, bias, residual, ldr, y, ldy, M, N, K,
                                                                     act, vec_ok);
    } else {
        dim3 grid(cdiv(N, GB_N), cdiv(M, GB_M));
        linear_f32_kernel<GB_M><<<grid, 256, 0, (cudaStream_t)stream>>>(x, x_add, ldx, w, bias, residual, ldr, y, ldy, M, N,
                                                                       K, act, vec_ok);
    }
    return launched("linear_f32_kernel");
}

extern "C" int far3d_layernorm(const float* x, const float* add, const float* gamma, const float* beta, float* y, int M,
                               int C, float eps, int relu_before, int relu_after, void* stream) {
    FAR3D_REQUIRE(x && y, "null pointer");
    FAR3D_REQUIRE(M > 0 && C > 0, "non-positive size");
    if (C > 1024) return fail(FAR3D_E_UNSUPPORTED, "%slayernorm supports C <= 1024 (got %ld)", "", C);
    layernorm_kernel<<<cdiv((long)M * 32, 256), 256, 0, (cudaStream_t)stream>>>(x, add, gamma, beta, y, M, C, eps,
                                                                              relu_before, relu_after);
    return launched("layernorm_kernel");
}

static int mha_impl(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* o, int ldo, int B, int Nq,
                    int Nk, int H, int Dh, const int* key_skip, void* stream);
// tensor-core form (mha_mma.cu)
int far3d_mha_mma_launch(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* o, int ldo, int B, int Nq,
                         int Nk, int H, const int* key_skip, void* stream);
void far3d_mha_mma_set_key_groups(int kg);
static int g_mha_simt = 0;
// bit 0: SIMT kernel; bits 4-7: key groups per CTA of the tensor-core kernel (0: default)
extern "C" void far3d_mha_tune(int simt) { g_mha_simt = (simt & 1) ? 1 : 0; far3d_mha_mma_set_key_groups((simt >> 4) & 15 ? (simt >> 4) & 15 : 3); }
extern "C" int far3d_mha_fwd(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* o,
                             int ldo, int B, int Nq, int Nk, int H, int Dh, void* stream) {
    return mha_impl(q, ldq, k, ldk, v, ldv, o, ldo, B, Nq, Nk, H, Dh, nullptr, stream);
}
extern "C" int far3d_mha_fwd_masked(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* o,
                                    int ldo, int B, int Nq, int Nk, int H, int Dh, const int32_t* key_skip, void* stream) {
    return mha_impl(q, ldq, k, ldk, v, ldv, o, ldo, B, Nq, Nk, H, Dh, key_skip, stream);
}
static int mha_impl(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* o, int ldo, int B, int Nq,
                    int Nk, int H, int Dh, const int* key_skip, void* stream) {
    FAR3D_REQUIRE(q && k && v && o, "null pointer");
    FAR3D_REQUIRE(B > 0 && Nq > 0 && Nk > 0 && H > 0, "non-positive size");
    if (Dh != 32) return fail(FAR3D_E_UNSUPPORTED, "%smha supports head dim 32 (got %ld)", "", Dh);
    FAR3D_REQUIRE(((uintptr_t)k % 16 == 0) && ((uintptr_t)v % 16 == 0) && ldk % 4 == 0 && ldv % 4 == 0,
                  "k/v must be 16-byte aligned");
    FAR3D_REQUIRE(((uintptr_t)q % 16 == 0) && ((uintptr_t)o % 16 == 0) && ldq % 4 == 0 && ldo % 4 == 0,
                  "q/o must be 16-byte aligned");
    if (!g_mha_simt && ldq % 2 == 0 && ldo % 2 == 0)
        return far3d_mha_mma_launch(q, ldq, k, ldk, v, ldv, o, ldo, B, Nq, Nk, H, key_skip, stream);
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(mha_d32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MHA_SMEM) != cudaSuccess)
            return fail(FAR3D_E_CUDA, "%smha: cannot opt in to %ld bytes of shared memory", "", (long)MHA_SMEM);
        attr_set = true;
    }
    int qtiles = cdiv(Nq, MHA_QT);
    mha_d32_kernel<<<B * H * qtiles, MHA_WARPS * 32, MHA_SMEM, (cudaStream_t)stream>>>(q, ldq, k, ldk, v, ldv, o, ldo, B, Nq, Nk, H,
                                                                                      key_skip);
    return launched("mha_d32_kernel");
}

extern "C" int far3d_pos2posemb3d(const float* pos, float* emb, int M, int F, void* stream) {
    FAR3D_REQUIRE(pos && emb && M > 0 && F > 0, "bad argument");
    pos2posemb3d_kernel<<<cdiv((long)M * 3 * F, 256), 256, 0, (cudaStream_t)stream>>>(pos, emb, M, F);
    return launched("pos2posemb3d_kernel");
}
extern "C" int far3d_pos2posemb1d(const float* pos, int ldp, float* emb, int M, int F, void* stream) {
    FAR3D_REQUIRE(pos && emb && M > 0 && F > 0 && ldp > 0, "bad argument");
    pos2posemb1d_kernel<<<cdiv((long)M * F, 256), 256, 0, (cudaStream_t)stream>>>(pos, ldp, emb, M, F);
    return launched("pos2posemb1d_kernel");
}
extern "C" int far3d_nerf_posenc(const float* x, float* emb, int M, int Cin, int nfreq, void* stream) {
    FAR3D_REQUIRE(x && emb && M > 0 && Cin > 0 && nfreq > 0, "bad argument");
    nerf_posenc_kernel<<<cdiv((long)M * Cin * 2 * nfreq, 256), 256, 0, (cudaStream_t)stream>>>(x, emb, M, Cin, nfreq);
    return launched("nerf_posenc_kernel");
}

extern "C" int far3d_mln_flatten(const float* x, const float* gamma, const float* beta, float* out, int BN, int HW,
                                 int C, int S, int start, int channels_last, void* stream) {
    FAR3D_REQUIRE(x && gamma && beta && out, "null pointer");
    FAR3D_REQUIRE(BN > 0 && HW > 0 && C > 0 && S >= start + HW && start >= 0, "bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    if (channels_last) {
        FAR3D_REQUIRE(C % 4 == 0 && (uintptr_t)x % 16 == 0 && (uintptr_t)out % 16 == 0, "C % 4 and 16-byte alignment");
        mln_flatten_nhwc_kernel<<<cdiv((long)BN * HW * (C / 4), 256), 256, 0, st>>>(x, gamma, beta, out, BN, HW, C, S, start);
        return launched("mln_flatten_nhwc_kernel");
    }
    dim3 grid(cdiv(HW, 32), cdiv(C, 32), BN), block(32, 8);
    mln_flatten_nchw_kernel<<<grid, block, 0, st>>>(x, gamma, beta, out, BN, HW, C, S, start);
    return launched("mln_flatten_nchw_kernel");
}

extern "C" int far3d_mln_tokens(const float* x, const float* gamma, const float* beta, float* out, int M, int C,
                                int use_ln, void* stream) {
    FAR3D_REQUIRE(x && gamma && beta && out && M > 0 && C > 0, "bad argument");
    if (C > 1024) return fail(FAR3D_E_UNSUPPORTED, "%smln supports C <= 1024 (got %ld)", "", C);
    mln_tokens_kernel<<<cdiv((long)M * 32, 256), 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, out, M, C, use_ln);
    return launched("mln_tokens_kernel");
}

extern "C" int far3d_cam_logits(const float* lidar2img, const float* const* layer_ptrs, int layers, int rows, int E, int H, int J,
                                float eps, float* out, void* stream) {
    FAR3D_REQUIRE(lidar2img && layer_ptrs && out, "null pointer");
    FAR3D_REQUIRE(layers > 0 && layers <= FAR3D_MAX_CAM_LAYERS, "layers must be 1..FAR3D_MAX_CAM_LAYERS");
    FAR3D_REQUIRE(rows > 0 && E > 0 && H > 0 && J > 0 && (size_t)(E + H) * sizeof(float) <= 48 * 1024, "bad sizes");
    CamParams p;
    p.layers = layers; p.rows = rows; p.E = E; p.H = H; p.J = J; p.eps = eps;
    for (int i = 0; i < layers; ++i) {
        const float* const* q = layer_ptrs + 7 * i;                  // host array: w0, b0, w1, b1, ln weight, ln bias, weights_fc.weight
        for (int k = 0; k < 7; ++k) FAR3D_REQUIRE(q[k], "null layer pointer");
        p.l[i].w0 = q[0]; p.l[i].b0 = q[1]; p.l[i].w1 = q[2]; p.l[i].b1 = q[3]; p.l[i].g = q[4]; p.l[i].beta = q[5]; p.l[i].wfc = q[6];
    }
    cam_logits_kernel<<<dim3(rows, layers), 256, (size_t)(E + H) * sizeof(float), (cudaStream_t)stream>>>(p, lidar2img, out);
    return launched("cam_logits_kernel");
}
