// Perspective-aware deformable aggregation for sm_100a.
//
// Replaces (one kernel instead of ~12 + a third-party one):
//   models/utils/detr3d_transformer.py:547-552  projection of key points into every camera
//   models/utils/detr3d_transformer.py:555      21 MB repeat of sampling locations over groups x levels
//   models/utils/detr3d_transformer.py:561-563  mmcv MultiScaleDeformableAttnFunction (ms_deformable_im2col)
//   models/utils/detr3d_transformer.py:565-569  sum over cameras
//
// Design (HBM / L2-gather bound, no tensor cores); details at deform_agg_kernel below:
//   * work item = (batch, query, half of the channel groups): a 4-warp CTA, warp w owns a channel group (32 channels = one
//     128 B row per pixel in fp32), so every corner fetch is a fully used line.
//   * phase A: one thread per (camera, point) pair projects ONCE and tests the bounds of every level; a block scan of the
//     per-pair counts gives every in-view sample a deterministic position (no atomics).  Typically 1-2 of 7 cameras see a
//     point, so the gather loop runs over ~15-25 % of the cam x level x point grid.
//   * records {row index, bilinear weight} x 4 corners are written per chunk of 128 positions (8 KB of shared memory per
//     CTA: the rest of the 256 KB array stays L1 for the gather); an out-of-map corner aliases an in-map one with weight 0.
//   * phase B: a warp instruction fetches the four corner rows of TWO samples (256-bit load per lane), 4 in flight per lane;
//     three shuffle levels fold corners and sample halves; the output row (128 B per group) is written once.  No per-camera
//     outputs, no materialised sampling locations.
//
// Index/mask arithmetic is fixed (fused multiply-adds spelled out) so the CPU oracle
// (oracle/deform_agg_ref.c) reproduces floor indices and in-bounds masks bit-exactly.
#include <atomic>
#include "common.cuh"

namespace far3d {

struct LevelInfo {
    int H[FAR3D_MAX_LEVELS];
    int W[FAR3D_MAX_LEVELS];
    int start[FAR3D_MAX_LEVELS];
};

// p = M @ [x,y,z,1]; u = p0/max(p2,1e-5)/pad_w; v = p1/max(p2,1e-5)/pad_h   (detr3d_transformer.py:547-552)
__device__ __forceinline__ void project_point(const float* __restrict__ m, float x, float y, float z, float pad_h,
                                              float pad_w, float& u, float& v) {
    float p[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float acc = __fmul_rn(m[4 * i + 0], x);
        acc = __fmaf_rn(m[4 * i + 1], y, acc);
        acc = __fmaf_rn(m[4 * i + 2], z, acc);
        p[i] = __fadd_rn(acc, m[4 * i + 3]);
    }
    float zc = (p[2] != p[2]) ? p[2] : fmaxf(p[2], 1e-5f);
    u = __fdiv_rn(__fdiv_rn(p[0], zc), pad_w);
    v = __fdiv_rn(__fdiv_rn(p[1], zc), pad_h);
}

// mmcv im2col coordinates + bounds rule. Returns validity.
__device__ __forceinline__ bool sample_coords(float u, float v, int H, int W, float& h_im, float& w_im) {
    w_im = __fmaf_rn(u, (float)W, -0.5f);
    h_im = __fmaf_rn(v, (float)H, -0.5f);
    return (h_im > -1.f) && (w_im > -1.f) && (h_im < (float)H) && (w_im < (float)W);
}

struct __align__(16) SampleRec {
    int off[4];    // pixel index (level start + y*W + x) of the 4 corners, -1 if outside the map
    float cw[4];   // bilinear corner weights hh*hw, hh*lw, lh*hw, lh*lw
};

__device__ __forceinline__ void make_rec(float h_im, float w_im, int H, int W, int start, SampleRec& r) {
    float hf = floorf(h_im), wf = floorf(w_im);
    int h_low = (int)hf, w_low = (int)wf;
    float lh = h_im - hf, lw = w_im - wf;
    float hh = 1.f - lh, hw = 1.f - lw;
    bool y0 = h_low >= 0, y1 = h_low + 1 <= H - 1, x0 = w_low >= 0, x1 = w_low + 1 <= W - 1;
    int base = start + h_low * W + w_low;
    r.off[0] = (y0 && x0) ? base : -1;
    r.off[1] = (y0 && x1) ? base + 1 : -1;
    r.off[2] = (y1 && x0) ? base + W : -1;
    r.off[3] = (y1 && x1) ? base + W + 1 : -1;
    r.cw[0] = hh * hw; r.cw[1] = hh * lw; r.cw[2] = lh * hw; r.cw[3] = lh * lw;
}

template <typename T> struct Quad;
template <> struct Quad<float> {
    static __device__ __forceinline__ float4 load(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
};
template <> struct Quad<__nv_bfloat16> {
    static __device__ __forceinline__ float4 load(const __nv_bfloat16* p) {
        uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
        __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x), b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
        float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
        return make_float4(fa.x, fa.y, fb.x, fb.y);
    }
};

constexpr int DA_THREADS = 256;
constexpr int DA_MAX_CAMS = 16;

template <> struct Quad<__half> {
    static __device__ __forceinline__ float4 load(const __half* p) {
        uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
        float2 fa = __half22float2(*reinterpret_cast<__half2*>(&r.x)), fb = __half22float2(*reinterpret_cast<__half2*>(&r.y));
        return make_float4(fa.x, fa.y, fb.x, fb.y);
    }
};

// 8 consecutive channels of one pixel: one 256-bit load (LDG.E.256, sm_100) for fp32 features, one 128-bit load for 16-bit ones
template <typename T> struct Oct;
template <> struct Oct<float> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
        asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                     : "l"(p));
    }
};
template <> struct Oct<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
        const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    }
};
template <> struct Oct<__half> {
    static __device__ __forceinline__ void load(const __half* p, float (&v)[8]) {
        const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    }
};

// One corner of one in-view sample as phase B consumes it: `idx` = pixel row (cam*S + level start + y*W + x) in units of
// 4 channels (x C/4), so a lane's address is base + idx with one IMAD.WIDE; `cw` = bilinear corner weight.  A corner that
// falls outside the map (mmcv skips it) points at a valid corner of the same sample with cw = 0: no predicate, no zeroing,
// identical sum for finite features.
struct __align__(8) Corner { uint32_t idx; float cw; };

// (u, v) of pair t = n*P + p and the bit mask of levels whose bounds test passes
__device__ __forceinline__ unsigned pair_levels(const float* __restrict__ lidar2img_b, const float* __restrict__ kp_q, int t,
                                                int P, const LevelInfo& lv, int L, float pad_h, float pad_w, float& u,
                                                float& v) {
    const int n = t / P, p = t - n * P;
    float m[12];
    const float4* mp = reinterpret_cast<const float4*>(lidar2img_b + (size_t)n * 16);
#pragma unroll
    for (int i = 0; i < 3; ++i) { float4 r = __ldg(mp + i); m[4 * i] = r.x; m[4 * i + 1] = r.y; m[4 * i + 2] = r.z; m[4 * i + 3] = r.w; }
    project_point(m, __ldg(kp_q + p * 3), __ldg(kp_q + p * 3 + 1), __ldg(kp_q + p * 3 + 2), pad_h, pad_w, u, v);
    unsigned mask = 0;
#pragma unroll
    for (int l = 0; l < FAR3D_MAX_LEVELS; ++l) {
        float h_im, w_im;
        if (l < L && sample_coords(u, v, lv.H[l], lv.W[l], h_im, w_im)) mask |= 1u << l;
    }
    return mask;
}

// Samples a CTA holds in shared memory at a time.  The records of ALL N*L*P entries would be 19 KB per CTA (150 KB per SM at
// 8 CTAs), taken from the same 256 KB array that is the L1 the gather lives in (62 % L1 hit rate at cfg-2); a query sees 68
// samples on average (19 % of 364), so the kernel keeps DA_CHUNK of them and loops in the rare case there are more.
constexpr int DA_CHUNK = 128;

// -DFAR3D_DA_PHASES (tools/agg_phases.py): per-CTA cycle counts of the kernel's phases, summed over the CTA's work items:
// [0] phase A (projection + scan), [1] records, [2] softmax-weight gather, [3] feature gather, [4] fold + store, [5] items,
// [6] in-view samples, [7] queue pull
#ifdef FAR3D_DA_PHASES
__device__ long long* g_da_phases = nullptr;
#define DA_T(i) do { const long long _n = clock64(); if (tid == 0) ph[i] += _n - tph; tph = _n; } while (0)
#else
#define DA_T(i) do { } while (0)
#endif

struct __align__(16) PairRec { float u, v; uint32_t mask; int pos0; };

// Work items = B*Nq*parts, block = WARPS*32.  `sched` == nullptr: one CTA per item (grid = items).  Otherwise the grid is one
// resident wave (SMs x CTAs per SM) and every CTA pulls items from sched[0] until they run out: the work per query varies
// with the number of in-view samples (0 .. N*L*P) and 1800-2100 static CTAs are 1.5-1.8 waves, so the SMs sat idle for a third
// of the launch (profiles/r1i_deform_agg_ncu_summary.txt: SMs active 65 %).  sched[1] counts finished CTAs; the last one
// re-arms both counters for the next launch.
// dynamic smem: PairRec[N*P] | Corner[4*DA_CHUNK] | int widx[DA_CHUNK] | float wc[WARPS][DA_CHUNK]
//
// phase A (one (camera, point) pair per thread): project ONCE (the pair's (u,v) serves all levels), test the per-level
//   bounds, block-scan the counts (warp shuffles + one barrier): every in-view sample gets a position in (camera, point,
//   level) order - deterministic, no atomics.
// per chunk of DA_CHUNK positions: the owning threads write the samples' corner records, then
// phase B (warp = channel group): WIDE: lane = half*16 + corner*4 + oct, one 256-bit load per lane, a warp instruction
//   fetches the 4 corner rows of TWO samples; narrow: lane = corner*8 + quad, 128-bit loads, one sample per instruction.
//   Per sample and lane: one LDS.64 (corner record), one broadcast LDS (the group's softmax weight, gathered once per warp
//   into smem), one IMAD.WIDE, one load, one FMUL, the FFMAs.  The first version of this kernel spent 25 instructions per
//   sample and was issue-bound (profiles/r1d_deform_agg_ncu_summary.txt); this one is bound by the L1/L2 latency of the gather.
// PRE: the records {corner row, bilinear weight} x 4, the in-view count and the softmax weights COMPACTED to the in-view
// samples come from far3d_dfa_prepare (the softmax kernel projects its query's key points while it is at it): phase A and the
// record building - 2.5 of the 12.9 us a work item takes, repeated by both CTAs of a query - become one coalesced copy, and the
// softmax-weight gather through s_widx (1.7 us) a contiguous read (profiles/r2w_agg_phases.txt).
// Tried and dropped (profiles/r2y_agg_variants.txt): cp.async.bulk.prefetch.L2 of the corner rows when the records are built
// (slower: 46.5 vs 43.0 us), fetching the softmax weight inside the gather loop (no change).
template <typename FeatT, int U, int WARPS, bool WIDE, int NG, bool PRE>
__global__ void __launch_bounds__(WARPS * 32, (U > 4 && WIDE ? 16 : 32) / WARPS)
deform_agg_kernel(const FeatT* __restrict__ feat, LevelInfo lv, const float* __restrict__ key_points,
                  const float* __restrict__ lidar2img, const float* __restrict__ weights, float pad_h, float pad_w,
                  float* __restrict__ out, int B, int N, int S, int C, int G, int Nq, int L, int P, int parts, int* __restrict__ sched,
                  const int* __restrict__ pre_cnt, const Corner* __restrict__ pre_rec, const float* __restrict__ pre_w) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_item;
    const int items = B * Nq * parts;
    const int NP = N * P, LP = L * P;
    PairRec* s_pair = reinterpret_cast<PairRec*>(smem_raw);                                // [NP]
    Corner* s_rec = reinterpret_cast<Corner*>(smem_raw + (size_t)NP * sizeof(PairRec));    // [DA_CHUNK][4]
    int* s_widx = reinterpret_cast<int*>(s_rec + 4 * DA_CHUNK);                            // [DA_CHUNK]  n*Nq*G*LP + l*P + p
    float* s_wc = reinterpret_cast<float*>(s_widx + DA_CHUNK);                             // [WARPS][DA_CHUNK]
    __shared__ int s_wtot[WARPS];
    constexpr int THREADS = WARPS * 32;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef FAR3D_DA_PHASES
    long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tph = clock64();
#endif
  for (;;) {
    int item = blockIdx.x;
    if (sched) {
        if (tid == 0) s_item = atomicAdd(&sched[0], 1);
        __syncthreads();                               // also: every warp is done with the previous item's shared memory
        item = s_item;
        if (item >= items) break;
    }
    DA_T(7);
    // with WARPS < 8 a query's channel groups are split over `parts` items (each repeats the cheap phase A): finer work items
    const int bq = item / parts, part = item - bq * parts;
    const int b = bq / Nq, q = bq - b * Nq;
    const int wstride_n = Nq * G * LP;                 // weights: [(b*N+n), q, g, lp]
    const int C4 = C >> 2;
    const float* l2i_b = lidar2img + (size_t)b * N * 16;
    const float* kp_q = key_points + ((size_t)b * Nq + q) * P * 3;

    // ---- phase A
    int run = 0;
    if constexpr (PRE) run = __ldg(pre_cnt + bq);
    else for (int t0 = 0; t0 < NP; t0 += THREADS) {    // one round when N*P <= THREADS (cfg-2: 91 pairs)
        const int t = t0 + tid;
        float u = 0.f, v = 0.f;
        unsigned mask = 0;
        if (t < NP) mask = pair_levels(l2i_b, kp_q, t, P, lv, L, pad_h, pad_w, u, v);
        const int cnt = __popc(mask);
        int incl = cnt;                                // inclusive warp scan of the per-pair counts
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        if (t0 > 0) __syncthreads();                   // everyone has read the previous round's warp totals
        if (lane == 31) s_wtot[warp] = incl;
        __syncthreads();
        int pos = run + incl - cnt;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            const int c = s_wtot[w];
            if (w < warp) pos += c;
            run += c;
        }
        if (t < NP) {
            PairRec pr; pr.u = u; pr.v = v; pr.mask = mask; pr.pos0 = pos;
            s_pair[t] = pr;
        }
    }
    const int total = run;                             // (the barrier at the top of the chunk loop publishes s_pair)
    DA_T(0);
#ifdef FAR3D_DA_PHASES
    if (tid == 0) { ph[5] += 1; ph[6] += total; }
#endif

    const FeatT* fb = feat + (size_t)b * N * S * C;
    float* wc = s_wc + (size_t)warp * DA_CHUNK;
    const int g0 = part * WARPS + warp, gstep = WARPS * parts;
    // accumulators of the warp's NG groups live across chunks (G <= NG * WARPS * parts, chosen by the host)
    float acc[NG][8];
#pragma unroll
    for (int k = 0; k < NG; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[k][i] = 0.f;

    for (int c0 = 0; c0 < total || c0 == 0; c0 += DA_CHUNK) {
        __syncthreads();                               // s_pair written / previous chunk's records consumed
        const int nc = min(DA_CHUNK, total - c0);
        // ---- records of positions [c0, c0 + nc)
        if constexpr (PRE) {
            const uint4* src = reinterpret_cast<const uint4*>(pre_rec + ((size_t)bq * (N * LP) + c0) * 4);
            uint4* dst = reinterpret_cast<uint4*>(s_rec);
            for (int i = tid; i < 2 * nc; i += THREADS) dst[i] = __ldg(src + i);
        } else
        for (int t = tid; t < NP; t += THREADS) {
            const PairRec pr = s_pair[t];
            if (!pr.mask || pr.pos0 >= c0 + nc || pr.pos0 + (int)__popc(pr.mask) <= c0) continue;
            const int n = t / P, p = t - n * P;
            int pos = pr.pos0 - c0;
#pragma unroll
            for (int l = 0; l < FAR3D_MAX_LEVELS; ++l) {
                if (!(pr.mask & (1u << l))) continue;
                if (pos >= 0 && pos < nc) {
                    float h_im, w_im;
                    sample_coords(pr.u, pr.v, lv.H[l], lv.W[l], h_im, w_im);
                    SampleRec rec;
                    make_rec(h_im, w_im, lv.H[l], lv.W[l], lv.start[l], rec);
                    // >= 1 corner is inside whenever the sample passes the bounds test (-1 < h_im < H  =>  -1 <= floor <= H-1)
                    const int any = rec.off[0] >= 0 ? rec.off[0] : rec.off[1] >= 0 ? rec.off[1] : rec.off[2] >= 0 ? rec.off[2] : rec.off[3];
                    uint32_t ci[4]; float cw[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const bool in = rec.off[c] >= 0;
                        ci[c] = (uint32_t)((in ? rec.off[c] : any) + n * S) * (uint32_t)C4;
                        cw[c] = in ? rec.cw[c] : 0.f;
                    }
                    uint4* dst = reinterpret_cast<uint4*>(s_rec + (size_t)pos * 4);
                    dst[0] = make_uint4(ci[0], __float_as_uint(cw[0]), ci[1], __float_as_uint(cw[1]));
                    dst[1] = make_uint4(ci[2], __float_as_uint(cw[2]), ci[3], __float_as_uint(cw[3]));
                    s_widx[pos] = n * wstride_n + l * P + p;
                }
                ++pos;
            }
        }
        __syncthreads();
        DA_T(1);

        // ---- phase B on this chunk
#pragma unroll
        for (int k = 0; k < NG; ++k) {
            const int g = g0 + k * gstep;
            if (g >= G) break;
            const float* wrow = weights + (((size_t)b * N) * Nq + q) * G * LP + (size_t)g * LP;
            __syncwarp();
            if constexpr (PRE) {
                const float* wsrc = pre_w + ((size_t)bq * G + g) * (N * LP) + c0;
                for (int j = lane; j < nc; j += 32) wc[j] = __ldg(wsrc + j);
            } else {
                for (int j = lane; j < nc; j += 32) wc[j] = __ldg(wrow + s_widx[j]);
            }
            __syncwarp();
            DA_T(2);
            if constexpr (WIDE) {
                const int half = lane >> 4, corner = (lane >> 2) & 3, oct = lane & 3;
                const FeatT* fg = fb + g * 32 + oct * 8;
                const Corner* rc = s_rec + corner;
                int j = 0;
                for (; j + 2 * U <= nc; j += 2 * U) {
                    Corner r[U]; float w[U]; float x[U][8];
#pragma unroll
                    for (int t = 0; t < U; ++t) {
                        const int sidx = j + 2 * t + half;
                        r[t] = rc[sidx * 4];
                        w[t] = wc[sidx];
                    }
#pragma unroll
                    for (int t = 0; t < U; ++t) Oct<FeatT>::load(fg + (size_t)r[t].idx * 4, x[t]);
#pragma unroll
                    for (int t = 0; t < U; ++t) {
                        const float cw = r[t].cw * w[t];
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[k][i] = fmaf(cw, x[t][i], acc[k][i]);
                    }
                }
                for (; j < nc; j += 2) {
                    const int sidx = j + half;
                    if (sidx < nc) {
                        const Corner r = rc[sidx * 4];
                        const float cw = r.cw * wc[sidx];
                        float x[8];
                        Oct<FeatT>::load(fg + (size_t)r.idx * 4, x);
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[k][i] = fmaf(cw, x[i], acc[k][i]);
                    }
                }
                __syncwarp();
            } else {
                const int corner = lane >> 3, quad = lane & 7;
                const FeatT* fg = fb + g * 32 + quad * 4;
                const Corner* rc = s_rec + corner;
                int j = 0;
                for (; j + U <= nc; j += U) {
                    Corner r[U]; float w[U]; float4 x[U];
#pragma unroll
                    for (int t = 0; t < U; ++t) { r[t] = rc[(j + t) * 4]; w[t] = wc[j + t]; }
#pragma unroll
                    for (int t = 0; t < U; ++t) x[t] = Quad<FeatT>::load(fg + (size_t)r[t].idx * 4);
#pragma unroll
                    for (int t = 0; t < U; ++t) {
                        const float cw = r[t].cw * w[t];
                        acc[k][0] = fmaf(cw, x[t].x, acc[k][0]); acc[k][1] = fmaf(cw, x[t].y, acc[k][1]);
                        acc[k][2] = fmaf(cw, x[t].z, acc[k][2]); acc[k][3] = fmaf(cw, x[t].w, acc[k][3]);
                    }
                }
                for (; j < nc; ++j) {
                    const Corner r = rc[j * 4];
                    const float cw = r.cw * wc[j];
                    const float4 x = Quad<FeatT>::load(fg + (size_t)r.idx * 4);
                    acc[k][0] = fmaf(cw, x.x, acc[k][0]); acc[k][1] = fmaf(cw, x.y, acc[k][1]);
                    acc[k][2] = fmaf(cw, x.z, acc[k][2]); acc[k][3] = fmaf(cw, x.w, acc[k][3]);
                }
            }
            DA_T(3);
        }
    }

    // ---- fold the corner (and, WIDE, the two-sample) partial sums across the warp; one 128-byte row per group
#pragma unroll
    for (int k = 0; k < NG; ++k) {
        const int g = g0 + k * gstep;
        if (g >= G) break;
        float* orow = out + ((size_t)b * Nq + q) * C + g * 32;
        if constexpr (WIDE) {
#pragma unroll
            for (int o = 4; o <= 16; o <<= 1)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[k][i] += __shfl_xor_sync(0xffffffffu, acc[k][i], o);
            if (lane < 4) {
                *reinterpret_cast<float4*>(orow + lane * 8) = make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
                *reinterpret_cast<float4*>(orow + lane * 8 + 4) = make_float4(acc[k][4], acc[k][5], acc[k][6], acc[k][7]);
            }
        } else {
#pragma unroll
            for (int o = 8; o <= 16; o <<= 1)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[k][i] += __shfl_xor_sync(0xffffffffu, acc[k][i], o);
            if (lane < 8) *reinterpret_cast<float4*>(orow + lane * 4) = make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
        }
    }
    DA_T(4);
    if (!sched) break;
  }
#ifdef FAR3D_DA_PHASES
    if (tid == 0 && g_da_phases)
        for (int i = 0; i < 8; ++i) g_da_phases[(size_t)blockIdx.x * 8 + i] = ph[i];
#endif
    if (sched && tid == 0) {
        __threadfence();
        if (atomicAdd(&sched[1], 1) == (int)gridDim.x - 1) {      // every other CTA has made its last pull
            sched[0] = 0; sched[1] = 0;
            __threadfence();
        }
    }
}

// Generic group width (any D): one thread per (b, q, channel); plain loops, same arithmetic.
template <typename FeatT>
__global__ void deform_agg_generic_kernel(const FeatT* __restrict__ feat, LevelInfo lv,
                                          const float* __restrict__ key_points, const float* __restrict__ lidar2img,
                                          const float* __restrict__ weights, float pad_h, float pad_w,
                                          float* __restrict__ out, int B, int N, int S, int C, int G, int Nq, int L,
                                          int P) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)B * Nq * C) return;
    int c = (int)(idx % C);
    long bq = idx / C;
    int q = (int)(bq % Nq), b = (int)(bq / Nq);
    int D = C / G, g = c / D, LP = L * P;
    float acc = 0.f;
    for (int n = 0; n < N; ++n) {
        const float* m = lidar2img + ((size_t)b * N + n) * 16;
        const FeatT* fcam = feat + ((size_t)b * N + n) * S * C + c;
        const float* wrow = weights + ((((size_t)b * N + n) * Nq + q) * G + g) * LP;
        for (int p = 0; p < P; ++p) {
            const float* kp = key_points + (((size_t)b * Nq + q) * P + p) * 3;
            float u, v;
            project_point(m, kp[0], kp[1], kp[2], pad_h, pad_w, u, v);
            for (int l = 0; l < L; ++l) {
                float h_im, w_im;
                if (!sample_coords(u, v, lv.H[l], lv.W[l], h_im, w_im)) continue;
                SampleRec r;
                make_rec(h_im, w_im, lv.H[l], lv.W[l], lv.start[l], r);
                float wgt = wrow[l * P + p];
                float val = 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (r.off[k] >= 0) val = fmaf(r.cw[k], (float)fcam[(size_t)r.off[k] * C], val);
                acc = fmaf(wgt, val, acc);
            }
        }
    }
    out[idx] = acc;
}

__global__ void deform_agg_debug_kernel(LevelInfo lv, const float* __restrict__ key_points,
                                        const float* __restrict__ lidar2img, float pad_h, float pad_w,
                                        float* __restrict__ uv, int32_t* __restrict__ idx, uint8_t* __restrict__ valid,
                                        int B, int N, int Nq, int L, int P) {
    long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long total = (long)B * N * Nq * P;
    if (t >= total) return;
    int p = (int)(t % P); long r = t / P;
    int q = (int)(r % Nq); r /= Nq;
    int n = (int)(r % N); int b = (int)(r / N);
    const float* kp = key_points + (((size_t)b * Nq + q) * P + p) * 3;
    float u, v;
    project_point(lidar2img + ((size_t)b * N + n) * 16, kp[0], kp[1], kp[2], pad_h, pad_w, u, v);
    if (uv) { uv[2 * t] = u; uv[2 * t + 1] = v; }
    for (int l = 0; l < L; ++l) {
        float h_im, w_im;
        bool ok = sample_coords(u, v, lv.H[l], lv.W[l], h_im, w_im);
        size_t si = ((((size_t)b * N + n) * Nq + q) * L + l) * P + p;
        float hf = floorf(h_im), wf = floorf(w_im);
        int hl = (hf >= -2147483000.f && hf <= 2147483000.f) ? (int)hf : 0;
        int wl = (wf >= -2147483000.f && wf <= 2147483000.f) ? (int)wf : 0;
        if (idx) { idx[2 * si] = hl; idx[2 * si + 1] = wl; }
        if (valid) valid[si] = ok ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------ MSDA (mmcv layout)
// warp per (bn, q, g) for D == 32: lane = corner*8 + quad.
__global__ void __launch_bounds__(256)
msda_d32_kernel(const float* __restrict__ value, const int64_t* __restrict__ shapes, const int64_t* __restrict__ start,
                const float* __restrict__ loc, const float* __restrict__ attw, float* __restrict__ out, int BN, int S,
                int G, int Nq, int L, int P) {
    long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wid >= (long)BN * Nq * G) return;
    int lane = threadIdx.x & 31, corner = lane >> 3, quad = lane & 7;
    int g = (int)(wid % G); long r = wid / G;
    int q = (int)(r % Nq); int bn = (int)(r / Nq);
    const int C = G * 32;
    const float* vb = value + (size_t)bn * S * C + g * 32 + quad * 4;
    const float* lrow = loc + (size_t)wid * L * P * 2;
    const float* wrow = attw + (size_t)wid * L * P;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int l = 0; l < L; ++l) {
        int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1], st = (int)start[l];
        for (int p = 0; p < P; ++p) {
            float u = __ldg(lrow + (l * P + p) * 2), v = __ldg(lrow + (l * P + p) * 2 + 1);
            float h_im, w_im;
            if (!sample_coords(u, v, H, W, h_im, w_im)) continue;     // warp-uniform
            SampleRec rec;
            make_rec(h_im, w_im, H, W, st, rec);
            int off = rec.off[corner];
            float cw = rec.cw[corner] * __ldg(wrow + l * P + p);
            if (off >= 0) {
                float4 x = ldg_f4(vb + (size_t)off * C);
                acc.x = fmaf(cw, x.x, acc.x); acc.y = fmaf(cw, x.y, acc.y);
                acc.z = fmaf(cw, x.z, acc.z); acc.w = fmaf(cw, x.w, acc.w);
            }
        }
    }
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
        acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
    }
    if (corner == 0) *reinterpret_cast<float4*>(out + ((size_t)bn * Nq + q) * C + g * 32 + quad * 4) = acc;
}

__global__ void msda_generic_kernel(const float* __restrict__ value, const int64_t* __restrict__ shapes,
                                    const int64_t* __restrict__ start, const float* __restrict__ loc,
                                    const float* __restrict__ attw, float* __restrict__ out, int BN, int S, int G, int D,
                                    int Nq, int L, int P) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)BN * Nq * G * D) return;
    int d = (int)(idx % D); long r = idx / D;
    int g = (int)(r % G); r /= G;
    int q = (int)(r % Nq); int bn = (int)(r / Nq);
    const int C = G * D;
    const float* vb = value + (size_t)bn * S * C + g * D + d;
    size_t w0 = (((size_t)bn * Nq + q) * G + g) * L * P;
    float acc = 0.f;
    for (int l = 0; l < L; ++l) {
        int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1], st = (int)start[l];
        for (int p = 0; p < P; ++p) {
            float u = loc[(w0 + l * P + p) * 2], v = loc[(w0 + l * P + p) * 2 + 1];
            float h_im, w_im;
            if (!sample_coords(u, v, H, W, h_im, w_im)) continue;
            SampleRec rec;
            make_rec(h_im, w_im, H, W, st, rec);
            float val = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (rec.off[k] >= 0) val = fmaf(rec.cw[k], vb[(size_t)rec.off[k] * C], val);
            acc = fmaf(attw[w0 + l * P + p], val, acc);
        }
    }
    out[idx] = acc;
}

// ------------------------------------------------------------------------------------------ weights softmax
// logits[b,q,n,lp,g] = wq[b,q,lp*G+g] + wc[b,n,lp*G+g]; softmax over (n,lp) per (b,q,g); out [b*N+n, q, g, lp].
// One warp per (b,q,g).
__global__ void __launch_bounds__(256)
dfa_weights_softmax_kernel(const float* __restrict__ wq, const float* __restrict__ wc, float* __restrict__ weights,
                           int B, int N, int Nq, int G, int LP) {
    long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wid >= (long)B * Nq * G) return;
    int lane = threadIdx.x & 31;
    int g = (int)(wid % G); long r = wid / G;
    int q = (int)(r % Nq); int b = (int)(r / Nq);
    const float* a = wq + ((size_t)b * Nq + q) * LP * G + g;
    const float* c = wc + (size_t)b * N * LP * G + g;
    const int E = N * LP;
    float mx = -INFINITY;
    for (int e = lane; e < E; e += 32) {
        int n = e / LP, lp = e - n * LP;
        mx = fmaxf(mx, a[(size_t)lp * G] + c[((size_t)n * LP + lp) * G]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int e = lane; e < E; e += 32) {
        int n = e / LP, lp = e - n * LP;
        sum += expf(a[(size_t)lp * G] + c[((size_t)n * LP + lp) * G] - mx);
    }
    sum = warp_sum(sum);
    float inv = 1.f / sum;
    for (int e = lane; e < E; e += 32) {
        int n = e / LP, lp = e - n * LP;
        float v = expf(a[(size_t)lp * G] + c[((size_t)n * LP + lp) * G] - mx) * inv;
        weights[((((size_t)b * N + n) * Nq + q) * G + g) * LP + lp] = v;
    }
}

// Same op, one CTA per (b, q): the query's logit row and the camera rows are staged in smem transposed to [g][.] (the
// per-warp version read both with a stride of G floats, one 32-byte sector per lane), each warp owns a group, keeps its
// <= 512 logits in registers and evaluates exp once.
constexpr int DWS_MAXK = 16;
__global__ void __launch_bounds__(256)
dfa_weights_softmax_q_kernel(const float* __restrict__ wq, const float* __restrict__ wc, float* __restrict__ weights,
                             int B, int N, int Nq, int G, int LP) {
    extern __shared__ float dws_smem[];                   // sa[G][LP] | sc[G][N*LP]
    const int E = N * LP;
    float* sa = dws_smem;
    float* sc = dws_smem + G * LP;
    const int q = blockIdx.x % Nq, b = blockIdx.x / Nq;
    const float* a = wq + ((size_t)b * Nq + q) * LP * G;
    const float* c = wc + (size_t)b * N * LP * G;
    for (int i = threadIdx.x; i < LP * G; i += blockDim.x) sa[(i % G) * LP + i / G] = a[i];
    for (int i = threadIdx.x; i < E * G; i += blockDim.x) sc[(i % G) * E + i / G] = c[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int g = warp; g < G; g += 8) {
        float v[DWS_MAXK];
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < DWS_MAXK; ++k) {
            const int e = lane + 32 * k;
            v[k] = -INFINITY;
            if (e < E) { v[k] = sa[g * LP + e % LP] + sc[g * E + e]; mx = fmaxf(mx, v[k]); }
        }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < DWS_MAXK; ++k) {
            const int e = lane + 32 * k;
            if (e < E) { v[k] = expf(v[k] - mx); sum += v[k]; }
        }
        sum = warp_sum(sum);
        const float inv = 1.f / sum;
#pragma unroll
        for (int k = 0; k < DWS_MAXK; ++k) {
            const int e = lane + 32 * k;
            if (e < E) {
                const int n = e / LP, lp = e - n * LP;
                weights[((((size_t)b * N + n) * Nq + q) * G + g) * LP + lp] = v[k] * inv;
            }
        }
    }
}

// far3d_dfa_prepare: the softmax above + the projection of the query's key points, one CTA per (b, q).  Emits what the PRE form
// of deform_agg_kernel consumes: cnt[bq] in-view samples, their corner records in (camera, point, level) order - the order
// and the arithmetic of deform_agg_kernel's phase A, so the masks / indices stay bit-exact - and the softmax weights of
// exactly those samples, [bq][g][position].  The full [B*N, Nq, G, LP] weight tensor is written only when `weights` is given.
__global__ void __launch_bounds__(256)
dfa_prepare_kernel(const float* __restrict__ wq, const float* __restrict__ wc, const float* __restrict__ key_points,
                   const float* __restrict__ lidar2img, LevelInfo lv, float pad_h, float pad_w, float* __restrict__ weights,
                   int* __restrict__ cnt_out, Corner* __restrict__ rec_out, float* __restrict__ w_out, int B, int N, int Nq, int G,
                   int L, int P, int S, int C) {
    extern __shared__ float dws_smem[];                   // pos[N*LP]
    const int LP = L * P, E = N * LP, NP = N * P;
    int* s_pos = reinterpret_cast<int*>(dws_smem);
    __shared__ int s_wtot[8];
    const int bq = blockIdx.x, q = bq % Nq, b = bq / Nq;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nthreads = blockDim.x, nwarps = nthreads >> 5;             // 128 or 256 threads (far3d_dfa_prepare picks)
    const float* a = wq + (size_t)bq * LP * G;
    const float* c = wc + (size_t)b * N * LP * G;
    for (int i = tid; i < E; i += nthreads) s_pos[i] = -1;
    __syncthreads();
    // ---- projection + positions (deform_agg_kernel phase A) + records
    const float* l2i_b = lidar2img + (size_t)b * N * 16;
    const float* kp_q = key_points + (size_t)bq * P * 3;
    const int C4 = C >> 2;
    int run = 0;
    for (int t0 = 0; t0 < NP; t0 += nthreads) {
        const int t = t0 + tid;
        float u = 0.f, v = 0.f;
        unsigned mask = 0;
        if (t < NP) mask = pair_levels(l2i_b, kp_q, t, P, lv, L, pad_h, pad_w, u, v);
        const int n_in = __popc(mask);
        int incl = n_in;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        if (t0 > 0) __syncthreads();
        if (lane == 31) s_wtot[warp] = incl;
        __syncthreads();
        int pos = run + incl - n_in;
        for (int w = 0; w < nwarps; ++w) {
            const int cw_ = s_wtot[w];
            if (w < warp) pos += cw_;
            run += cw_;
        }
        if (mask) {
            const int n = t / P, p = t - n * P;
#pragma unroll
            for (int l = 0; l < FAR3D_MAX_LEVELS; ++l) {
                if (!(mask & (1u << l))) continue;
                float h_im, w_im;
                sample_coords(u, v, lv.H[l], lv.W[l], h_im, w_im);
                SampleRec rec;
                make_rec(h_im, w_im, lv.H[l], lv.W[l], lv.start[l], rec);
                const int any = rec.off[0] >= 0 ? rec.off[0] : rec.off[1] >= 0 ? rec.off[1] : rec.off[2] >= 0 ? rec.off[2] : rec.off[3];
                uint32_t ci[4]; float cw[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const bool in = rec.off[k] >= 0;
                    ci[k] = (uint32_t)((in ? rec.off[k] : any) + n * S) * (uint32_t)C4;
                    cw[k] = in ? rec.cw[k] : 0.f;
                }
                uint4* dst = reinterpret_cast<uint4*>(rec_out + ((size_t)bq * E + pos) * 4);
                dst[0] = make_uint4(ci[0], __float_as_uint(cw[0]), ci[1], __float_as_uint(cw[1]));
                dst[1] = make_uint4(ci[2], __float_as_uint(cw[2]), ci[3], __float_as_uint(cw[3]));
                s_pos[n * LP + l * P + p] = pos;
                ++pos;
            }
        }
    }
    if (tid == 0) cnt_out[bq] = run;
    __syncthreads();
    // ---- softmax over cameras x levels x points per group (dfa_weights_softmax_q_kernel); in-view entries go out compacted.
    // A warp owns a group and reads its logits straight from the [.., lp, g] rows (stride G floats: the 8 warps of the CTA read the
    // same sectors for their 8 groups, so all but the first hit L1); entry e = lane + 32 k is (camera n, lp) with lp / n advanced
    // incrementally - the smem transpose and the per-element divisions of the first version were a third of its instructions.
    for (int g = warp; g < G; g += nwarps) {
        float v[DWS_MAXK];
        float mx = -INFINITY;
        int lp = lane % LP;
#pragma unroll
        for (int k = 0; k < DWS_MAXK; ++k) {
            const int e = lane + 32 * k;
            v[k] = -INFINITY;
            if (e < E) { v[k] = __ldg(a + lp * G + g) + __ldg(c + (size_t)e * G + g); mx = fmaxf(mx, v[k]); }
            lp += 32;
            while (lp >= LP) lp -= LP;
        }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < DWS_MAXK; ++k) {
            const int e = lane + 32 * k;
            if (e < E) { v[k] = expf(v[k] - mx); sum += v[k]; }
        }
        sum = warp_sum(sum);
        const float inv = 1.f / sum;
        lp = lane % LP;
        int n = lane / LP;
#pragma unroll
        for (int k = 0; k < DWS_MAXK; ++k) {
            const int e = lane + 32 * k;
            if (e < E) {
                const float wv = v[k] * inv;
                const int pos = s_pos[e];
                if (pos >= 0) w_out[((size_t)bq * G + g) * E + pos] = wv;
                if (weights) weights[((((size_t)b * N + n) * Nq + q) * G + g) * LP + lp] = wv;
            }
            lp += 32;
            while (lp >= LP) { lp -= LP; ++n; }
        }
    }
}

}  // namespace far3d

using namespace far3d;

// kernel variant (tools / tests): warps per CTA (4 | 8 | 2); `wide` bit 0: 256-bit two-sample loads (default) or the 128-bit
// one-sample form; bit 1: one resident wave of CTAs pulling work items from a device-side queue instead of one CTA per item
// (measured 1-2 us slower at cfg-2, profiles/r2y_agg_variants.txt: the pull costs more than the shorter tail saves); bit 2: 4
// instead of 8 two-sample loads in flight per lane (64 registers and 8 CTAs per SM instead of 128 and 4: ~8 % slower)
static int g_da_warps = 4, g_da_wide = 1, g_da_static = 1, g_da_u8 = 1, g_da_prep256 = 0;
extern "C" void far3d_deform_agg_tune(int warps, int wide) {
    g_da_warps = (warps == 8 || warps == 2) ? warps : 4;
    g_da_wide = (wide & 1) ? 1 : 0;
    g_da_static = (wide & 2) ? 0 : 1;
    g_da_u8 = (wide & 4) ? 0 : 1;
    g_da_prep256 = (wide & 8) ? 1 : 0;       // far3d_dfa_prepare with 256-thread CTAs
}

// Work-queue counters {next item, finished CTAs} of the dynamic grid: 64 pairs handed out round-robin, so launches in flight
// on different streams (or baked into different CUDA graphs) do not share a pair; the kernel leaves its pair at {0, 0}.
namespace far3d { __device__ int g_da_sched[64][2]; }
static int da_num_sms() {
    static int n[16] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return 148;
    if (!n[dev] && (cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n[dev] <= 0)) n[dev] = 148;
    return n[dev];
}
static int* da_sched_slot() {
    static int* base[16] = {nullptr};
    static std::atomic<unsigned> next{0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    if (!base[dev]) {
        void* p = nullptr;
        if (cudaGetSymbolAddress(&p, far3d::g_da_sched) != cudaSuccess) return nullptr;
        base[dev] = (int*)p;
    }
    return base[dev] + 2 * (next.fetch_add(1) % 64u);
}

static int fill_levels(LevelInfo& lv, const int32_t* hw_host, const int32_t* start_host, int L, int S) {
    if (L < 1 || L > FAR3D_MAX_LEVELS) return fail(FAR3D_E_UNSUPPORTED, "%snum_levels %ld out of range", "", L);
    for (int l = 0; l < L; ++l) {
        lv.H[l] = hw_host[2 * l]; lv.W[l] = hw_host[2 * l + 1];
        lv.start[l] = start_host ? start_host[l] : 0;
        if (lv.H[l] <= 0 || lv.W[l] <= 0) return fail(FAR3D_E_INVALID, "%sbad level shape", "");
        if (start_host && S > 0 && lv.start[l] + (long)lv.H[l] * lv.W[l] > S)
            return fail(FAR3D_E_INVALID, "%slevel %ld exceeds S=%ld", "", l, S);
    }
    return FAR3D_OK;
}

#ifdef FAR3D_DA_PHASES
extern "C" int far3d_deform_agg_phases(long long* buf) {
    return cudaMemcpyToSymbol(far3d::g_da_phases, &buf, sizeof(buf)) == cudaSuccess ? 0 : 1;
}
#endif

static int agg_launch(const void* feat, int feat_dtype, const int32_t* hw_host, const int32_t* start_host, const float* key_points,
                      const float* lidar2img, const float* weights, float pad_h, float pad_w, float* out, int B, int N, int S, int C,
                      int G, int Nq, int L, int P, const int32_t* pre_cnt, const void* pre_rec, const float* pre_w, void* stream) {
    FAR3D_REQUIRE(feat && hw_host && start_host && out, "null pointer");
    FAR3D_REQUIRE(pre_cnt ? (pre_rec && pre_w) : (key_points && lidar2img && weights), "null pointer");
    FAR3D_REQUIRE(B > 0 && N > 0 && S > 0 && C > 0 && G > 0 && Nq > 0 && L > 0 && P > 0, "non-positive size");
    FAR3D_REQUIRE(C % G == 0, "C must be divisible by G");
    FAR3D_REQUIRE(feat_dtype >= 0 && feat_dtype <= 2, "feat_dtype must be 0 (fp32), 1 (bf16) or 2 (fp16)");
    FAR3D_REQUIRE((long)N * S < (1L << 31), "N*S must fit int32");
    FAR3D_REQUIRE((long)N * Nq * G * L * P < (1L << 31), "N*Nq*G*L*P must fit int32");
    LevelInfo lv;
    int rc = fill_levels(lv, hw_host, start_host, L, S);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int D = C / G;
    const int E = N * L * P;
    const bool fast = (D == 32) && N <= DA_MAX_CAMS && P <= 64 && G <= 16 && (long)N * S * (C / 4) < (1L << 32) &&
                      ((uintptr_t)feat % 32 == 0) && ((uintptr_t)out % 16 == 0) && ((uintptr_t)lidar2img % 16 == 0);
    if (fast) {
        // default: 4 warps per CTA (a query's 8 groups over two CTAs), 256-bit two-sample loads; see DESIGN.md 4.2
        // work item = (query, G / parts channel groups); finer items (2 warps, 4 parts) shorten the last, partly filled wave
        const int warps = pre_cnt ? 4 : (g_da_warps == 8 || G < 8) ? 8 : (g_da_warps == 2 && G % 4 == 0 && g_da_wide) ? 2 : 4;
        const int parts = warps == 8 ? 1 : warps == 2 ? 4 : 2;
        const int ng = G > warps * parts ? 2 : 1;
        const size_t smem = (size_t)N * P * sizeof(PairRec) +
                            (size_t)DA_CHUNK * (4 * sizeof(Corner) + sizeof(int) + warps * sizeof(float));
        const int items = B * Nq * parts;
        const bool u8 = g_da_u8 && warps == 4 && (pre_cnt || g_da_wide);
        const int wave = da_num_sms() * ((u8 ? 16 : 32) / warps);     // resident CTAs (__launch_bounds__)
        int* sched = (!g_da_static && items > wave) ? da_sched_slot() : nullptr;
        FAR3D_REQUIRE(g_da_static || items <= wave || sched, "work-queue counters unavailable");
        const int grid = sched ? wave : items;
#define FAR3D_DA_LAUNCH(T, UU, WW, WD, NG, PR)                                                                          \
    do {                                                                                                                \
        if (smem > 48 * 1024)                                                                                           \
            cudaFuncSetAttribute(deform_agg_kernel<T, UU, WW, WD, NG, PR>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                 (int)smem);                                                                            \
        deform_agg_kernel<T, UU, WW, WD, NG, PR><<<grid, WW * 32, smem, st>>>(                                          \
            (const T*)feat, lv, key_points, lidar2img, weights, pad_h, pad_w, out, B, N, S, C, G, Nq, L, P, parts, sched, \
            pre_cnt, (const Corner*)pre_rec, pre_w);                                                                    \
    } while (0)
#define FAR3D_DA_LAUNCH_NG(T, UU, WW, WD, PR)                                                                           \
    do { if (ng == 2) FAR3D_DA_LAUNCH(T, UU, WW, WD, 2, PR); else FAR3D_DA_LAUNCH(T, UU, WW, WD, 1, PR); } while (0)
#define FAR3D_DA_LAUNCH_T(T)                                                                                            \
    do {                                                                                                                \
        if (pre_cnt) { if (u8) FAR3D_DA_LAUNCH_NG(T, 8, 4, true, true); else FAR3D_DA_LAUNCH_NG(T, 4, 4, true, true); } \
        else if (g_da_wide) {                                                                                           \
            if (warps == 8) FAR3D_DA_LAUNCH_NG(T, 4, 8, true, false);                                                   \
            else if (warps == 2) FAR3D_DA_LAUNCH_NG(T, 4, 2, true, false);                                              \
            else if (u8) FAR3D_DA_LAUNCH_NG(T, 8, 4, true, false);                                                      \
            else FAR3D_DA_LAUNCH_NG(T, 4, 4, true, false);                                                              \
        } else {                                                                                                        \
            if (warps == 8) FAR3D_DA_LAUNCH_NG(T, 8, 8, false, false);                                                  \
            else FAR3D_DA_LAUNCH_NG(T, 8, 4, false, false);                                                             \
        }                                                                                                               \
    } while (0)
        if (feat_dtype == 0) FAR3D_DA_LAUNCH_T(float);
        else if (feat_dtype == 1) FAR3D_DA_LAUNCH_T(__nv_bfloat16);
        else FAR3D_DA_LAUNCH_T(__half);
#undef FAR3D_DA_LAUNCH_T
#undef FAR3D_DA_LAUNCH_NG
#undef FAR3D_DA_LAUNCH
        return launched("deform_agg_kernel");
    }
    if (pre_cnt) return fail(FAR3D_E_UNSUPPORTED, "%sprepared aggregation needs 32-channel groups (C / G = %ld)", "", D);
    long total = (long)B * Nq * C;
    if (feat_dtype == 0)
        deform_agg_generic_kernel<float><<<cdiv(total, 256), 256, 0, st>>>((const float*)feat, lv, key_points, lidar2img,
                                                                           weights, pad_h, pad_w, out, B, N, S, C, G, Nq, L, P);
    else if (feat_dtype == 1)
        deform_agg_generic_kernel<__nv_bfloat16><<<cdiv(total, 256), 256, 0, st>>>(
            (const __nv_bfloat16*)feat, lv, key_points, lidar2img, weights, pad_h, pad_w, out, B, N, S, C, G, Nq, L, P);
    else
        deform_agg_generic_kernel<__half><<<cdiv(total, 256), 256, 0, st>>>(
            (const __half*)feat, lv, key_points, lidar2img, weights, pad_h, pad_w, out, B, N, S, C, G, Nq, L, P);
    return launched("deform_agg_generic_kernel");
}

extern "C" int far3d_deform_agg_fwd(const void* feat, int feat_dtype, const int32_t* hw_host, const int32_t* start_host,
                                    const float* key_points, const float* lidar2img, const float* weights, float pad_h,
                                    float pad_w, float* out, int B, int N, int S, int C, int G, int Nq, int L, int P,
                                    void* stream) {
    FAR3D_REQUIRE(feat && hw_host && start_host && key_points && lidar2img && weights && out, "null pointer");
    FAR3D_REQUIRE(B > 0 && N > 0 && S > 0 && C > 0 && G > 0 && Nq > 0 && L > 0 && P > 0, "non-positive size");
    return agg_launch(feat, feat_dtype, hw_host, start_host, key_points, lidar2img, weights, pad_h, pad_w, out, B, N, S, C, G,
                      Nq, L, P, nullptr, nullptr, nullptr, stream);
}

// ---- two-kernel form: far3d_dfa_prepare (softmax + projection + records) -> far3d_deform_agg_gather
extern "C" int far3d_dfa_prepare_supported(int N, int G, int L, int P, int C) {
    const long E = (long)N * L * P;
    const size_t bytes = (size_t)E * sizeof(int);
    return (G > 0 && C % G == 0 && C / G == 32 && N <= DA_MAX_CAMS && P <= 64 && G <= 16 && L <= FAR3D_MAX_LEVELS &&
            E <= 32 * DWS_MAXK && bytes <= 48 * 1024) ? 1 : 0;
}

extern "C" int far3d_dfa_prepare(const float* wq, const float* wc, const float* key_points, const float* lidar2img,
                                 const int32_t* hw_host, const int32_t* start_host, float pad_h, float pad_w, int B, int N, int Nq,
                                 int G, int L, int P, int S, int C, float* weights, int32_t* cnt, void* rec, float* wts,
                                 void* stream) {
    FAR3D_REQUIRE(wq && wc && key_points && lidar2img && hw_host && start_host && cnt && rec && wts, "null pointer");
    FAR3D_REQUIRE(B > 0 && N > 0 && Nq > 0 && G > 0 && L > 0 && P > 0 && S > 0 && C > 0, "non-positive size");
    FAR3D_REQUIRE(far3d_dfa_prepare_supported(N, G, L, P, C), "shape outside the prepared path (see far3d_dfa_prepare_supported)");
    FAR3D_REQUIRE((long)N * S * (C / 4) < (1L << 32), "N*S*C/4 must fit uint32");
    FAR3D_REQUIRE((uintptr_t)rec % 16 == 0 && (uintptr_t)lidar2img % 16 == 0, "rec / lidar2img must be 16-byte aligned");
    LevelInfo lv;
    int rc = fill_levels(lv, hw_host, start_host, L, S);
    if (rc) return rc;
    const long E = (long)N * L * P;
    const size_t bytes = (size_t)E * sizeof(int);
    // 128-thread CTAs: 47 registers x 128 threads let every CTA of a ~1000-query launch be resident at once (10 per SM); with 256
    // threads the launch is 1.4 waves
    dfa_prepare_kernel<<<B * Nq, g_da_prep256 ? 256 : 128, bytes, (cudaStream_t)stream>>>(wq, wc, key_points, lidar2img, lv, pad_h, pad_w, weights, cnt,
                                                                    (Corner*)rec, wts, B, N, Nq, G, L, P, S, C);
    return launched("dfa_prepare_kernel");
}

extern "C" int far3d_deform_agg_gather(const void* feat, int feat_dtype, const int32_t* hw_host, const int32_t* start_host,
                                       const int32_t* cnt, const void* rec, const float* wts, float* out, int B, int N, int S,
                                       int C, int G, int Nq, int L, int P, void* stream) {
    FAR3D_REQUIRE(feat && hw_host && start_host && out && cnt && rec && wts, "null pointer");
    FAR3D_REQUIRE(B > 0 && N > 0 && S > 0 && C > 0 && G > 0 && Nq > 0 && L > 0 && P > 0, "non-positive size");
    FAR3D_REQUIRE((uintptr_t)rec % 16 == 0, "rec must be 16-byte aligned");
    return agg_launch(feat, feat_dtype, hw_host, start_host, nullptr, nullptr, nullptr, 0.f, 0.f, out, B, N, S, C, G, Nq, L, P, cnt,
                      rec, wts, stream);
}

extern "C" int far3d_deform_agg_debug(const int32_t* hw_host, const float* key_points, const float* lidar2img,
                                      float pad_h, float pad_w, float* uv, int32_t* idx, uint8_t* valid, int B, int N,
                                      int Nq, int L, int P, void* stream) {
    FAR3D_REQUIRE(hw_host && key_points && lidar2img, "null pointer");
    FAR3D_REQUIRE(B > 0 && N > 0 && Nq > 0 && L > 0 && P > 0, "non-positive size");
    LevelInfo lv;
    int rc = fill_levels(lv, hw_host, nullptr, L, 0);
    if (rc) return rc;
    long total = (long)B * N * Nq * P;
    deform_agg_debug_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(lv, key_points, lidar2img, pad_h, pad_w,
                                                                              uv, idx, valid, B, N, Nq, L, P);
    return launched("deform_agg_debug_kernel");
}

extern "C" int far3d_msda_fwd(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                              const float* sampling_locations, const float* attention_weights, float* out, int BN,
                              int S, int G, int D, int Nq, int L, int P, void* stream) {
    FAR3D_REQUIRE(value && spatial_shapes && level_start_index && sampling_locations && attention_weights && out,
                  "null pointer");
    FAR3D_REQUIRE(BN > 0 && S > 0 && G > 0 && D > 0 && Nq > 0 && L > 0 && P > 0, "non-positive size");
    cudaStream_t st = (cudaStream_t)stream;
    if (D == 32 && (uintptr_t)value % 16 == 0 && (uintptr_t)out % 16 == 0) {
        long warps = (long)BN * Nq * G;
        msda_d32_kernel<<<cdiv(warps * 32, 256), 256, 0, st>>>(value, spatial_shapes, level_start_index,
                                                              sampling_locations, attention_weights, out, BN, S, G, Nq, L, P);
        return launched("msda_d32_kernel");
    }
    long total = (long)BN * Nq * G * D;
    msda_generic_kernel<<<cdiv(total, 256), 256, 0, st>>>(value, spatial_shapes, level_start_index, sampling_locations,
                                                         attention_weights, out, BN, S, G, D, Nq, L, P);
    return launched("msda_generic_kernel");
}

extern "C" int far3d_dfa_weights_softmax(const float* wq, const float* wc, float* weights, int B, int N, int Nq, int G,
                                         int LP, void* stream) {
    FAR3D_REQUIRE(wq && wc && weights, "null pointer");
    FAR3D_REQUIRE(B > 0 && N > 0 && Nq > 0 && G > 0 && LP > 0, "non-positive size");
    const size_t dws_bytes = (size_t)(G * LP + (size_t)G * N * LP) * sizeof(float);
    if (N * LP <= 32 * DWS_MAXK && dws_bytes <= 48 * 1024) {
        dfa_weights_softmax_q_kernel<<<B * Nq, 256, dws_bytes, (cudaStream_t)stream>>>(wq, wc, weights, B, N, Nq, G, LP);
        return launched("dfa_weights_softmax_q_kernel");
    }
    long warps = (long)B * Nq * G;
    dfa_weights_softmax_kernel<<<cdiv(warps * 32, 256), 256, 0, (cudaStream_t)stream>>>(wq, wc, weights, B, N, Nq, G, LP);
    return launched("dfa_weights_softmax_kernel");
}
