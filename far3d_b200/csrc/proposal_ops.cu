// 2D proposals -> adaptive 3D queries on the device, with fixed capacities and no host round trip until the count is needed.
//
// Replaces the data-dependent torch glue of
//   models/dense_heads/yolox_head.py:355-489   get_bboxes: sigmoid(obj) * sigmoid(max cls), 3x3 local-max peak pick, score
//                                              threshold, boolean gathers (:454, :467), box decode (:491-501)
//   models/dense_heads/farhead.py:710-827      build_query2d_proposal: depth-bin lookup at the box centre, top-k depth bins,
//                                              multi-depth duplication, un-projection with inverse(lidar2img), pc_range
//                                              normalisation, context features (farhead.py:585-590) + the 2D score channel
// whose boolean-mask gathers and len() / .item() calls force a device->host synchronisation each.  Here:
//   roi_score_kernel     score map w = sigmoid(obj) * sigmoid(max_c cls) of every pixel of every level (one pass, NHWC inputs)
//   roi_select_kernel    one CTA per camera: peak test against the 3x3 neighbourhood, threshold, block-scan compaction in the
//                        reference's order (level-major, then row-major) into per-camera slots of fixed capacity; box decode
//   query2d_lift_kernel  one CTA: camera offsets, depth softmax + top-k at the box centre, multi-depth duplicates in the
//                        reference's order (all primaries, then depth rank 1 of the qualifying ones, ...), un-projection with
//                        the 4x4 inverse (fp64 adjugate, per camera), normalisation; writes the query count
//   ctx_gather_kernel    context rows: feat_flatten[row of the peak] ++ score channel
// Everything downstream (FarHead) runs at a padded, bucketed query count; the padded rows are masked as attention keys.
#include "common.cuh"

namespace far3d {

constexpr int PR_MAX_LEVELS = FAR3D_MAX_LEVELS;
constexpr int PR_MAX_CAMS = 16;
constexpr int PR_MAX_TOPK = 4;

struct RoiLevels {
    const float* cls[PR_MAX_LEVELS];   // [N, H, W, cls_cs] fp32 NHWC
    const float* reg[PR_MAX_LEVELS];   // [N, H, W, reg_cs]: 0-3 box, 4 objectness, 5-6 centre offsets
    int H[PR_MAX_LEVELS], W[PR_MAX_LEVELS], start[PR_MAX_LEVELS], stride[PR_MAX_LEVELS];
    int L, S2;
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// position p in [0, S2) of one camera -> level, y, x
__device__ __forceinline__ void locate(const RoiLevels& lv, int p, int& l, int& y, int& x) {
    l = 0;
#pragma unroll
    for (int i = 1; i < PR_MAX_LEVELS; ++i)
        if (i < lv.L && p >= lv.start[i]) l = i;
    const int r = p - lv.start[l];
    y = r / lv.W[l];
    x = r - y * lv.W[l];
}

__global__ void __launch_bounds__(256)
roi_score_kernel(RoiLevels lv, int N, int num_classes, int cls_cs, int reg_cs, float* __restrict__ score) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)N * lv.S2) return;
    const int n = (int)(idx / lv.S2), p = (int)(idx - (long)n * lv.S2);
    int l, y, x;
    locate(lv, p, l, y, x);
    const size_t pix = ((size_t)n * lv.H[l] + y) * lv.W[l] + x;
    const float* c = lv.cls[l] + pix * cls_cs;
    float m = -INFINITY;
    for (int k = 0; k < num_classes; ++k) m = fmaxf(m, __ldg(c + k));
    const float obj = __ldg(lv.reg[l] + pix * reg_cs + 4);
    score[idx] = sigmoidf_(obj) * sigmoidf_(m);          // == obj.sigmoid() * cls.amax(1).sigmoid()  (yolox_head.py:455)
}

constexpr int SEL_THREADS = 1024;

// grid = N cameras.  Peak: w == max over the 3x3 window (F.max_pool2d(w, 3, 1, 1): out-of-map neighbours do not count).
__global__ void __launch_bounds__(SEL_THREADS)
roi_select_kernel(RoiLevels lv, int reg_cs, const float* __restrict__ score, float threshold, int cap,
                  int32_t* __restrict__ sel_pos, float* __restrict__ sel_score, float* __restrict__ sel_box,
                  int32_t* __restrict__ counts) {
    __shared__ int s_warp[SEL_THREADS / 32];
    __shared__ int s_run;
    const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* sc = score + (size_t)n * lv.S2;
    if (tid == 0) s_run = 0;
    __syncthreads();
    for (int p0 = 0; p0 < lv.S2; p0 += SEL_THREADS) {
        const int p = p0 + tid;
        bool keep = false;
        float w = 0.f;
        int l = 0, y = 0, x = 0;
        if (p < lv.S2) {
            locate(lv, p, l, y, x);
            w = sc[p];
            float wn = w;
            const int H = lv.H[l], W = lv.W[l];
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    const int yy = y + dy, xx = x + dx;
                    if (yy >= 0 && yy < H && xx >= 0 && xx < W) wn = fmaxf(wn, sc[lv.start[l] + yy * W + xx]);
                }
            keep = (w == wn) && (w > threshold);           // score = w * (w == wn); valid = score > threshold (:461-467)
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int base = s_run;
        for (int i = 0; i < warp; ++i) base += s_warp[i];
        const int slot = base + __popc(bal & ((1u << lane) - 1u));
        if (keep && slot < cap) {
            const size_t pix = ((size_t)n * lv.H[l] + y) * lv.W[l] + x;
            const float4 bp = __ldg(reinterpret_cast<const float4*>(lv.reg[l] + pix * reg_cs));
            const float s = (float)lv.stride[l];
            // bbox decode (yolox_head.py:491-501) then corner -> centre form, in the reference's operation order
            const float cx0 = bp.x * s + (float)x * s, cy0 = bp.y * s + (float)y * s;
            const float bw = expf(bp.z) * s, bh = expf(bp.w) * s;
            const float x1 = cx0 - bw / 2, y1 = cy0 - bh / 2, x2 = cx0 + bw / 2, y2 = cy0 + bh / 2;
            const size_t o = (size_t)n * cap + slot;
            sel_pos[o] = p;
            sel_score[o] = w;
            reinterpret_cast<float4*>(sel_box)[o] = make_float4((x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1);
        }
        __syncthreads();
        if (tid == 0) {
            int t = s_run;
            for (int i = 0; i < SEL_THREADS / 32; ++i) t += s_warp[i];
            s_run = t;
        }
        __syncthreads();
    }
    if (tid == 0) counts[n] = s_run;                        // may exceed cap: the caller sees the overflow
}

struct LiftParams {
    int N, cap, S, Hd, Wd, D, Dcs, down, topk, rmin_bin, cap_total;
    float dmin, bin_size, thr_logit;
};

// 4x4 inverse by the adjugate in fp64 (the reference calls torch.inverse on fp32 data: LU with its own rounding; cond ~1e3-1e4)
__device__ void inv4(const float* __restrict__ m, float* __restrict__ out) {
    double a[16], inv[16];
    for (int i = 0; i < 16; ++i) a[i] = (double)m[i];
    inv[0] = a[5] * a[10] * a[15] - a[5] * a[11] * a[14] - a[9] * a[6] * a[15] + a[9] * a[7] * a[14] + a[13] * a[6] * a[11] - a[13] * a[7] * a[10];
    inv[4] = -a[4] * a[10] * a[15] + a[4] * a[11] * a[14] + a[8] * a[6] * a[15] - a[8] * a[7] * a[14] - a[12] * a[6] * a[11] + a[12] * a[7] * a[10];
    inv[8] = a[4] * a[9] * a[15] - a[4] * a[11] * a[13] - a[8] * a[5] * a[15] + a[8] * a[7] * a[13] + a[12] * a[5] * a[11] - a[12] * a[7] * a[9];
    inv[12] = -a[4] * a[9] * a[14] + a[4] * a[10] * a[13] + a[8] * a[5] * a[14] - a[8] * a[6] * a[13] - a[12] * a[5] * a[10] + a[12] * a[6] * a[9];
    inv[1] = -a[1] * a[10] * a[15] + a[1] * a[11] * a[14] + a[9] * a[2] * a[15] - a[9] * a[3] * a[14] - a[13] * a[2] * a[11] + a[13] * a[3] * a[10];
    inv[5] = a[0] * a[10] * a[15] - a[0] * a[11] * a[14] - a[8] * a[2] * a[15] + a[8] * a[3] * a[14] + a[12] * a[2] * a[11] - a[12] * a[3] * a[10];
    inv[9] = -a[0] * a[9] * a[15] + a[0] * a[11] * a[13] + a[8] * a[1] * a[15] - a[8] * a[3] * a[13] - a[12] * a[1] * a[11] + a[12] * a[3] * a[9];
    inv[13] = a[0] * a[9] * a[14] - a[0] * a[10] * a[13] - a[8] * a[1] * a[14] + a[8] * a[2] * a[13] + a[12] * a[1] * a[10] - a[12] * a[2] * a[9];
    inv[2] = a[1] * a[6] * a[15] - a[1] * a[7] * a[14] - a[5] * a[2] * a[15] + a[5] * a[3] * a[14] + a[13] * a[2] * a[7] - a[13] * a[3] * a[6];
    inv[6] = -a[0] * a[6] * a[15] + a[0] * a[7] * a[14] + a[4] * a[2] * a[15] - a[4] * a[3] * a[14] - a[12] * a[2] * a[7] + a[12] * a[3] * a[6];
    inv[10] = a[0] * a[5] * a[15] - a[0] * a[7] * a[13] - a[4] * a[1] * a[15] + a[4] * a[3] * a[13] + a[12] * a[1] * a[7] - a[12] * a[3] * a[5];
    inv[14] = -a[0] * a[5] * a[14] + a[0] * a[6] * a[13] + a[4] * a[1] * a[14] - a[4] * a[2] * a[13] - a[12] * a[1] * a[6] + a[12] * a[2] * a[5];
    inv[3] = -a[1] * a[6] * a[11] + a[1] * a[7] * a[10] + a[5] * a[2] * a[11] - a[5] * a[3] * a[10] - a[9] * a[2] * a[7] + a[9] * a[3] * a[6];
    inv[7] = a[0] * a[6] * a[11] - a[0] * a[7] * a[10] - a[4] * a[2] * a[11] + a[4] * a[3] * a[10] + a[8] * a[2] * a[7] - a[8] * a[3] * a[6];
    inv[11] = -a[0] * a[5] * a[11] + a[0] * a[7] * a[9] + a[4] * a[1] * a[11] - a[4] * a[3] * a[9] - a[8] * a[1] * a[7] + a[8] * a[3] * a[5];
    inv[15] = a[0] * a[5] * a[10] - a[0] * a[6] * a[9] - a[4] * a[1] * a[10] + a[4] * a[2] * a[9] + a[8] * a[1] * a[6] - a[8] * a[2] * a[5];
    const double det = a[0] * inv[0] + a[1] * inv[4] + a[2] * inv[8] + a[3] * inv[12];
    const double r = 1.0 / det;
    for (int i = 0; i < 16; ++i) out[i] = (float)(inv[i] * r);
}

// one proposal row -> normalised 3D reference point (farhead.py:789-826)
__device__ __forceinline__ void lift_point(const LiftParams& q, const float* __restrict__ i2l, const float* __restrict__ pcr,
                                           float cx, float cy, int bin, float* __restrict__ out3) {
    const float t = (float)bin / 0.5f + 1.f;
    const float d = q.dmin + q.bin_size / 8.f * (t * t - 1.f);                 // _convert_bin_depth_to_specific (:521-531)
    const float dd = fmaxf(d, 1e-5f);
    const float c[4] = {cx * dd, cy * dd, d, 1.f};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float a = i2l[4 * r] * c[0];
        a = fmaf(i2l[4 * r + 1], c[1], a);
        a = fmaf(i2l[4 * r + 2], c[2], a);
        a = fmaf(i2l[4 * r + 3], c[3], a);
        out3[r] = (a - pcr[r]) / (pcr[3 + r] - pcr[r]);
    }
}

constexpr int LIFT_THREADS = 1024;

// one CTA.  ws: int [cap_total] ok flags | int [cap_total] rank among ok | int [cap_total * (topk-1)] bins | float [same] ds
__global__ void __launch_bounds__(LIFT_THREADS)
query2d_lift_kernel(LiftParams q, const int32_t* __restrict__ sel_pos, const float* __restrict__ sel_score,
                    const float* __restrict__ sel_box, const int32_t* __restrict__ counts,
                    const float* __restrict__ depth_logits, const float* __restrict__ lidar2img,
                    const float* __restrict__ pc_range, float* __restrict__ ref2d, int32_t* __restrict__ src_row,
                    float* __restrict__ score_feat, int32_t* __restrict__ meta, int32_t* __restrict__ ws) {
    __shared__ float s_i2l[PR_MAX_CAMS][16];
    __shared__ float s_pcr[6];
    __shared__ int s_off[PR_MAX_CAMS + 1];
    __shared__ int s_warp[LIFT_THREADS / 32];
    __shared__ int s_run;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < q.N) inv4(lidar2img + tid * 16, s_i2l[tid]);
    if (tid < 6) s_pcr[tid] = pc_range[tid];
    if (tid == 0) {
        int t = 0, over = 0;
        for (int i = 0; i < q.N; ++i) {
            s_off[i] = t;
            const int c = counts[i];
            if (c > q.cap) over = 1;
            t += min(c, q.cap);
        }
        if (t > q.cap_total) { t = q.cap_total; over = 1; }
        s_off[q.N] = t;
        s_run = 0;
        meta[3] = over;
    }
    __syncthreads();
    const int M = s_off[q.N];
    int* w_ok = ws;
    int* w_rank = ws + q.cap_total;
    int* w_bin = ws + 2 * q.cap_total;
    float* w_ds = reinterpret_cast<float*>(ws + 2 * q.cap_total + q.cap_total * (PR_MAX_TOPK - 1));
    const int K1 = q.topk - 1;
    for (int g0 = 0; g0 < M; g0 += LIFT_THREADS) {
        const int g = g0 + tid;
        bool ok = false;
        if (g < M) {
            int cam = 0;
            for (int i = 1; i < q.N; ++i)
                if (g >= s_off[i]) cam = i;
            const size_t o = (size_t)cam * q.cap + (g - s_off[cam]);
            const float4 box = __ldg(reinterpret_cast<const float4*>(sel_box) + o);
            // depth bins at the box centre: (bb[:, :2] / down).round().long().clamp(...)  (farhead.py:731-737)
            int cxi = (int)rintf(box.x / (float)q.down), cyi = (int)rintf(box.y / (float)q.down);
            cxi = min(max(cxi, 0), q.Wd - 1);
            cyi = min(max(cyi, 0), q.Hd - 1);
            const float* lg = depth_logits + (((size_t)cam * q.Hd + cyi) * q.Wd + cxi) * q.Dcs;
            float mx = -INFINITY;
            for (int k = 0; k < q.D; ++k) mx = fmaxf(mx, __ldg(lg + k));
            float sum = 0.f;
            for (int k = 0; k < q.D; ++k) sum += expf(__ldg(lg + k) - mx);
            int ti[PR_MAX_TOPK];
            float tv[PR_MAX_TOPK];
            for (int r = 0; r < q.topk; ++r) {                                  // top-k of the softmax = top-k of the logits
                float best = -INFINITY;
                int bi = 0;
                for (int k = 0; k < q.D; ++k) {
                    bool used = false;
                    for (int s = 0; s < r; ++s) used |= ti[s] == k;
                    const float v = __ldg(lg + k);
                    if (!used && v > best) { best = v; bi = k; }
                }
                ti[r] = bi;
                tv[r] = expf(best - mx) / sum;
            }
            ok = ti[0] >= q.rmin_bin;                                           // :744
            lift_point(q, s_i2l[cam], s_pcr, box.x, box.y, ti[0], ref2d + (size_t)g * 3);
            src_row[g] = cam * q.S + sel_pos[o];
            const float s = sel_score[o];
            score_feat[g] = (logf(s / (1.f - s)) - q.thr_logit) * (tv[0] / tv[0]);     // lo * ds, ds = tv / tv[:, 0:1]  (:757-763)
            w_ok[g] = ok ? 1 : 0;
            for (int r = 0; r < K1; ++r) { w_bin[g * (PR_MAX_TOPK - 1) + r] = ti[r + 1]; w_ds[g * (PR_MAX_TOPK - 1) + r] = tv[r + 1] / tv[0]; }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int base = s_run;
        for (int i = 0; i < warp; ++i) base += s_warp[i];
        if (g < M) w_rank[g] = base + __popc(bal & ((1u << lane) - 1u));
        __syncthreads();
        if (tid == 0) {
            int t = s_run;
            for (int i = 0; i < LIFT_THREADS / 32; ++i) t += s_warp[i];
            s_run = t;
        }
        __syncthreads();
    }
    const int n_ok = K1 > 0 ? s_run : 0;
    // multi-depth duplicates, k-major after all primaries (boxes.repeat(topk - 1, 1)[ok.repeat(topk - 1)], :745-751)
    int total = M + K1 * n_ok;
    if (K1 > 0) {
        for (int g = tid; g < M; g += LIFT_THREADS) {
            if (!w_ok[g]) continue;
            int cam = 0;
            for (int i = 1; i < q.N; ++i)
                if (g >= s_off[i]) cam = i;
            const size_t o = (size_t)cam * q.cap + (g - s_off[cam]);
            const float4 box = __ldg(reinterpret_cast<const float4*>(sel_box) + o);
            const float s = sel_score[o];
            const float lo = logf(s / (1.f - s)) - q.thr_logit;
            for (int r = 0; r < K1; ++r) {
                const int e = M + r * n_ok + w_rank[g];
                if (e >= q.cap_total) continue;
                lift_point(q, s_i2l[cam], s_pcr, box.x, box.y, w_bin[g * (PR_MAX_TOPK - 1) + r], ref2d + (size_t)e * 3);
                src_row[e] = src_row[g];
                score_feat[e] = lo * w_ds[g * (PR_MAX_TOPK - 1) + r];
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        if (total > q.cap_total) { total = q.cap_total; meta[3] = 1; }
        meta[0] = total; meta[1] = M; meta[2] = n_ok;
    }
    // padding rows: a harmless point in the middle of the range (masked as attention keys, dropped from the outputs)
    for (int e = total + tid; e < q.cap_total; e += LIFT_THREADS) {
        ref2d[(size_t)e * 3] = 0.5f; ref2d[(size_t)e * 3 + 1] = 0.5f; ref2d[(size_t)e * 3 + 2] = 0.5f;
        src_row[e] = -1;
        score_feat[e] = 0.f;
    }
}

// ctx[r] = feat_flatten[src_row[r]] ++ score_feat[r]   (zeros for padding rows); 4 channels per thread
__global__ void __launch_bounds__(256)
ctx_gather_kernel(const float* __restrict__ feat, const int32_t* __restrict__ src_row, const float* __restrict__ score_feat,
                  int C, int rows, float* __restrict__ ctx) {
    const int C4 = C >> 2;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)rows * (C4 + 1)) return;
    const int r = (int)(idx / (C4 + 1)), c = (int)(idx - (long)r * (C4 + 1));
    const int src = src_row[r];
    float* o = ctx + (size_t)r * (C + 1);
    if (c == C4) { o[C] = src >= 0 ? score_feat[r] : 0.f; return; }
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (src >= 0) v = __ldg(reinterpret_cast<const float4*>(feat + (size_t)src * C) + c);
    o[4 * c] = v.x; o[4 * c + 1] = v.y; o[4 * c + 2] = v.z; o[4 * c + 3] = v.w;     // row stride C + 1: not 16-byte aligned
}

}  // namespace far3d

using namespace far3d;

static int fill_roi_levels(RoiLevels& lv, const void* const* cls_host, const void* const* reg_host, const int32_t* hw_host,
                           const int32_t* stride_host, int L) {
    if (L < 1 || L > PR_MAX_LEVELS) return fail(FAR3D_E_UNSUPPORTED, "%snum_levels %ld out of range", "", L);
    int s = 0;
    for (int l = 0; l < L; ++l) {
        lv.cls[l] = (const float*)cls_host[l]; lv.reg[l] = (const float*)reg_host[l];
        lv.H[l] = hw_host[2 * l]; lv.W[l] = hw_host[2 * l + 1]; lv.stride[l] = stride_host[l];
        lv.start[l] = s;
        if (!lv.cls[l] || !lv.reg[l] || lv.H[l] <= 0 || lv.W[l] <= 0) return fail(FAR3D_E_INVALID, "%sbad level", "");
        if ((uintptr_t)lv.reg[l] % 16) return fail(FAR3D_E_INVALID, "%sreg maps must be 16-byte aligned", "");
        s += lv.H[l] * lv.W[l];
    }
    lv.L = L; lv.S2 = s;
    return FAR3D_OK;
}

extern "C" int far3d_roi_select(const void* const* cls_host, const void* const* reg_host, const int32_t* hw_host,
                                const int32_t* stride_host, int L, int N, int num_classes, int cls_cs, int reg_cs,
                                float threshold, float* score_ws, int cap_per_cam, int32_t* sel_pos, float* sel_score,
                                float* sel_box, int32_t* counts, void* stream) {
    FAR3D_REQUIRE(cls_host && reg_host && hw_host && stride_host && score_ws && sel_pos && sel_score && sel_box && counts, "null pointer");
    FAR3D_REQUIRE(N > 0 && N <= PR_MAX_CAMS && num_classes > 0 && cls_cs >= num_classes && reg_cs >= 5 && reg_cs % 4 == 0 && cap_per_cam > 0,
                  "bad sizes");
    FAR3D_REQUIRE((uintptr_t)sel_box % 16 == 0, "sel_box must be 16-byte aligned");
    RoiLevels lv;
    int rc = fill_roi_levels(lv, cls_host, reg_host, hw_host, stride_host, L);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    roi_score_kernel<<<cdiv((long)N * lv.S2, 256), 256, 0, st>>>(lv, N, num_classes, cls_cs, reg_cs, score_ws);
    if ((rc = launched("roi_score_kernel"))) return rc;
    roi_select_kernel<<<N, SEL_THREADS, 0, st>>>(lv, reg_cs, score_ws, threshold, cap_per_cam, sel_pos, sel_score, sel_box, counts);
    return launched("roi_select_kernel");
}

extern "C" int far3d_query2d_lift(const int32_t* sel_pos, const float* sel_score, const float* sel_box, const int32_t* counts,
                                  int N, int cap_per_cam, int S, const float* depth_logits, int Hd, int Wd, int D, int Dcs,
                                  int down, int topk, int rmin_bin, float dmin, float bin_size, float thr_logit,
                                  const float* lidar2img, const float* pc_range, int cap_total, float* ref2d,
                                  int32_t* src_row, float* score_feat, int32_t* meta, int32_t* workspace, void* stream) {
    FAR3D_REQUIRE(sel_pos && sel_score && sel_box && counts && depth_logits && lidar2img && pc_range && ref2d && src_row && score_feat &&
                      meta && workspace, "null pointer");
    FAR3D_REQUIRE(N > 0 && N <= PR_MAX_CAMS && cap_per_cam > 0 && cap_total > 0 && D > 0 && Dcs >= D && down > 0, "bad sizes");
    FAR3D_REQUIRE(topk >= 1 && topk <= PR_MAX_TOPK && topk <= D, "topk must be in 1..4");
    FAR3D_REQUIRE((uintptr_t)sel_box % 16 == 0, "sel_box must be 16-byte aligned");
    LiftParams q;
    q.N = N; q.cap = cap_per_cam; q.S = S; q.Hd = Hd; q.Wd = Wd; q.D = D; q.Dcs = Dcs; q.down = down; q.topk = topk;
    q.rmin_bin = rmin_bin; q.cap_total = cap_total; q.dmin = dmin; q.bin_size = bin_size; q.thr_logit = thr_logit;
    query2d_lift_kernel<<<1, LIFT_THREADS, 0, (cudaStream_t)stream>>>(q, sel_pos, sel_score, sel_box, counts, depth_logits, lidar2img,
                                                                   pc_range, ref2d, src_row, score_feat, meta, workspace);
    return launched("query2d_lift_kernel");
}

extern "C" int64_t far3d_query2d_lift_workspace_ints(int cap_total) { return (int64_t)cap_total * (2 + 2 * (PR_MAX_TOPK - 1)); }

extern "C" int far3d_ctx_gather(const float* feat_flatten, const int32_t* src_row, const float* score_feat, int C, int rows,
                                float* ctx, void* stream) {
    FAR3D_REQUIRE(feat_flatten && src_row && score_feat && ctx && C > 0 && C % 4 == 0 && rows > 0, "bad argument");
    FAR3D_REQUIRE((uintptr_t)feat_flatten % 16 == 0, "feat_flatten must be 16-byte aligned");
    ctx_gather_kernel<<<cdiv((long)rows * (C / 4 + 1), 256), 256, 0, (cudaStream_t)stream>>>(feat_flatten, src_row, score_feat, C, rows, ctx);
    return launched("ctx_gather_kernel");
}
