// Fused NMS-free box decode: one kernel instead of sigmoid -> view -> topk (radix sort) -> div / mod -> gather -> exp / atan2 ->
// range mask -> boolean gathers -> z shift.
//
// Replaces core/bbox/coders/nms_free_coder.py:39-112 (decode_single: sigmoid, top-`max_num` over queries x classes, label /
// query index, denormalize_bbox core/bbox/util.py:25-52, post_center_range mask, optional score threshold) and the bottom-centre
// shift of FarHead.get_bboxes (models/dense_heads/farhead.py:1224-1245), for one sample.  The number of boxes that survive the
// range mask is data dependent: the kernel writes fixed-size [K, ...] outputs compacted in score order plus the count, so the
// caller moves ONE block to the host and slices there (the reference's bbox3d2result moves the boxes to the host anyway).
//
// One CTA: (1) 4-pass MSB radix select of the K-th largest logit over the Nq * C candidates (sigmoid is monotone, so the
// selection runs on the logits), (2) ordered collection of the candidates above the threshold value and of the lowest-index
// ties, (3) bitonic sort of the K candidates by (score descending, flat index ascending), (4) decode + mask + ordered compaction.
#include "common.cuh"

namespace far3d {

constexpr int DEC_THREADS = 1024;
constexpr int DEC_MAXK = 512;

__device__ __forceinline__ uint32_t order_key(float f) {       // larger float <=> larger unsigned key
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// ordered append: every thread with `take` gets the next free slot in thread order; returns the slot or -1; *run is advanced
__device__ __forceinline__ int ordered_slot(bool take, int* s_warp, int* s_run) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned bal = __ballot_sync(0xffffffffu, take);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int base = *s_run;
    for (int i = 0; i < warp; ++i) base += s_warp[i];
    const int slot = take ? base + __popc(bal & ((1u << lane) - 1u)) : -1;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = *s_run;
        for (int i = 0; i < DEC_THREADS / 32; ++i) t += s_warp[i];
        *s_run = t;
    }
    __syncthreads();
    return slot;
}

struct DecodeParams {
    int Nq, C, code, K, bottom_center;
    float range[6];
    float score_thr;          // <= 0: none
};

__global__ void __launch_bounds__(DEC_THREADS)
box_decode_kernel(DecodeParams p, const float* __restrict__ cls, const float* __restrict__ box, float* __restrict__ out_boxes,
                  float* __restrict__ out_scores, int32_t* __restrict__ out_labels, int32_t* __restrict__ out_query,
                  int32_t* __restrict__ out_count) {
    __shared__ int s_hist[256];
    __shared__ int s_warp[DEC_THREADS / 32];
    __shared__ int s_run;
    __shared__ uint32_t s_prefix, s_mask;
    __shared__ int s_need;
    __shared__ unsigned long long s_cand[DEC_MAXK];
    const int tid = threadIdx.x;
    const int n = p.Nq * p.C;
    const int K = min(p.K, n);
    // ---- (1) radix select: after the passes s_prefix is the key of the K-th largest candidate
    if (tid == 0) { s_prefix = 0; s_mask = 0; s_need = K; }
    __syncthreads();
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = tid; i < 256; i += DEC_THREADS) s_hist[i] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix, mask = s_mask;
        for (int i = tid; i < n; i += DEC_THREADS) {
            const uint32_t k = order_key(__ldg(cls + i));
            if ((k & mask) == prefix) atomicAdd(&s_hist[(k >> shift) & 255u], 1);
        }
        __syncthreads();
        if (tid == 0) {
            int need = s_need, b = 255;
            for (; b > 0; --b) {                           // walk down from the largest digit
                if (s_hist[b] >= need) break;
                need -= s_hist[b];
            }
            s_need = need;                                 // rank of the K-th inside digit b's bucket
            s_prefix = prefix | ((uint32_t)b << shift);
            s_mask = mask | (255u << shift);
        }
        __syncthreads();
    }
    const uint32_t T = s_prefix;
    // ---- (2) candidates: every key > T (in index order), then the first s_need keys == T
    if (tid == 0) s_run = 0;
    for (int i = tid; i < DEC_MAXK; i += DEC_THREADS) s_cand[i] = 0ull;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += DEC_THREADS) {
        const int i = i0 + tid;
        uint32_t k = 0;
        if (i < n) k = order_key(__ldg(cls + i));
        const int slot = ordered_slot(i < n && k > T, s_warp, &s_run);
        if (slot >= 0 && slot < DEC_MAXK) s_cand[slot] = ((unsigned long long)k << 32) | (uint32_t)(0xFFFFFFFFu - (uint32_t)i);
    }
    const int above = s_run;
    const int ties = K - above;
    __syncthreads();
    if (tid == 0) s_run = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += DEC_THREADS) {
        if (s_run >= ties) break;                          // uniform: s_run is read after the barrier inside ordered_slot
        const int i = i0 + tid;
        uint32_t k = 0;
        if (i < n) k = order_key(__ldg(cls + i));
        const int slot = ordered_slot(i < n && k == T, s_warp, &s_run);
        if (slot >= 0 && slot < ties) s_cand[above + slot] = ((unsigned long long)k << 32) | (uint32_t)(0xFFFFFFFFu - (uint32_t)i);
    }
    __syncthreads();
    // ---- (3) bitonic sort, descending, of DEC_MAXK composite keys (unused slots are 0 = smallest)
    for (int size = 2; size <= DEC_MAXK; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < DEC_MAXK / 2; t += DEC_THREADS) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const unsigned long long a = s_cand[lo], b = s_cand[hi];
                if ((a < b) == desc) { s_cand[lo] = b; s_cand[hi] = a; }
            }
            __syncthreads();
        }
    // ---- (4) decode, mask, compact in score order
    if (tid == 0) s_run = 0;
    __syncthreads();
    const int W = p.code > 8 ? 9 : 7;
    for (int t0 = 0; t0 < K; t0 += DEC_THREADS) {
        const int t = t0 + tid;
        bool keep = false;
        float b[10], score = 0.f;
        int label = 0, q = 0;
        if (t < K) {
            const uint32_t idx = 0xFFFFFFFFu - (uint32_t)(s_cand[t] & 0xFFFFFFFFull);
            q = (int)(idx / (uint32_t)p.C);
            label = (int)(idx - (uint32_t)q * (uint32_t)p.C);
            score = 1.f / (1.f + expf(-__ldg(cls + idx)));
            const float* r = box + (size_t)q * p.code;
            b[0] = r[0]; b[1] = r[1]; b[2] = r[2];
            b[3] = expf(r[3]); b[4] = expf(r[4]); b[5] = expf(r[5]);
            b[6] = atan2f(r[6], r[7]);                     // denormalize_bbox (util.py:25-52)
            if (W == 9) { b[7] = r[8]; b[8] = r[9]; }
            keep = b[0] >= p.range[0] && b[1] >= p.range[1] && b[2] >= p.range[2] && b[0] <= p.range[3] && b[1] <= p.range[4] &&
                   b[2] <= p.range[5];
            if (p.score_thr > 0.f) keep = keep && score >= p.score_thr;
        }
        const int slot = ordered_slot(keep, s_warp, &s_run);
        if (slot >= 0) {
            if (p.bottom_center) b[2] = b[2] - b[5] * 0.5f;      // FarHead.get_bboxes: box centre -> bottom centre (farhead.py:1237)
            for (int c = 0; c < W; ++c) out_boxes[(size_t)slot * W + c] = b[c];
            out_scores[slot] = score;
            out_labels[slot] = label;
            out_query[slot] = q;
        }
    }
    if (tid == 0) *out_count = s_run;
}

}  // namespace far3d

using namespace far3d;

extern "C" int far3d_box_decode(const float* cls, const float* box, int Nq, int C, int code, int max_num,
                                const float* post_center_range_host, float score_threshold, int bottom_center, float* out_boxes,
                                float* out_scores, int32_t* out_labels, int32_t* out_query, int32_t* out_count, void* stream) {
    FAR3D_REQUIRE(cls && box && post_center_range_host && out_boxes && out_scores && out_labels && out_query && out_count, "null pointer");
    FAR3D_REQUIRE(Nq > 0 && C > 0 && (code == 8 || code == 10), "bad sizes (code_size 8 or 10)");
    FAR3D_REQUIRE(max_num > 0 && max_num <= DEC_MAXK, "max_num must be in 1..512");
    FAR3D_REQUIRE((long)Nq * C < (1L << 31), "Nq * C must fit int32");
    DecodeParams p;
    p.Nq = Nq; p.C = C; p.code = code; p.K = max_num; p.score_thr = score_threshold; p.bottom_center = bottom_center;
    for (int i = 0; i < 6; ++i) p.range[i] = post_center_range_host[i];
    box_decode_kernel<<<1, DEC_THREADS, 0, (cudaStream_t)stream>>>(p, cls, box, out_boxes, out_scores, out_labels, out_query, out_count);
    return launched("box_decode_kernel");
}
