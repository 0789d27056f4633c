// Temporal memory bank of FarHead on the device: two kernels per frame instead of ~35 torch launches (top-k, gathers,
// concatenations, 4x4 pose products, timestamp arithmetic).
//
// Replaces models/dense_heads/farhead.py:446-508:
//   post_update_memory (:479-508)  rec_score = max-class sigmoid of the last layer, top-`topk_proposals` queries, gathers of
//                                  their embedding / reference point / velocity, push in front of the bank, move every row
//                                  into the next ego frame (ego_pose @ .), timestamps minus the frame's
//   pre_update_memory  (:446-477)  (next frame) timestamps plus the frame's, rows into the current ego frame (ego_pose_inv @ .),
//                                  truncation to memory_len, reset by prev_exists, pseudo reference points for a new scene
// Row math is fp32 fused multiply-adds (the reference: fp32 cuBLAS bmm), timestamps are fp64 as in the reference.
#include "common.cuh"

namespace far3d {

constexpr int MEM_THREADS = 1024;
constexpr int MEM_MAXQ = 4096;

__device__ __forceinline__ uint32_t mem_order_key(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// one CTA: idx[0..K) = queries with the largest max-class logit, descending (ties: lowest index first) == torch.topk of
// sigmoid(cls).max(-1) up to the order of exact ties
__global__ void __launch_bounds__(MEM_THREADS)
memory_topk_kernel(const float* __restrict__ cls, int Nq, int C, int K, int32_t* __restrict__ idx) {
    __shared__ unsigned long long s_key[MEM_MAXQ];
    int n2 = 1;
    while (n2 < Nq) n2 <<= 1;
    for (int q = threadIdx.x; q < n2; q += MEM_THREADS) {
        unsigned long long k = 0ull;
        if (q < Nq) {
            float m = -INFINITY;
            for (int c = 0; c < C; ++c) m = fmaxf(m, __ldg(cls + (size_t)q * C + c));
            k = ((unsigned long long)mem_order_key(m) << 32) | (uint32_t)(0xFFFFFFFFu - (uint32_t)q);
        }
        s_key[q] = k;
    }
    __syncthreads();
    for (int size = 2; size <= n2; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < n2 / 2; t += MEM_THREADS) {
                const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const unsigned long long a = s_key[lo], b = s_key[hi];
                if ((a < b) == desc) { s_key[lo] = b; s_key[hi] = a; }
            }
            __syncthreads();
        }
    for (int r = threadIdx.x; r < K; r += MEM_THREADS) idx[r] = (int32_t)(0xFFFFFFFFu - (uint32_t)(s_key[r] & 0xFFFFFFFFull));
}

__device__ __forceinline__ void mat4_mul(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float s = a[4 * i] * b[j];
            s = fmaf(a[4 * i + 1], b[4 + j], s);
            s = fmaf(a[4 * i + 2], b[8 + j], s);
            s = fmaf(a[4 * i + 3], b[12 + j], s);
            o[4 * i + j] = s;
        }
}
__device__ __forceinline__ void mat4_point(const float* __restrict__ a, const float* __restrict__ p, float* __restrict__ o) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float s = a[4 * i] * p[0];
        s = fmaf(a[4 * i + 1], p[1], s);
        s = fmaf(a[4 * i + 2], p[2], s);
        o[i] = s + a[4 * i + 3];
    }
}

// post_update_memory: new bank of K + M rows.  One warp per row; lanes copy the embedding, lane 0 does the small fields.
__global__ void __launch_bounds__(256)
memory_post_kernel(const int32_t* __restrict__ idx, int K, int M, int E, int code, const float* __restrict__ dec_last,
                   const float* __restrict__ box_last, const float* __restrict__ ego_pose, const double* __restrict__ timestamp,
                   const float* __restrict__ o_emb, const float* __restrict__ o_ref, const double* __restrict__ o_ts,
                   const float* __restrict__ o_pose, const float* __restrict__ o_velo, float* __restrict__ n_emb,
                   float* __restrict__ n_ref, double* __restrict__ n_ts, float* __restrict__ n_pose, float* __restrict__ n_velo) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= K + M) return;
    const bool fresh = row < K;
    const int q = fresh ? idx[row] : 0, m = row - K;
    const float* src = fresh ? dec_last + (size_t)q * E : o_emb + (size_t)m * E;
    for (int c = lane; c < E; c += 32) n_emb[(size_t)row * E + c] = src[c];
    if (lane == 0) {
        float pose[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pose[i] = __ldg(ego_pose + i);
        float p[3], v[2];
        if (fresh) {
            const float* b = box_last + (size_t)q * code;
            p[0] = b[0]; p[1] = b[1]; p[2] = b[2];
            v[0] = b[code - 2]; v[1] = b[code - 1];
#pragma unroll
            for (int i = 0; i < 16; ++i) n_pose[(size_t)row * 16 + i] = pose[i];       // ego_pose @ identity
            n_ts[row] = 0.0 - *timestamp;
        } else {
            p[0] = o_ref[(size_t)m * 3]; p[1] = o_ref[(size_t)m * 3 + 1]; p[2] = o_ref[(size_t)m * 3 + 2];
            v[0] = o_velo[(size_t)m * 2]; v[1] = o_velo[(size_t)m * 2 + 1];
            float old[16], np_[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) old[i] = o_pose[(size_t)m * 16 + i];
            mat4_mul(pose, old, np_);
#pragma unroll
            for (int i = 0; i < 16; ++i) n_pose[(size_t)row * 16 + i] = np_[i];
            n_ts[row] = o_ts[m] - *timestamp;
        }
        float o[3];
        mat4_point(pose, p, o);
        n_ref[(size_t)row * 3] = o[0]; n_ref[(size_t)row * 3 + 1] = o[1]; n_ref[(size_t)row * 3 + 2] = o[2];
        n_velo[(size_t)row * 2] = v[0]; n_velo[(size_t)row * 2 + 1] = v[1];
    }
}

// pre_update_memory on an existing bank of R >= n rows -> n rows.  x = prev_exists (0 | 1).
__global__ void __launch_bounds__(256)
memory_pre_kernel(int n, int E, int kprop, const float* __restrict__ prev_exists, const float* __restrict__ ego_pose_inv,
                  const double* __restrict__ timestamp, const float* __restrict__ pseudo /*[kprop,3] in metres*/,
                  const float* __restrict__ o_emb, const float* __restrict__ o_ref, const double* __restrict__ o_ts,
                  const float* __restrict__ o_pose, const float* __restrict__ o_velo, float* __restrict__ n_emb,
                  float* __restrict__ n_ref, double* __restrict__ n_ts, float* __restrict__ n_pose, float* __restrict__ n_velo) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= n) return;
    const float x = __ldg(prev_exists);
    for (int c = lane; c < E; c += 32) n_emb[(size_t)row * E + c] = o_emb[(size_t)row * E + c] * x;
    if (lane == 0) {
        float inv[16], old[16], np_[16], p[3], o[3];
#pragma unroll
        for (int i = 0; i < 16; ++i) { inv[i] = __ldg(ego_pose_inv + i); old[i] = o_pose[(size_t)row * 16 + i]; }
        mat4_mul(inv, old, np_);
        p[0] = o_ref[(size_t)row * 3]; p[1] = o_ref[(size_t)row * 3 + 1]; p[2] = o_ref[(size_t)row * 3 + 2];
        mat4_point(inv, p, o);
        const bool ps = row < kprop;                         // a new scene starts from the learned pseudo points / identity poses
#pragma unroll
        for (int i = 0; i < 16; ++i)
            n_pose[(size_t)row * 16 + i] = np_[i] * x + (ps ? (1.f - x) * ((i % 5 == 0) ? 1.f : 0.f) : 0.f);
#pragma unroll
        for (int i = 0; i < 3; ++i) n_ref[(size_t)row * 3 + i] = o[i] * x + (ps ? (1.f - x) * pseudo[(size_t)row * 3 + i] : 0.f);
        n_ts[row] = (o_ts[row] + *timestamp) * (double)x;
        n_velo[(size_t)row * 2] = o_velo[(size_t)row * 2] * x;
        n_velo[(size_t)row * 2 + 1] = o_velo[(size_t)row * 2 + 1] * x;
    }
}

}  // namespace far3d

using namespace far3d;

extern "C" int far3d_memory_post_update(const float* cls_last, const float* box_last, const float* dec_last, int Nq, int C,
                                        int code, int E, int K, int M, const float* ego_pose, const double* timestamp,
                                        const float* o_emb, const float* o_ref, const double* o_ts, const float* o_pose,
                                        const float* o_velo, int32_t* topk_idx, float* n_emb, float* n_ref, double* n_ts,
                                        float* n_pose, float* n_velo, void* stream) {
    FAR3D_REQUIRE(cls_last && box_last && dec_last && ego_pose && timestamp && topk_idx && n_emb && n_ref && n_ts && n_pose && n_velo,
                  "null pointer");
    FAR3D_REQUIRE(M == 0 || (o_emb && o_ref && o_ts && o_pose && o_velo), "null memory bank");
    FAR3D_REQUIRE(Nq > 0 && Nq <= MEM_MAXQ && C > 0 && code >= 5 && E > 0 && K > 0 && K <= Nq && M >= 0, "bad sizes (Nq <= 4096)");
    cudaStream_t st = (cudaStream_t)stream;
    memory_topk_kernel<<<1, MEM_THREADS, 0, st>>>(cls_last, Nq, C, K, topk_idx);
    int rc = launched("memory_topk_kernel");
    if (rc) return rc;
    memory_post_kernel<<<cdiv((long)(K + M) * 32, 256), 256, 0, st>>>(topk_idx, K, M, E, code, dec_last, box_last, ego_pose, timestamp, o_emb,
                                                                  o_ref, o_ts, o_pose, o_velo, n_emb, n_ref, n_ts, n_pose, n_velo);
    return launched("memory_post_kernel");
}

extern "C" int far3d_memory_pre_update(int n, int E, int kprop, const float* prev_exists, const float* ego_pose_inv,
                                       const double* timestamp, const float* pseudo_points, const float* o_emb,
                                       const float* o_ref, const double* o_ts, const float* o_pose, const float* o_velo,
                                       float* n_emb, float* n_ref, double* n_ts, float* n_pose, float* n_velo, void* stream) {
    FAR3D_REQUIRE(prev_exists && ego_pose_inv && timestamp && o_emb && o_ref && o_ts && o_pose && o_velo && n_emb && n_ref && n_ts &&
                      n_pose && n_velo, "null pointer");
    FAR3D_REQUIRE(n > 0 && E > 0 && kprop >= 0 && kprop <= n && (kprop == 0 || pseudo_points), "bad sizes");
    memory_pre_kernel<<<cdiv((long)n * 32, 256), 256, 0, (cudaStream_t)stream>>>(n, E, kprop, prev_exists, ego_pose_inv, timestamp,
                                                                              pseudo_points, o_emb, o_ref, o_ts, o_pose, o_velo, n_emb,
                                                                              n_ref, n_ts, n_pose, n_velo);
    return launched("memory_pre_kernel");
}
