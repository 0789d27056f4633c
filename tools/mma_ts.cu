// Micro-benchmark + correctness probe: tcgen05.mma with the A operand in TMEM (TS form), staged there from a
// SWIZZLE_128B K-major smem tile with tcgen05.cp.128x256b -- against the SS form (A read from smem by every MMA).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_ts tools/mma_ts.cu && tools/mma_ts
// Part 1: D[128 x N] = A[128 x 64] * B[N x 64]^T with exact small-integer bf16 data, SS vs TS, checked on the host.
// Part 2: cycles per k16 step of the split-bf16 pattern (3 MMAs: hi*hi, hi*lo, lo*hi) in SS form and in TS form
//         (2 tcgen05.cp + 3 TS MMAs), for N = 64..256.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t idesc_bf16(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void cp_128x256b(uint32_t taddr, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// element (row r, k) of a [rows x 64] bf16 K-major SWIZZLE_128B tile: 16-byte chunk c = k/8 lands at chunk c ^ (r & 7)
__device__ __forceinline__ int sw128_index(int r, int k) { return r * 64 + ((((k >> 3) ^ (r & 7)) << 3) | (k & 7)); }

constexpr int A_COL = 256;       // TMEM column where staged A slices start (accumulator occupies columns 0..255)

// mode 0: SS; mode 1: TS (cp each k16 slice of A to TMEM, then MMA with A from TMEM)
__global__ void __launch_bounds__(128, 1) check_kernel(int N, int mode, float* out) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ uint64_t bar;
    __shared__ uint32_t s_tmem;
    __nv_bfloat16* A = (__nv_bfloat16*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    __nv_bfloat16* B = A + 128 * 64;
    for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
        int r = i / 64, k = i % 64;
        A[sw128_index(r, k)] = __float2bfloat16((float)(((r * 3 + k * 5) % 7) - 3));
    }
    for (int i = threadIdx.x; i < N * 64; i += blockDim.x) {
        int n = i / 64, k = i % 64;
        B[sw128_index(n, k)] = __float2bfloat16((float)(((n * 2 + k) % 5) - 2));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    if (threadIdx.x == 0) {
        const uint32_t idesc = idesc_bf16(N);
        const uint64_t da = desc_sw128(smem_u32(A)), db = desc_sw128(smem_u32(B));
        for (int k = 0; k < 4; ++k) {
            if (mode == 0) {
                mma_ss(tmem, da + 2 * k, db + 2 * k, idesc, k ? 1u : 0u);
            } else {
                cp_128x256b(tmem + A_COL + 8 * k, da + 2 * k);
                mma_ts(tmem, tmem + A_COL + 8 * k, db + 2 * k, idesc, k ? 1u : 0u);
            }
        }
        commit(&bar);
    }
    wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                       "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) out[(size_t)threadIdx.x * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

// Unaligned / strided A view check: A lives in a 176-row SWIZZLE_128B region (rows 128 B apart, XOR phase = absolute row & 7, the
// way TMA writes a box).  The MMA tile row i is region row (i/8)*gstride + (i%8) + r0: an 8-row group every `gstride` rows
// (descriptor SBO = gstride*128 B) starting r0 rows into the region (descriptor start = base + r0*128 B).  Tells whether the
// hardware derives the swizzle phase from absolute smem address bits (then any r0 / gstride works: a full (8+2)x(16+2) halo
// patch can serve all nine taps) or from the row index inside the group.  bo_mode: descriptor base_offset field 0 or (start>>7)&7.
__global__ void __launch_bounds__(128, 1) view_check_kernel(int r0, int gstride, int bo_mode, float* out) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int N = 64, ROWS = 176;
    __nv_bfloat16* A = (__nv_bfloat16*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    __nv_bfloat16* B = A + ROWS * 64;
    for (int i = threadIdx.x; i < ROWS * 64; i += blockDim.x) {
        int r = i / 64, k = i % 64;
        A[sw128_index(r, k)] = __float2bfloat16((float)(((r * 3 + k * 5) % 7) - 3));
    }
    for (int i = threadIdx.x; i < N * 64; i += blockDim.x) {
        int n = i / 64, k = i % 64;
        B[sw128_index(n, k)] = __float2bfloat16((float)(((n * 2 + k) % 5) - 2));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    if (threadIdx.x == 0) {
        const uint32_t idesc = idesc_bf16(N);
        const uint32_t a_start = smem_u32(A) + (uint32_t)r0 * 128u;
        uint64_t da = 0;
        da |= (uint64_t)((a_start & 0x3FFFF) >> 4);
        da |= (uint64_t)1 << 16;
        da |= (uint64_t)((gstride * 128) >> 4) << 32;
        da |= (uint64_t)1 << 46;
        if (bo_mode) da |= (uint64_t)((a_start >> 7) & 7) << 49;
        da |= (uint64_t)2 << 61;
        const uint64_t db = desc_sw128(smem_u32(B));
        for (int k = 0; k < 4; ++k) mma_ss(tmem, da + 2 * k, db + 2 * k, idesc, k ? 1u : 0u);
        commit(&bar);
    }
    wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                       "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) out[(size_t)threadIdx.x * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
    }
}

// pattern 0: SS x3 (3 MMAs per k16); 1: TS x3 (cp hi, 2 MMAs, cp lo, 1 MMA); 2: TS MMAs only (A resident, no cp);
// 3: cp only (2 per k16)
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int reps, int pattern, long long* out) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ uint64_t bar, bar2, bar3;
    __shared__ uint32_t s_tmem;
    unsigned char* tiles = (unsigned char*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < 4 * 96 * 1024 / 4 / 2; i += blockDim.x) ((uint32_t*)tiles)[i] = 0;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar2)), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar3)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    // patterns 9-12: SS x3 while the other three warps poll an mbarrier that completes only at the end, the way idle pipeline
    // roles do in the conv kernel: 9 every lane spins on try_wait, 10 one lane per warp spins, 11 every lane, with a
    // 20 us suspend-time hint, 12 one lane per warp with nanosleep(200) back-off
    __shared__ uint64_t bar4;
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar4)), "r"(1));
    __syncthreads();
    if (pattern >= 9 && warp > 0) {
        const bool poll = (pattern == 9 || pattern == 11) || (threadIdx.x & 31) == 0;
        if (poll) {
            uint32_t ok = 0;
            while (!ok) {
                if (pattern == 11)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(ok) : "r"(smem_u32(&bar4)), "r"(0), "r"(20000) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(ok) : "r"(smem_u32(&bar4)), "r"(0) : "memory");
                if (pattern == 12 && !ok) __nanosleep(200);
            }
        }
        __syncwarp();
    }
    const int pat_in = pattern;
    if (pattern >= 9) pattern = 0;
    if (threadIdx.x == 0) {
        const uint32_t idesc = idesc_bf16(N);
        const uint32_t base = smem_u32(tiles);
        const long long t0 = clock64();
        if (pattern >= 4) {
            // SS x3 with the conv kernel's per-stage extras: 4 commit (no wait) per 12 MMAs, 5 try_wait on a ready barrier +
            // tcgen05.fence per 12 MMAs, 6 both, 7 both with TWO commits per stage, 8 like 6 but alternating accumulators
            for (int r = 0; r < reps; ++r) {
                const uint32_t st = base + (r & 1) * 96 * 1024;
                if (pattern == 5 || pattern >= 6) {
                    wait(&bar3, 1);                        // fresh barrier: the "previous phase" is complete -> returns at once
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                const uint64_t ah = desc_sw128(st), al = desc_sw128(st + 16 * 1024), bh = desc_sw128(st + 32 * 1024),
                               bl = desc_sw128(st + 64 * 1024);
                const uint32_t d = tmem + ((pattern == 8 && ((r >> 4) & 1)) ? 256u : 0u);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    mma_ss(d, al + 2 * k, bh + 2 * k, idesc, (r | k) ? 1u : 0u);
                    mma_ss(d, ah + 2 * k, bl + 2 * k, idesc, 1u);
                    mma_ss(d, ah + 2 * k, bh + 2 * k, idesc, 1u);
                }
                if (pattern == 4 || pattern >= 6) commit(&bar2);
                if (pattern == 7) commit(&bar2);
            }
            commit(&bar);
            wait(&bar, 0);
            out[blockIdx.x] = clock64() - t0;
        } else {
        for (int r = 0; r < reps; ++r) {
            // stage = A_hi 16K, A_lo 16K, B_hi 32K, B_lo 32K; two stages rotate
            const uint32_t st = base + (r & 1) * 96 * 1024;
            const uint64_t ah = desc_sw128(st), al = desc_sw128(st + 16 * 1024), bh = desc_sw128(st + 32 * 1024),
                           bl = desc_sw128(st + 64 * 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t slot = tmem + A_COL + 16 * ((r * 4 + k) & 7);      // 8 rotating (hi, lo) slots of 8+8 columns
                if (pattern == 0) {
                    mma_ss(tmem, al + 2 * k, bh + 2 * k, idesc, (r | k) ? 1u : 0u);
                    mma_ss(tmem, ah + 2 * k, bl + 2 * k, idesc, 1u);
                    mma_ss(tmem, ah + 2 * k, bh + 2 * k, idesc, 1u);
                } else if (pattern == 1) {
                    cp_128x256b(slot, ah + 2 * k);
                    cp_128x256b(slot + 8, al + 2 * k);
                    mma_ts(tmem, slot + 8, bh + 2 * k, idesc, (r | k) ? 1u : 0u);
                    mma_ts(tmem, slot, bl + 2 * k, idesc, 1u);
                    mma_ts(tmem, slot, bh + 2 * k, idesc, 1u);
                } else if (pattern == 2) {
                    mma_ts(tmem, slot + 8, bh + 2 * k, idesc, (r | k) ? 1u : 0u);
                    mma_ts(tmem, slot, bl + 2 * k, idesc, 1u);
                    mma_ts(tmem, slot, bh + 2 * k, idesc, 1u);
                } else {
                    cp_128x256b(slot, ah + 2 * k);
                    cp_128x256b(slot + 8, al + 2 * k);
                }
            }
        }
        commit(&bar);
        wait(&bar, 0);
        out[blockIdx.x] = clock64() - t0;
        }
        if (pat_in >= 9) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar4)) : "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

int main() {
    const size_t smem = 200 * 1024;
    cudaFuncSetAttribute(check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float* d_out;
    cudaMalloc(&d_out, 128 * 256 * sizeof(float));
    float* h = (float*)malloc(128 * 256 * sizeof(float));
    for (int N : {64, 160, 256})
        for (int mode : {0, 1}) {
            cudaMemset(d_out, 0xff, 128 * 256 * sizeof(float));
            check_kernel<<<1, 128, smem>>>(N, mode, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("check N %d mode %d: error %s\n", N, mode, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h, d_out, 128 * N * sizeof(float), cudaMemcpyDeviceToHost);
            int bad = 0; double maxd = 0;
            for (int r = 0; r < 128; ++r)
                for (int n = 0; n < N; ++n) {
                    float ref = 0;
                    for (int k = 0; k < 64; ++k) ref += (float)(((r * 3 + k * 5) % 7) - 3) * (float)(((n * 2 + k) % 5) - 2);
                    double d = fabs((double)h[r * N + n] - ref);
                    if (!(d == 0)) { if (bad < 4) printf("   mismatch r %d n %d got %g want %g\n", r, n, h[r * N + n], ref); ++bad; }
                    if (d > maxd) maxd = d;
                }
            printf("check N %3d %s: %s (%d mismatches of %d)\n", N, mode ? "TS (tcgen05.cp.128x256b + A in TMEM)" : "SS", bad ? "WRONG" : "exact",
                   bad, 128 * N);
        }
    cudaFuncSetAttribute(view_check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int gstride : {8, 10})
        for (int r0 : {0, 1, 2, 5, 10, 11, 21})
            for (int bo : {0, 1}) {
                if (gstride == 8 && r0 > 5) continue;
                cudaMemset(d_out, 0xff, 128 * 256 * sizeof(float));
                view_check_kernel<<<1, 128, smem>>>(r0, gstride, bo, d_out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("view check: error %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h, d_out, 128 * 64 * sizeof(float), cudaMemcpyDeviceToHost);
                int bad = 0;
                for (int i = 0; i < 128; ++i) {
                    const int r = (i / 8) * gstride + (i % 8) + r0;
                    for (int n = 0; n < 64; ++n) {
                        float ref = 0;
                        for (int k = 0; k < 64; ++k) ref += (float)(((r * 3 + k * 5) % 7) - 3) * (float)(((n * 2 + k) % 5) - 2);
                        if (h[i * 64 + n] != ref) ++bad;
                    }
                }
                printf("view check: 8-row groups every %2d rows, start row %2d, base_offset %s: %s (%d of %d wrong)\n", gstride, r0,
                       bo ? "(start>>7)&7" : "0", bad ? "WRONG" : "exact", bad, 128 * 64);
            }
    long long* d;
    cudaMalloc(&d, 148 * sizeof(long long));
    const int reps = 1000;
    const char* names[] = {"SS x3 (3 MMA)", "TS x3 (2 cp + 3 MMA)", "TS x3 MMAs only", "cp only (2 per k16)",
                           "SS x3 + commit/12", "SS x3 + try_wait/12", "SS x3 + both", "SS x3 + wait + 2 commits", "SS x3 + both, 2 accumulators",
                           "SS x3, 96 lanes polling", "SS x3, 3 lanes polling", "SS x3, 96 lanes, 20us hint", "SS x3, 3 lanes + nanosleep"};
    for (int grid : {1, 148})
        for (int N : {64, 128, 192, 256})
            for (int pattern : {0, 1, 2, 3, 6, 9}) {
                rate_kernel<<<grid, 128, smem>>>(N, reps, pattern, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("rate: error %s\n", cudaGetErrorString(e)); return 1; }
                long long hh[148];
                cudaMemcpy(hh, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
                double tot = 0;
                for (int i = 0; i < grid; ++i) tot += hh[i];
                printf("grid %3d N %3d %-30s: %7.1f cyc per k16 step (nominal x3 = %d)\n", grid, N, names[pattern],
                       tot / grid / (reps * 4.0), 3 * N / 2);
            }
    return 0;
}
