// Micro-benchmark: raw issue/execute rate of tcgen05.mma (cta_group::1, kind::f16, SS operands) on one SM and on all SMs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate tools/mma_rate.cu && ./mma_rate
// Reports cycles per M128 x N x K16 MMA for several N and operand-rotation patterns (nominal: N/2 cycles).
#include <cstdio>
#include <cstdint>
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
__device__ __forceinline__ void mma(uint32_t tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// rot: number of distinct (A,B) smem tile pairs cycled through (1 = same operands every MMA)
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int reps, int rot, int commit_every, long long* out) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ uint64_t bar;
    __shared__ uint32_t s_tmem;
    unsigned char* tiles = (unsigned char*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < 8 * 48 * 1024 / 4 && i < 49152; i += blockDim.x) ((uint32_t*)tiles)[i] = 0;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    if (threadIdx.x == 0) {
        const uint32_t idesc = idesc_bf16(N);
        const uint32_t base = smem_u32(tiles);
        const long long t0 = clock64();
        uint32_t parity = 0;
        for (int r = 0; r < reps; ++r) {
            const int t = r % rot;
            const uint32_t a = base + t * (16 * 1024 + 32 * 1024), b = a + 16 * 1024;
#pragma unroll
            for (int k = 0; k < 4; ++k) mma(tmem, desc_sw128(a) + 2 * k, desc_sw128(b) + 2 * k, idesc, (r | k) ? 1u : 0u);
            if (commit_every && (r % commit_every) == commit_every - 1) {
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
                uint32_t ok = 0;
                while (!ok)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(ok) : "r"(smem_u32(&bar)), "r"(parity) : "memory");
                parity ^= 1;
            }
        }
        const long long t_issue = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(parity) : "memory");
        const long long t1 = clock64();
        out[blockIdx.x * 2] = t1 - t0;
        out[blockIdx.x * 2 + 1] = t_issue - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
    }
}

int main() {
    long long* d;
    cudaMalloc(&d, 148 * 2 * sizeof(long long));
    const size_t smem = 200 * 1024;
    cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int reps = 2000;
    printf("cycles per MMA (M128 x N x K16, bf16, SS, cta_group::1); nominal = N/2\n");
    for (int grid : {1, 148})
        for (int N : {64, 128, 192, 256})
            for (int rot : {1, 4})
                for (int ce : {0, 1, 3}) {
                    mma_rate_kernel<<<grid, 128, smem>>>(N, reps, rot, ce, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    long long h[296];
                    cudaMemcpy(h, d, grid * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
                    double tot = 0, iss = 0;
                    for (int i = 0; i < grid; ++i) { tot += h[2 * i]; iss += h[2 * i + 1]; }
                    printf("grid %3d N %3d rot %d commit/wait every %d stage(s): %.1f cyc/MMA total, %.1f cyc/MMA issue-side (nominal %d)\n",
                           grid, N, rot, ce, tot / grid / (reps * 4.0), iss / grid / (reps * 4.0), N / 2);
                }
    return 0;
}
