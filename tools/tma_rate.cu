// Micro-benchmark: per-SM TMA ingest rate (global/L2 -> shared memory) for the box shapes of the conv kernel, all SMs loading at once.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_rate tools/tma_rate.cu -lcuda && tools/tma_rate
// The r2 per-CTA timelines show the single-tile 3x3 layers pacing at ~30 bytes/clk/SM of operand loads (13.2 us for 801 KB), most
// of it weight (B) tiles fetched as 96 separate 128-byte rows (tap-major weights: a row every taps * Cin * 2 bytes).  Patterns:
//   0  B tile as today:      3-D map [Cout][taps][Cin], box {64, 1, rows}   - rows 128 B each, far apart in global memory
//   1  B tile pre-tiled:     2-D map [tiles * rows][64],  box {64, rows}    - the tile's rows contiguous (rows * 128 B)
//   2  B tile as bulk copy:  cp.async.bulk of rows * 128 contiguous bytes   - no tensor map (weights pre-swizzled on the host)
//   3  A halo patch:         4-D map NHWC [7][H][W][C], box {64, 10, 18, 1}
//   4  A generic tile:       4-D map NHWC, box {64, 16, 8, 1}
// Each CTA: one producer thread keeps `depth` loads in flight into a ring of smem slots; bytes / elapsed cycles is reported.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
        if (!ok && clock64() - t0 > 2000000000LL) __trap();
    }
}
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int SLOT = 24 * 1024;      // bytes per ring slot (>= the largest box: 23 KB halo patch)
constexpr int MAXD = 8;

struct Params {
    int pattern, iters, depth, rows, producers, batch, lanes;   // rows: B tile rows (N / 2 of a CTA pair)
    int Cout, taps, Cin, H, W, C, N;
    const unsigned char* wtiled;             // pre-tiled weights (pattern 2)
    long long* out;
};

// P producer threads (lane 0 of warps 0..P-1), each with its own `depth` slots and barriers
__global__ void __launch_bounds__(256, 1)
tma_rate_kernel(const __grid_constant__ CUtensorMap m0, const __grid_constant__ CUtensorMap m1, const __grid_constant__ CUtensorMap mA,
                const __grid_constant__ CUtensorMap mG, const Params p) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ uint64_t full_all[MAXD];
    unsigned char* ring_all = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    if (threadIdx.x == 0) {
        for (int i = 0; i < MAXD; ++i) mbar_init(&full_all[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // lanes = 1: the producers are lanes 0..P-1 of warp 0 (do loads of different THREADS of one warp overlap?)
    const int prod = p.lanes ? (int)threadIdx.x : (int)(threadIdx.x >> 5);
    if (p.lanes ? threadIdx.x >= p.producers : ((threadIdx.x & 31) != 0 || prod >= p.producers)) return;
    uint64_t* full = full_all + prod * p.depth;                  // producers * depth <= MAXD
    unsigned char* ring = ring_all + (size_t)prod * p.depth * SLOT;
    const uint32_t bbytes = (uint32_t)p.rows * 128u;
    const uint32_t bytes = p.pattern <= 2 ? bbytes : p.pattern == 3 ? 10u * 18u * 128u : 16u * 8u * 128u;
    const int kchunks = p.Cin / 64, ntile = p.Cout / p.rows;
    const int wt = p.W / 16, ht = p.H / 8;
    auto issue = [&](int it, int slot, bool expect = true) {
        if (expect) mbar_expect(&full[slot], bytes);
        unsigned char* dst = ring + (size_t)slot * SLOT;
        const int u = it * p.producers + prod + (int)blockIdx.x * 7;     // CTAs / producers walk the data at different phases
        if (p.pattern == 0) {
            tma_3d(dst, &m0, &full[slot], (u % kchunks) * 64, (u / kchunks) % p.taps, ((u / (kchunks * p.taps)) % ntile) * p.rows);
        } else if (p.pattern == 1) {
            tma_2d(dst, &m1, &full[slot], 0, (u % (kchunks * p.taps * ntile)) * p.rows);
        } else if (p.pattern == 2) {
            bulk_1d(dst, p.wtiled + (size_t)(u % (kchunks * p.taps * ntile)) * bbytes, bbytes, &full[slot]);
        } else if (p.pattern == 3) {
            tma_4d(dst, &mA, &full[slot], ((u / 3) % (p.C / 64)) * 64, ((u % wt) * 16) - 1, (((u / wt) % ht) * 8) - 1, u % p.N);
        } else {
            tma_4d(dst, &mG, &full[slot], ((u / 3) % (p.C / 64)) * 64, (u % wt) * 16, ((u / wt) % ht) * 8, u % p.N);
        }
    };
    const long long t0 = clock64();
    if (p.batch > 1) {
        // batches of `batch` loads: all the barrier work first, then the loads back to back (depth = 2 * batch slots)
        const int nb = p.iters / p.batch;
        auto issue_batch = [&](int b) {
            const int s0 = (b & 1) * p.batch;
            for (int j = 0; j < p.batch; ++j) mbar_expect(&full[s0 + j], bytes);
            for (int j = 0; j < p.batch; ++j) issue(b * p.batch + j, s0 + j, false);
        };
        issue_batch(0);
        if (nb > 1) issue_batch(1);
        for (int b = 0; b < nb; ++b) {
            const int s0 = (b & 1) * p.batch;
            for (int j = 0; j < p.batch; ++j) mbar_wait(&full[s0 + j], (uint32_t)(b >> 1) & 1u);
            if (b + 2 < nb) issue_batch(b + 2);
        }
        const long long t1 = clock64();
        p.out[blockIdx.x * 2] = t1 - t0;
        p.out[blockIdx.x * 2 + 1] = (long long)bytes * nb * p.batch;
        return;
    }
    for (int i = 0; i < p.depth && i < p.iters; ++i) issue(i, i);
    for (int it = 0; it < p.iters; ++it) {
        const int slot = it % p.depth;
        mbar_wait(&full[slot], (uint32_t)(it / p.depth) & 1u);
        if (it + p.depth < p.iters) issue(it + p.depth, slot);
    }
    const long long t1 = clock64();
    if (prod == 0) {
        p.out[blockIdx.x * 2] = t1 - t0;
        p.out[blockIdx.x * 2 + 1] = (long long)bytes * p.iters * p.producers;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode(EncodeTiledFn fn, CUtensorMap* tm, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                  const cuuint32_t* box) {
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d (rank %d)\n", (int)r, rank); return 1; }
    return 0;
}

int main() {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
        printf("no cuTensorMapEncodeTiled\n");
        return 1;
    }
    EncodeTiledFn fn = (EncodeTiledFn)f;
    const int Cout = 192, taps = 9, Cin = 192, N = 7, H = 40, W = 60, C = 768;
    unsigned char *w, *wt, *x;
    const size_t wbytes = (size_t)Cout * taps * Cin * 2, xbytes = (size_t)N * H * W * C * 2;
    cudaMalloc(&w, wbytes); cudaMalloc(&wt, wbytes); cudaMalloc(&x, xbytes);
    cudaMemset(w, 0, wbytes); cudaMemset(wt, 0, wbytes); cudaMemset(x, 0, xbytes);
    long long* out;
    cudaMalloc(&out, 148 * 2 * sizeof(long long));
    cudaFuncSetAttribute(tma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MAXD * SLOT + 1024);
    for (int rows : {96}) {
        CUtensorMap m0, m1, mA, mG;
        {
            cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)taps, (cuuint64_t)Cout};
            cuuint64_t str[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)taps * Cin * 2};
            cuuint32_t box[3] = {64, 1, (cuuint32_t)rows};
            if (encode(fn, &m0, w, 3, dims, str, box)) return 1;
        }
        {
            cuuint64_t dims[2] = {64, (cuuint64_t)(wbytes / 128)};
            cuuint64_t str[1] = {128};
            cuuint32_t box[2] = {64, (cuuint32_t)rows};
            if (encode(fn, &m1, wt, 2, dims, str, box)) return 1;
        }
        {
            cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
            cuuint64_t str[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
            cuuint32_t boxA[4] = {64, 18, 10, 1}, boxG[4] = {64, 16, 8, 1};
            if (encode(fn, &mA, x, 4, dims, str, boxA)) return 1;
            if (encode(fn, &mG, x, 4, dims, str, boxG)) return 1;
        }
        for (int pattern = 0; pattern < 5; ++pattern) {
            if (pattern >= 3 && rows != 96) continue;
            for (int cfg = 0; cfg < 9; ++cfg)
                for (int grid : {148}) {
                    // {producer threads, slots each, batch, producers are lanes of one warp}
                    static const int PD[9][4] = {{1, 1, 1, 0}, {1, 8, 1, 0}, {2, 4, 1, 0}, {4, 2, 1, 0}, {1, 4, 2, 0}, {1, 8, 4, 0},
                                                 {2, 4, 1, 1}, {4, 2, 1, 1}, {8, 1, 1, 1}};
                    const int depth = PD[cfg][1];
                    Params p = {};
                    p.pattern = pattern; p.iters = 2000; p.depth = depth; p.rows = rows; p.producers = PD[cfg][0];
                    p.batch = PD[cfg][2]; p.lanes = PD[cfg][3];
                    p.Cout = Cout; p.taps = taps; p.Cin = Cin; p.H = H; p.W = W; p.C = C; p.N = N;
                    p.wtiled = wt; p.out = out;
                    tma_rate_kernel<<<grid, 256, MAXD * SLOT + 1024>>>(m0, m1, mA, mG, p);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s (pattern %d)\n", cudaGetErrorString(e), pattern); return 1; }
                    std::vector<long long> h(grid * 2);
                    cudaMemcpy(h.data(), out, grid * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
                    double worst = 0, bytes = 0;
                    for (int i = 0; i < grid; ++i) { if ((double)h[2 * i] > worst) worst = (double)h[2 * i]; bytes = (double)h[2 * i + 1]; }
                    static const char* names[] = {"B rows strided (3-D map, today)", "B tile contiguous (2-D map)", "B tile cp.async.bulk 1-D",
                                                  "A halo patch {64,18,10}", "A generic tile {64,16,8}"};
                    const double per_load = (pattern <= 2 ? rows * 128 : pattern == 3 ? 23040 : 16384);
                    printf("%-34s %5.1f KB/load  %d producer %s x %d slot(s) batch %d  grid %3d : %6.1f bytes/clk/SM = one load per %6.0f clk  (%.2f TB/s)\n",
                           names[pattern], per_load / 1024.0, p.producers, p.lanes ? "lanes of one warp" : "warps", depth, p.batch, grid,
                           bytes / worst, per_load / (bytes / worst),
                           bytes / worst * grid * 1.965e9 / 1e12);
                }
        }
    }
    return 0;
}
