// Implicit-GEMM convolution / GEMM on Blackwell tensor cores (tcgen05.mma, TMEM accumulators, TMA-fed smem).
//
// Replaces the cuDNN conv + BatchNorm(eval) + ReLU triples of models/backbones/vovnet.py:124-161 (80 conv3x3,
// 16 concat conv1x1, stem convs 2-3), the mmdet FPN lateral / output / extra convs (config far3d.py:50-57), the 2D
// head's towers (yolox_head.py:164-231) and - with ksize 1 on a [rows, K] "image" - nn.Linear layers.
//
// GEMM view: M = output pixels, N = Cout, K = taps * Cin.  A persistent worker is a CTA pair (cluster of 2,
// cta_group::2: one M=256 MMA stream over two adjacent 128-pixel tiles, half a B tile staged per CTA) or, for one-tile
// problems, a single CTA.  Per CTA:
//   warp 0: TMA producer, warp 1: tcgen05.mma issuer (leader CTA only; whole warp in lockstep, one elected lane issues;
//   fp32 accumulator in TMEM), warps 2-9: epilogue (two per TMEM lane quadrant; tcgen05.ld -> bias + activation -> fp32 / fp16 / split-fp16 stores at
//   a channel offset, so OSA concat buffers are written in place and torch.cat of vovnet.py:230 disappears).
//
// What the r1 measurements say limits these kernels, in the order it was found (profiles/r1b_*):
//   * the single MMA-issuing thread (fixed: elect_one_sync, warp-uniform loop), then the epilogue's code generation
//     (fixed: branch-free vector math, shared-window pointers), then shared-memory bandwidth (MMA operand reads + TMA
//     writes against 128 B/clk/SM: CTA pairs halve the B side);
//   * halo mode (3x3, stride 1 - 80 % of the FLOPs): per 64-channel chunk and per filter offset along the
//     tile's 8-pixel side, ONE TMA box brings an 8 x (16+2) pixel column patch into smem (144 SWIZZLE_128B rows); the
//     three taps along the 16-pixel side are three UMMA descriptors into that patch (start advanced by whole 8-row
//     groups = 1024 B, i.e. swizzle-atom aligned), so activations cross L2->SM 3x instead of 9x.  TMA zero-fill outside
//     the image is the conv padding.
//   * generic mode (1x1, stride-2 3x3): A tile per (tap, chunk) as a shifted tiled-TMA box {64, tw, th, 1}
//     (stride 2 views the tensor as {2C, W/2, 2, H/2, N} so a tap is again a dense box).
//   * persistent workers with two TMEM accumulator stages: the epilogue of tile i overlaps the MMAs of tile i+1 and the
//     producer prefetches across tile boundaries.  TMA multicast was measured and dropped (it cuts L2 reads, not per-SM
//     ingest).
//
// "fp16x3" (split) mode: activations and weights are stored as fp16 hi + fp16 lo planes (value = hi + lo); each k-step
// issues lo*hi + hi*lo + hi*hi into the same fp32 accumulator, which reproduces fp32 convolution to ~2^-17 relative
// (the reference computes in fp32/TF32) while staying on the fp16 tensor pipe.
//
// "fp16mx" mode (round 2; MODE == 2): the two correction terms need only ~4 significant bits of each operand, so they run as
// ONE stream of e4m3 x e4m3 MMAs (K = 32 per instruction, same issue time as a K = 16 fp16 MMA - tools/mma_mx.cu,
// profiles/r2a_mma_mx.txt) into the SAME fp32 TMEM accumulator as the fp16 main term: the "lo" planes of activations and
// weights are replaced by e4m3 correction planes of the same size (common.cuh), 32 channels cost 2 fp16 + 2 e4m3 MMAs instead
// of 6 fp16 MMAs: 2 tensor-pipe passes per MAC instead of 3.  Round 3: the e4m3 MMAs are plain kind::f8f6f4 - both correction
// products carry the same power of two, 2^(11 + EA + w_exp), and the (static) fp16 weight plane is stored pre-multiplied by
// it, so the epilogue removes one common factor.  (Round 2 used kind::mxf8f6f4.block_scale with uniform UE8M0 scale factors:
// their 16 TMEM columns capped the N tile at 224, i.e. Cout = 256 / 512 / 1024 ran as 128-wide tiles at the shared-memory
// bandwidth bound; without them the accumulator stages are 256 columns wide.)
#include <cuda.h>
#include <algorithm>
#include "common.cuh"

namespace far3d {

typedef __half fp16;

struct ConvParams {
    int N, H, W, Ho, Wo;          // input / output spatial dims
    int Cin, Cout, ks, stride;
    int tw, th, tiles_w, tiles_h; // generic kernel: M tile = th x tw output pixels (tw*th == 128)
    int tiles_f, tiles_s;         // halo kernel: tiles along the fast (8 px) / slow (16 px) image dimension
    int transposed;               // halo kernel: fast dimension is H (1) or W (0)
    int m_tiles;                  // real tile count (grid.x may be padded to the cluster size)
    int cm;                       // cluster size along M (B multicast)
    int bn;                       // N tile
    int kchunks;                  // ceil(Cin / 64)
    int x_cs, x_co;               // used for stride-2 channel coordinate
    int num_stages;               // generic kernel ring depth / halo kernel B ring depth
    int a_stages;                 // halo kernel A ring depth
    int b_resident;               // halo kernel: the B ring holds every (chunk, tap) weight tile of the layer (num_stages == 9 * kchunks):
                                  // loaded during the worker's first tile, never released (stem conv 2: 64 -> 64)
    int relu;
    float acc_scale;              // 1 + (expected truncation loss of the TMEM accumulation), applied to the accumulator in the
                                  // epilogue before the bias (see kRzLossPerMma)
    const float* bias;
    float* y_f32; int yf_cs, yf_co; long long yf_ns;
    fp16* y_hi; fp16* y_lo; int yb_cs, yb_co;
    const float* res; int res_cs; // optional fp32 residual added after the activation, indexed like y_f32 (dense rows)
    float* colsum;                // optional [m_tiles * 4][Cout] per-(tile, epilogue warp) column sums of the fp32 output
                                  // (global average pool partials of the eSE block, fused into the concat conv)
    int y_fmt;                    // 0: y_lo is the fp16 residual plane; else FAR3D_LO_MX(EA): y_lo is an e4m3 correction plane
    long long* dbg;               // optional per-CTA timestamps (ns): 0 start, 1 first tile's MMAs issued, 2 all MMAs issued, 3 epilogue
                                  // done, 4 end, 5 loads issued, 6 first accumulator ready
};

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded spin: a pipeline bug traps (launch error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();     // ~2 s at 1.9 GHz
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ long long gtime() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// debug buffer: 16 int64 per CTA.  0-7: %globaltimer stamps; 8-15: cycles spent waiting per role (only in builds with
// -DFAR3D_CONV_WAITSTATS, tools/conv_timeline.py --waits): 8 MMA on acc_empty, 9 MMA on a_full, 10 MMA on b_full, 11 producer 0 on
// b_empty, 12 producer 0 on a_empty, 13 epilogue warp 2 on acc_full, 14 epilogue warp 2 busy, 15 tiles of this worker
#define DBG_STAMP(i) do { if (p.dbg && blockIdx.y == 0) p.dbg[(size_t)blockIdx.x * 16 + (i)] = gtime(); } while (0)
#ifdef FAR3D_CONV_WAITSTATS
#define WAIT_ACC(acc, bar, par) do { const long long _t = clock64(); mbar_wait(bar, par); acc += clock64() - _t; } while (0)
#define DBG_PUT(i, v) do { if (p.dbg && blockIdx.y == 0) p.dbg[(size_t)blockIdx.x * 16 + (i)] = (long long)(v); } while (0)
#else
#define WAIT_ACC(acc, bar, par) mbar_wait(bar, par)
#define DBG_PUT(i, v) do { } while (0)
#endif
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

// address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// "accumulator stage drained": what must be ordered before the arrive are this warp's TMEM reads, and tcgen05.wait::ld +
// tcgen05.fence::before_thread_sync do that; a .release arrive additionally waits for the warp's outstanding GLOBAL stores
// (the tile's output!) to be performed - ERRBAR + a stalled SYNCS.ARRIVE, 7 % of all stall samples in the r2 ncu source page
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_relaxed(uint64_t* bar) {
    asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// cta_group::2 TMA loads: the data lands in THIS CTA's smem, the transaction bytes are signalled on an mbarrier that may
// live in the peer CTA (`bar` is a shared::cluster address - the leader's "full" barrier of the stage)
__device__ __forceinline__ void tma2_load_3d(void* dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma2_load_4d(void* dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma2_load_5d(void* dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

// CG-generic forms: `bar` is a shared::cta address (CG == 1) or the shared::cluster address of the leader's barrier (CG == 2)
template <int CG>
__device__ __forceinline__ void tma_ld_3d(void* dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    if (CG == 2) tma2_load_3d(dst, map, bar, c0, c1, c2);
    else asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                      ::"r"(smem_u32(dst)), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
template <int CG>
__device__ __forceinline__ void tma_ld_4d(void* dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    if (CG == 2) tma2_load_4d(dst, map, bar, c0, c1, c2, c3);
    else asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                      ::"r"(smem_u32(dst)), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
template <int CG>
__device__ __forceinline__ void tma_ld_5d(void* dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    if (CG == 2) tma2_load_5d(dst, map, bar, c0, c1, c2, c3, c4);
    else asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                      ::"r"(smem_u32(dst)), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_3d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                               uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5, %6}], [%2], %3;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B; 8-row groups `sbo_bytes` apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t sbo_bytes = 1024) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major), bits [16,30)
    d |= (uint64_t)(sbo_bytes >> 4) << 32;           // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                          // descriptor version 1 (Blackwell), bits [46,48)
    // base_offset (bits [49,52)) stays 0: the hardware takes the swizzle phase from the absolute smem address bits, so a
    // descriptor may start at any 128 B row of a TMA-written region and step 8-row groups by any multiple of 128 B
    // (tools/mma_ts.cu "view check": exact for start rows 0..11 and group strides 8 / 10 rows; a non-zero field is wrong)
    d |= (uint64_t)2 << 61;                          // layout type SWIZZLE_128B, bits [61,64)
    return d;
}
// instruction descriptor for kind::f16: D=f32, A=B=fp16, both K-major, M=128 (256 for a CTA pair), N=bn (cute::UMMA::InstrDescriptor)
__device__ __forceinline__ uint32_t umma_idesc_fp16(int bn, int m = 128) {
    uint32_t d = 0;
    d |= 1u << 4;                    // c_format = F32
    // a_format (bits 7-9) = b_format (bits 10-12) = 0: F16 (1 would be BF16)
    d |= (uint32_t)(bn >> 3) << 17;  // n_dim
    d |= (uint32_t)(m >> 4) << 24;   // m_dim
    return d;
}
__device__ __forceinline__ void umma_fp16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// CTA-pair MMA: M = 256 (rows 0-127 from the leader's A tile and TMEM, 128-255 from the peer's), each CTA supplies half of
// the N rows of B from the same smem offsets; issued by the leader only
__device__ __forceinline__ void umma2_fp16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// kind::f8f6f4 (no block scale): A = B = e4m3 (format code 0 in the kind::f16 descriptor layout), K = 32 per instruction,
// fp32 accumulate - the same issue time as a K = 16 fp16 MMA
template <int CG>
__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    if (CG == 2)
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// arrive (when all MMAs issued so far retire) on the barrier at this smem offset in both CTAs of the pair
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive on the barrier at this smem offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// TMA-issuing warps.  1: the round-2 wait accounting (profiles/r2u_conv_waits_producers.txt) shows 1, 3 and 5 producer warps
// stall the MMA warp identically - the kernel is bound by shared-memory bandwidth (MMA operand reads + TMA writes), not by
// TMA issue - and 192 threads x 255 registers leave a quarter of the register file for the other frame's head kernels.
#ifndef FAR3D_UM_PRODUCERS
#define FAR3D_UM_PRODUCERS 1
#endif
constexpr int UM_PRODUCERS = FAR3D_UM_PRODUCERS;
// Epilogue warps.  A warp may only touch the TMEM lane quadrant (warp % 4), so 8 warps = two per quadrant, which take alternate
// 32-column rounds of the accumulator.  Round 3: the epilogue of one warp per scheduler is issue- and latency-bound (2.5 us per
// 64-column round: ~7 us un-overlapped on every single-tile launch, and the stem / 1x1 layers with short K are epilogue-bound
// outright); the first 8-warp attempt spilled (64-column rounds hold r[64] + v[64] + 64 packed words per thread,
// profiles/r2h_*) - 32-column rounds need half of that and fit 8 warps in the register footprint of the 4-warp kernel.
#ifndef FAR3D_UM_EPI_WARPS
#define FAR3D_UM_EPI_WARPS 8
#endif
constexpr int UM_EPI = FAR3D_UM_EPI_WARPS;                  // 4 (64-column rounds) or 8 (32-column rounds)
static_assert(UM_EPI == 4 || UM_EPI == 8, "4 or 8 epilogue warps");
constexpr int UM_THREADS = 64 + 32 * UM_EPI + 32 * (UM_PRODUCERS - 1);   // warp 0 TMA, warp 1 MMA (+TMEM alloc), warps 2..2+UM_EPI-1 epilogue, then extra TMA warps
constexpr int UM_BM = 128, UM_BK = 64;
constexpr int UM_A_BYTES = UM_BM * UM_BK * 2;   // 16 KB per plane
constexpr int HALO_F = 8, HALO_S = 16;          // halo tile: 8 pixels along the fast dim, 16 along the slow dim
constexpr int HALO_PF = HALO_F + 2, HALO_PS = HALO_S + 2;        // patch = (8+2) x (16+2) pixels: every tap of the 3x3 filter
constexpr int HALO_PATCH_TX = HALO_PF * HALO_PS * UM_BK * 2;     // bytes one TMA box delivers: 180 rows x 128 B
constexpr int HALO_PATCH_BYTES = (HALO_PATCH_TX + 1023) / 1024 * 1024;   // plane stride in smem (1024-aligned): 23 KB

__device__ __forceinline__ uint32_t tmem_cols_for(int bn) {
    uint32_t c = 32;
    while ((int)c < bn) c <<= 1;
    return c;
}

// One elected lane of a converged warp (cute::elect_one_sync).  Code guarded by it sits in warp-uniform control flow, so
// the compiler keeps descriptors in uniform registers and emits a bare UTCHMMA / UTMALDG; the same instructions inside an
// `if (lane == 0)` region get wrapped in a vote-and-branch waterfall each (r1: ~170 clk of issue overhead per MMA, which made
// the single MMA-issuing thread the bottleneck of every layer).
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, px;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// The MMAs of one 64-channel chunk (`ksteps` k-steps of 16 channels); descriptors point at the chunk's first 32 bytes.
// MODE 0: plain fp16.  MODE 1 (fp16x3): value = hi + lo for both operands, lo*hi + hi*lo + hi*hi into the same fp32 accumulator
// (the lo*lo term is below fp32 resolution); the two MMAs that share A_hi are adjacent: an SS-form MMA re-fetches A (4 KB) only
// when the A descriptor changes (tools/mma_ts.cu).  MODE 2 (fp16mx): per 32 channels two fp16 MMAs hi*hi and two e4m3 MMAs on
// the correction planes - block 0 (scale-factor byte 0) = a_lo8 * w_hi8, block 1 (byte 1) = a_hi8 * w_lo8.
template <int MODE, int CG>
__device__ __forceinline__ void mma_chunk(uint32_t tmem, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo, uint32_t idesc,
                                          uint32_t accum, int ksteps) {
    if (MODE == 2) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (2 * h < ksteps) {
                const uint32_t o = 4u * h;                                   // 64 bytes = 4 descriptor address units
                if (CG == 2) {
                    umma2_fp16(tmem, a_hi + o, b_hi + o, idesc, h ? 1u : accum);
                    umma2_fp16(tmem, a_hi + o + 2, b_hi + o + 2, idesc, 1u);
                } else {
                    umma_fp16(tmem, a_hi + o, b_hi + o, idesc, h ? 1u : accum);
                    umma_fp16(tmem, a_hi + o + 2, b_hi + o + 2, idesc, 1u);
                }
                umma_f8<CG>(tmem, a_lo + o, b_lo + o, idesc, 1u);            // lo8 * w_hi8   (the idesc of kind::f16 with format
                umma_f8<CG>(tmem, a_lo + o + 2, b_lo + o + 2, idesc, 1u);    // hi8 * w_lo8    code 0 reads as e4m3 x e4m3 here)
            }
        }
        return;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < ksteps) {
            const uint32_t o = 2u * k;                                       // 16 fp16 = 32 B = 2 address units
            const uint32_t acc = k ? 1u : accum;
            if (CG == 2) {
                if (MODE == 1) {
                    umma2_fp16(tmem, a_lo + o, b_hi + o, idesc, acc);
                    umma2_fp16(tmem, a_hi + o, b_lo + o, idesc, 1u);
                    umma2_fp16(tmem, a_hi + o, b_hi + o, idesc, 1u);
                } else {
                    umma2_fp16(tmem, a_hi + o, b_hi + o, idesc, acc);
                }
            } else {
                if (MODE == 1) {
                    umma_fp16(tmem, a_lo + o, b_hi + o, idesc, acc);
                    umma_fp16(tmem, a_hi + o, b_lo + o, idesc, 1u);
                    umma_fp16(tmem, a_hi + o, b_hi + o, idesc, 1u);
                } else {
                    umma_fp16(tmem, a_hi + o, b_hi + o, idesc, acc);
                }
            }
        }
    }
}

// TMEM -> registers -> bias + activation -> global (fp32 and/or fp16 hi [+ lo]); one thread per output pixel.
__device__ __forceinline__ void epilogue_store(const ConvParams& p, uint32_t tmem_base, int quad, int img, int oh, int ow,
                                               bool pix_ok, int n0, int sub = 0, int nsub = 1) {
    const size_t pix = ((size_t)img * p.Ho + oh) * p.Wo + ow;
    float* yf = p.y_f32 ? p.y_f32 + (size_t)img * p.yf_ns + ((size_t)oh * p.Wo + ow) * p.yf_cs + p.yf_co : nullptr;
    fp16* yh = p.y_hi ? p.y_hi + pix * p.yb_cs + p.yb_co : nullptr;
    fp16* yl = p.y_lo ? p.y_lo + pix * p.yb_cs + p.yb_co : nullptr;
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
    for (int c = sub * 16; c < p.bn; c += 16 * nsub) {     // the warps of a quadrant take alternate 16-column rounds
        uint32_t r[16];
        tmem_ld16(trow + (uint32_t)c, r);
        tmem_ld_wait();
        const int col0 = n0 + c;
        if (!pix_ok || col0 >= p.Cout) continue;
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float t = __uint_as_float(r[j]) * p.acc_scale;
            if (p.bias && col0 + j < p.Cout) t += __ldg(p.bias + col0 + j);
            if (p.relu == 1) t = fmaxf(t, 0.f);
            else if (p.relu == 2) t = t / (1.f + __expf(-t));      // Swish (YOLOX towers)
            if (p.res && col0 + j < p.Cout) t += __ldg(p.res + pix * p.res_cs + col0 + j);
            v[j] = t;
        }
        if (col0 + 16 <= p.Cout) {
            if (yf) {
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(yf + col0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
            if (yh) {
                uint32_t ph[8], pl[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    fp16 h0, l0, h1, l1;
                    split_fp16(v[2 * j], h0, l0);
                    split_fp16(v[2 * j + 1], h1, l1);
                    ph[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                    pl[j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
                }
                uint4* dh = reinterpret_cast<uint4*>(yh + col0);
                dh[0] = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                dh[1] = make_uint4(ph[4], ph[5], ph[6], ph[7]);
                if (yl) {
                    uint4* dl = reinterpret_cast<uint4*>(yl + col0);
                    dl[0] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                    dl[1] = make_uint4(pl[4], pl[5], pl[6], pl[7]);
                }
            }
        } else {
            for (int j = 0; j < 16 && col0 + j < p.Cout; ++j) {
                if (yf) yf[col0 + j] = v[j];
                if (yh) {
                    fp16 h, l;
                    split_fp16(v[j], h, l);
                    yh[col0 + j] = h;
                    if (yl) yl[col0 + j] = l;
                }
            }
        }
    }
}

// Coalesced epilogue: the warp's 32 pixel rows x 64 columns are staged through a 4 KB smem buffer (16-byte chunks
// XOR-swizzled by row to stay bank-conflict free) and written as full 128-byte segments, 4 pixels per store instruction,
// instead of 32 scattered 16-byte pieces.  Per-warp smem: wbuf 4 KB | bias slice 1 KB | row base addresses 3 x 32 x 8 B.
// 8-warp form (32-column rounds): 64-byte rows - wbuf 2 KB | bias slice 1 KB | row base addresses 768 B.
constexpr int EP_WBUF = UM_EPI == 8 ? 2048 : 4096, EP_BIAS = 1024, EP_ROWS = UM_EPI == 8 ? 768 : 1024;
constexpr int EP_WARP_BYTES = EP_WBUF + EP_BIAS + EP_ROWS;

__device__ __forceinline__ void stage_and_store(unsigned char* wbuf, int lane, const uint4* chunks,
                                                const unsigned long long* rowbase, unsigned col_bytes, unsigned okmask,
                                                int nchunks_valid) {
    // write this lane's row: chunk j at swizzled slot
#pragma unroll
    for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(wbuf + lane * 128 + ((j ^ (lane & 7)) << 4)) = chunks[j];
    __syncwarp();
    const int j = lane & 7;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int pr = it * 4 + (lane >> 3);
        const uint4 v = *reinterpret_cast<const uint4*>(wbuf + pr * 128 + ((j ^ (pr & 7)) << 4));
        if (((okmask >> pr) & 1u) && j < nchunks_valid)
            *reinterpret_cast<uint4*>(rowbase[pr] + col_bytes + (unsigned)(j << 4)) = v;
    }
    __syncwarp();
}

template <int ACT>
__device__ __forceinline__ float activate(float t) {
    if (ACT == 1) return fmaxf(t, 0.f);
    if (ACT == 2) return __fdividef(t, 1.f + __expf(-t));      // Swish (YOLOX towers)
    return t;
}

// bias + activation on one 64-column round; full rounds are straight-line vector code (the per-element predicated form
// compiled to a branch per element and was the slowest part of the r1 epilogue)
template <int ACT, int W = 64>
__device__ __forceinline__ void round_math(const uint32_t* r, float* v, const float* sb, int cvalid, float sc) {
    if (cvalid == W) {
#pragma unroll
        for (int j = 0; j < W; j += 4) {
            const float4 b = *reinterpret_cast<const float4*>(sb + j);
            v[j] = activate<ACT>(fmaf(__uint_as_float(r[j]), sc, b.x));
            v[j + 1] = activate<ACT>(fmaf(__uint_as_float(r[j + 1]), sc, b.y));
            v[j + 2] = activate<ACT>(fmaf(__uint_as_float(r[j + 2]), sc, b.z));
            v[j + 3] = activate<ACT>(fmaf(__uint_as_float(r[j + 3]), sc, b.w));
        }
    } else {
#pragma unroll
        for (int j = 0; j < W; j += 4) {
            const float4 b = *reinterpret_cast<const float4*>(sb + j);     // the slice is zero-padded to bn columns
            const bool ok = j < cvalid;                                    // cvalid is a multiple of 8
            v[j] = ok ? activate<ACT>(fmaf(__uint_as_float(r[j]), sc, b.x)) : 0.f;
            v[j + 1] = ok ? activate<ACT>(fmaf(__uint_as_float(r[j + 1]), sc, b.y)) : 0.f;
            v[j + 2] = ok ? activate<ACT>(fmaf(__uint_as_float(r[j + 2]), sc, b.z)) : 0.f;
            v[j + 3] = ok ? activate<ACT>(fmaf(__uint_as_float(r[j + 3]), sc, b.w)) : 0.f;
        }
    }
}

__device__ __forceinline__ void epilogue_store_coalesced(const ConvParams& p, uint32_t tacc, int quad, int lane, int img,
                                                         int oh, int ow, bool pix_ok, int n0, unsigned char* wsm,
                                                         int mt) {
    unsigned char* wbuf = wsm;
    const float* sbias = reinterpret_cast<const float*>(wsm + EP_WBUF);
    unsigned long long* rows = reinterpret_cast<unsigned long long*>(wsm + EP_WBUF + EP_BIAS);   // [3][32]: y_hi, y_lo, y_f32
    const size_t pix = ((size_t)img * p.Ho + oh) * p.Wo + ow;
    const unsigned okmask = __ballot_sync(0xffffffffu, pix_ok);
    rows[lane] = p.y_hi ? (unsigned long long)(p.y_hi + pix * p.yb_cs + p.yb_co) : 0ull;
    rows[32 + lane] = p.y_lo ? (unsigned long long)(p.y_lo + pix * p.yb_cs + p.yb_co) : 0ull;
    rows[64 + lane] = p.y_f32 ? (unsigned long long)(p.y_f32 + (size_t)img * p.yf_ns + ((size_t)oh * p.Wo + ow) * p.yf_cs + p.yf_co) : 0ull;
    __syncwarp();
    const uint32_t trow = tacc + ((uint32_t)(quad * 32) << 16);
    const int act = p.relu;
    for (int c = 0; c < p.bn; c += 64) {
        const int ncols = min(64, p.bn - c);                 // multiple of 16, warp-uniform
        uint32_t r[64];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (q * 16 < ncols) tmem_ld16(trow + (uint32_t)(c + q * 16), r + q * 16);
        tmem_ld_wait();
        const int col0 = n0 + c;
        if (col0 >= p.Cout) continue;                        // warp-uniform
        const int cvalid = min(ncols, p.Cout - col0);        // valid columns in this round (multiple of 8)
        float v[64];
        if (act == 1) round_math<1>(r, v, sbias + c, cvalid, p.acc_scale);
        else if (act == 2) round_math<2>(r, v, sbias + c, cvalid, p.acc_scale);
        else round_math<0>(r, v, sbias + c, cvalid, p.acc_scale);
        if (p.res && pix_ok) {
            const float* rr = p.res + pix * p.res_cs + col0;
#pragma unroll
            for (int j = 0; j < 64; j += 4)
                if (j < cvalid) {
                    const float4 q = __ldg(reinterpret_cast<const float4*>(rr + j));
                    v[j] += q.x; v[j + 1] += q.y; v[j + 2] += q.z; v[j + 3] += q.w;
                }
        }
        if (p.y_hi && p.y_fmt == 0) {
            uint4 ch[8], cl[8];
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                uint32_t ph[4], pl[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    fp16 h0, l0, h1, l1;
                    split_fp16(v[g * 8 + 2 * e], h0, l0);
                    split_fp16(v[g * 8 + 2 * e + 1], h1, l1);
                    ph[e] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                    pl[e] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
                }
                ch[g] = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                cl[g] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
            }
            stage_and_store(wbuf, lane, ch, rows, (unsigned)col0 * 2u, okmask, cvalid / 8);
            if (p.y_lo) stage_and_store(wbuf, lane, cl, rows + 32, (unsigned)col0 * 2u, okmask, cvalid / 8);
        } else if (p.y_hi) {
            // e4m3 correction plane (common.cuh): per 32 channels [lo8 x32 | hi8 x32] - the round's 64 columns are again one
            // contiguous 128-byte run of the row (yb_co + col0 is a multiple of 32, checked on the host)
            const float lo_scale = exp2f((float)(11 + lo_mx_exp(p.y_fmt))), hi_scale = exp2f((float)lo_mx_exp(p.y_fmt));
            uint4 ch[8], cc[8];
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t ph[16], l8[8], h8[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    float ls[4], hs[4];
                    fp16 hh[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) split_mx(v[hf * 32 + 4 * e + i], lo_scale, hi_scale, hh[i], ls[i], hs[i]);
                    ph[2 * e] = (uint32_t)__half_as_ushort(hh[0]) | ((uint32_t)__half_as_ushort(hh[1]) << 16);
                    ph[2 * e + 1] = (uint32_t)__half_as_ushort(hh[2]) | ((uint32_t)__half_as_ushort(hh[3]) << 16);
                    l8[e] = pack_e4m3x4(ls[0], ls[1], ls[2], ls[3]);
                    h8[e] = pack_e4m3x4(hs[0], hs[1], hs[2], hs[3]);
                }
#pragma unroll
                for (int g = 0; g < 4; ++g) ch[hf * 4 + g] = make_uint4(ph[4 * g], ph[4 * g + 1], ph[4 * g + 2], ph[4 * g + 3]);
                cc[hf * 4 + 0] = make_uint4(l8[0], l8[1], l8[2], l8[3]);
                cc[hf * 4 + 1] = make_uint4(l8[4], l8[5], l8[6], l8[7]);
                cc[hf * 4 + 2] = make_uint4(h8[0], h8[1], h8[2], h8[3]);
                cc[hf * 4 + 3] = make_uint4(h8[4], h8[5], h8[6], h8[7]);
            }
            stage_and_store(wbuf, lane, ch, rows, (unsigned)col0 * 2u, okmask, cvalid / 8);
            if (p.y_lo) stage_and_store(wbuf, lane, cc, rows + 32, (unsigned)col0 * 2u, okmask, cvalid / 8);
        }
        if (p.y_f32) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                if (half * 32 >= cvalid) break;
                uint4 cf[8];
#pragma unroll
                for (int g = 0; g < 8; ++g)
                    cf[g] = make_uint4(__float_as_uint(v[half * 32 + g * 4]), __float_as_uint(v[half * 32 + g * 4 + 1]),
                                       __float_as_uint(v[half * 32 + g * 4 + 2]), __float_as_uint(v[half * 32 + g * 4 + 3]));
                stage_and_store(wbuf, lane, cf, rows + 64, (unsigned)(col0 + half * 32) * 4u, okmask,
                                min(8, (cvalid - half * 32) / 4));
                if (p.colsum && mt < p.m_tiles) {
                    // the 32 rows x 32 columns just stored are still in wbuf: lane = column, sum the in-image rows
                    // (chunk c>>2 of row r sits at slot (c>>2) ^ (r&7): the 32 lanes hit 32 different banks)
                    float sacc = 0.f;
#pragma unroll
                    for (int rr = 0; rr < 32; ++rr) {
                        const float t = *reinterpret_cast<const float*>(wbuf + rr * 128 + ((((lane >> 2) ^ (rr & 7)) << 4) | ((lane & 3) << 2)));
                        if ((okmask >> rr) & 1u) sacc += t;
                    }
                    if (half * 32 + lane < cvalid)
                        p.colsum[((size_t)mt * 4 + quad) * p.Cout + col0 + half * 32 + lane] = sacc;
                    __syncwarp();
                }
            }
        }
    }
    __syncwarp();                                            // the row table is rewritten by the next tile
}

// ---- 8-warp form: 32-column rounds, 64-byte rows.  A round's 32 pixel rows x 64 bytes (32 fp16 / one 32-channel group of the
// e4m3 correction plane / 16 fp32) go through a 2 KB buffer, 16-byte chunk j of row r at slot j ^ ((r >> 1) & 3): both the
// row-wise writes and the 8-rows-per-instruction reads touch eight different 16-byte bank groups per quarter warp.
__device__ __forceinline__ void stage_and_store4(unsigned char* wbuf, int lane, const uint4* chunks,
                                                 const unsigned long long* rowbase, unsigned col_bytes, unsigned okmask,
                                                 int nchunks_valid) {
    const int sw = (lane >> 1) & 3;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(wbuf + lane * 64 + ((j ^ sw) << 4)) = chunks[j];
    __syncwarp();
    const int j = lane & 3;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int pr = it * 8 + (lane >> 2);
        const uint4 v = *reinterpret_cast<const uint4*>(wbuf + pr * 64 + ((j ^ ((pr >> 1) & 3)) << 4));
        if (((okmask >> pr) & 1u) && j < nchunks_valid)
            *reinterpret_cast<uint4*>(rowbase[pr] + col_bytes + (unsigned)(j << 4)) = v;
    }
    __syncwarp();
}

// `sub` (0 / 1): which of the quadrant's two warps this is - it takes the 32-column rounds sub, sub + 2, ...
__device__ __forceinline__ void epilogue_store_coalesced32(const ConvParams& p, uint32_t tacc, int quad, int sub, int lane, int img,
                                                           int oh, int ow, bool pix_ok, int n0, unsigned char* wsm, int mt) {
    unsigned char* wbuf = wsm;
    const float* sbias = reinterpret_cast<const float*>(wsm + EP_WBUF);
    unsigned long long* rows = reinterpret_cast<unsigned long long*>(wsm + EP_WBUF + EP_BIAS);   // [3][32]: y_hi, y_lo, y_f32
    const size_t pix = ((size_t)img * p.Ho + oh) * p.Wo + ow;
    const unsigned okmask = __ballot_sync(0xffffffffu, pix_ok);
    rows[lane] = p.y_hi ? (unsigned long long)(p.y_hi + pix * p.yb_cs + p.yb_co) : 0ull;
    rows[32 + lane] = p.y_lo ? (unsigned long long)(p.y_lo + pix * p.yb_cs + p.yb_co) : 0ull;
    rows[64 + lane] = p.y_f32 ? (unsigned long long)(p.y_f32 + (size_t)img * p.yf_ns + ((size_t)oh * p.Wo + ow) * p.yf_cs + p.yf_co) : 0ull;
    __syncwarp();
    const uint32_t trow = tacc + ((uint32_t)(quad * 32) << 16);
    const int act = p.relu;
    for (int c = sub * 32; c < p.bn; c += 64) {
        const int ncols = min(32, p.bn - c);                 // 16 or 32, warp-uniform
        uint32_t r[32];
        tmem_ld16(trow + (uint32_t)c, r);
        if (ncols > 16) tmem_ld16(trow + (uint32_t)(c + 16), r + 16);
        else {
#pragma unroll
            for (int j = 16; j < 32; ++j) r[j] = 0u;
        }
        tmem_ld_wait();
        const int col0 = n0 + c;
        if (col0 >= p.Cout) continue;                        // warp-uniform
        const int cvalid = min(ncols, p.Cout - col0);        // valid columns in this round (multiple of 8)
        float v[32];
        if (act == 1) round_math<1, 32>(r, v, sbias + c, cvalid, p.acc_scale);
        else if (act == 2) round_math<2, 32>(r, v, sbias + c, cvalid, p.acc_scale);
        else round_math<0, 32>(r, v, sbias + c, cvalid, p.acc_scale);
        if (p.res && pix_ok) {
            const float* rr = p.res + pix * p.res_cs + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                if (j < cvalid) {
                    const float4 q = __ldg(reinterpret_cast<const float4*>(rr + j));
                    v[j] += q.x; v[j + 1] += q.y; v[j + 2] += q.z; v[j + 3] += q.w;
                }
        }
        if (p.y_hi && p.y_fmt == 0) {
            uint4 ch[4], cl[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint32_t ph[4], pl[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    fp16 h0, l0, h1, l1;
                    split_fp16(v[g * 8 + 2 * e], h0, l0);
                    split_fp16(v[g * 8 + 2 * e + 1], h1, l1);
                    ph[e] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                    pl[e] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
                }
                ch[g] = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                cl[g] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
            }
            stage_and_store4(wbuf, lane, ch, rows, (unsigned)col0 * 2u, okmask, cvalid / 8);
            if (p.y_lo) stage_and_store4(wbuf, lane, cl, rows + 32, (unsigned)col0 * 2u, okmask, cvalid / 8);
        } else if (p.y_hi) {
            // e4m3 correction plane (common.cuh): this round is exactly one 32-channel group, [lo8 x32 | hi8 x32] = 64 bytes at
            // byte offset 2 * col0 of the row (yb_co + col0 is a multiple of 32, checked on the host)
            const float lo_scale = exp2f((float)(11 + lo_mx_exp(p.y_fmt))), hi_scale = exp2f((float)lo_mx_exp(p.y_fmt));
            uint4 ch[4], cc[4];
            uint32_t ph[16], l8[8], h8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float ls[4], hs[4];
                fp16 hh[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) split_mx(v[4 * e + i], lo_scale, hi_scale, hh[i], ls[i], hs[i]);
                ph[2 * e] = (uint32_t)__half_as_ushort(hh[0]) | ((uint32_t)__half_as_ushort(hh[1]) << 16);
                ph[2 * e + 1] = (uint32_t)__half_as_ushort(hh[2]) | ((uint32_t)__half_as_ushort(hh[3]) << 16);
                l8[e] = pack_e4m3x4(ls[0], ls[1], ls[2], ls[3]);
                h8[e] = pack_e4m3x4(hs[0], hs[1], hs[2], hs[3]);
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) ch[g] = make_uint4(ph[4 * g], ph[4 * g + 1], ph[4 * g + 2], ph[4 * g + 3]);
            cc[0] = make_uint4(l8[0], l8[1], l8[2], l8[3]);
            cc[1] = make_uint4(l8[4], l8[5], l8[6], l8[7]);
            cc[2] = make_uint4(h8[0], h8[1], h8[2], h8[3]);
            cc[3] = make_uint4(h8[4], h8[5], h8[6], h8[7]);
            stage_and_store4(wbuf, lane, ch, rows, (unsigned)col0 * 2u, okmask, cvalid / 8);
            if (p.y_lo) stage_and_store4(wbuf, lane, cc, rows + 32, (unsigned)col0 * 2u, okmask, cvalid / 8);
        }
        if (p.y_f32) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                if (half * 16 >= cvalid) break;
                uint4 cf[4];
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    cf[g] = make_uint4(__float_as_uint(v[half * 16 + g * 4]), __float_as_uint(v[half * 16 + g * 4 + 1]),
                                       __float_as_uint(v[half * 16 + g * 4 + 2]), __float_as_uint(v[half * 16 + g * 4 + 3]));
                stage_and_store4(wbuf, lane, cf, rows + 64, (unsigned)(col0 + half * 16) * 4u, okmask,
                                 min(4, (cvalid - half * 16) / 4));
                if (p.colsum && mt < p.m_tiles) {
                    // the 32 rows x 16 columns just stored are still in wbuf: lanes 0-15 = column, rows summed in order 0..31 (the
                    // 4-warp form's order: the pooled means stay bit-identical between the two builds); the 16 columns of a row are
                    // 16 different banks
                    const int cc_ = lane & 15;
                    float sacc = 0.f;
#pragma unroll
                    for (int rr = 0; rr < 32; ++rr) {
                        const float t = *reinterpret_cast<const float*>(wbuf + rr * 64 + ((((cc_ >> 2) ^ ((rr >> 1) & 3)) << 4) | ((cc_ & 3) << 2)));
                        if ((okmask >> rr) & 1u) sacc += t;
                    }
                    if (lane < 16 && half * 16 + cc_ < cvalid)
                        p.colsum[((size_t)mt * 4 + quad) * p.Cout + col0 + half * 16 + cc_] = sacc;
                    __syncwarp();
                }
            }
        }
    }
    __syncwarp();                                            // the row table is rewritten by the next tile
}

// ============================================================================================ persistent kernel
// One CTA per SM loops over output tiles (static round-robin).  Three decoupled pipelines:
//   TMA producer (warp 0)  --smem ring(s), full/empty mbarriers-->  MMA issuer (warp 1)
//   MMA issuer             --2 TMEM accumulator stages, tmem_full/tmem_empty-->  epilogue (warps 2-9)
// so the producer prefetches the next tile's operands during the current tile's tail and the epilogue of tile i
// overlaps the MMAs of tile i+1: no per-tile launch / setup / drain bubbles.
//
// HALO = false: generic implicit GEMM.  Stage = A tile (128 pixels x 64 ch of one tap, shifted tiled-TMA box) + B tile.
// HALO = true : 3x3 stride-1.  Tile = 8 pixels along the "fast" image dimension x 16 along the "slow" one.  For each
//   64-channel chunk and each fast-dimension filter offset df, ONE box {64, 8, 18, 1} (8 x 18 pixel column patch, 144
//   canonical SWIZZLE_128B rows, 8-row groups 1024 B apart) serves the three slow-dimension taps: tap ds is the same
//   patch with the descriptor start advanced by ds groups (ds * 1024 B, swizzle-atom aligned).  Activations cross
//   L2->SM 3x instead of 9x per chunk; B tiles ride their own ring, one per tap.
template <int MODE, bool HALO, int CG>
__global__ void __launch_bounds__(UM_THREADS, 1)
conv_persistent_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                       const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                       const ConvParams p) {
    constexpr bool SPLIT = MODE != 0;                      // two operand planes per tensor (hi + lo, or hi + e4m3 correction)
    extern __shared__ unsigned char smem_dyn[];
    __shared__ uint64_t a_full[4], a_empty[4], b_full[9], b_empty[9], acc_full[2], acc_empty[2];
    __shared__ uint32_t s_tmem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) DBG_STAMP(7);                    // kernel entry (before barrier init / TMEM allocation / scale-factor fill)
    const int NA = p.a_stages, NB = p.num_stages;          // generic: only the "B" ring is used (stage = A + B)
    // CTA pair (CG == 2): cluster of two CTAs works on two adjacent M tiles with ONE M=256 MMA stream issued by the leader
    // (rank 0).  Each CTA stages its own A tile and only HALF of the B tile, so the smem read traffic of the MMAs and the TMA
    // write traffic per output pixel drop (the r1 kernel was shared-memory-bandwidth bound: every SS-form MMA re-reads
    // 4 KB of A and 32 B x N of B).  Barriers: "full" lives in the leader (both CTAs' TMA bytes are signalled there),
    // "empty" / "acc_full" are multicast-committed to both CTAs, "acc_empty" collects both epilogues in the leader.
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
    const int cta = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;            // persistent worker (CTA or CTA pair) index
    const int nworkers = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const uint32_t b_bytes = (uint32_t)(p.bn / CG) * UM_BK * 2;
    const uint32_t a_stage_bytes = (SPLIT ? 2u : 1u) * (HALO ? HALO_PATCH_BYTES : UM_A_BYTES);
    const uint32_t b_stage_bytes = (SPLIT ? 2u : 1u) * b_bytes + (HALO ? 0u : a_stage_bytes);
    // 1024-byte alignment by offset (not by integer round trip), so the compiler keeps these pointers in the shared window
    unsigned char* a_ring = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    unsigned char* b_ring = a_ring + (HALO ? (size_t)NA * a_stage_bytes : 0);
    unsigned char* ep_buf = b_ring + (size_t)NB * b_stage_bytes;      // UM_EPI x EP_WARP_BYTES (one block per epilogue warp)
    const int n_tiles = (p.Cout + p.bn - 1) / p.bn;
    const int total_tiles = ((p.m_tiles + CG - 1) / CG) * n_tiles;                 // CG == 2: tiles of 2 M tiles
    const int taps = p.ks * p.ks, pad = p.ks / 2;
    const uint32_t acc_cols = tmem_cols_for(p.bn);           // accumulator columns per tile (power of two)
    const uint32_t tmem_cols = acc_cols * 2 > 512 ? 512 : acc_cols * 2;      // two accumulator stages

    if (threadIdx.x == 0) {
        for (int s = 0; s < 4; ++s) { mbar_init(&a_full[s], SPLIT ? 2 : 1); mbar_init(&a_empty[s], 1); }     // one arrive per operand plane
        for (int s = 0; s < 9; ++s) { mbar_init(&b_full[s], SPLIT ? 2 : 1); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], UM_EPI * CG); }
        fence_barrier_init();
    }
    if (warp == 1) {
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(tmem_cols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(tmem_cols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();     // pair: the peer's barriers must be initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    if (threadIdx.x == 0) DBG_STAMP(0);
    // Programmatic dependent launch (far3d_conv_umma_tune8): everything above touches no global memory, so it may run while the
    // previous kernel of the stream drains; from here on this grid reads that kernel's output (and overwrites buffers it may still
    // read).  Without the launch attribute both instructions are no-ops.  The trigger comes first: the next conv's CTAs are
    // placed as soon as this grid's CTAs leave their SMs (a conv CTA holds the whole shared memory, so never before) and run
    // THEIR prologue behind our tail.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // tile id -> (m tile, n tile); m tile -> image + pixel origin
    auto decode = [&](int tile, int& img, int& c0, int& c1, int& n0, int* mt_out = nullptr) {
        const int nt = tile % n_tiles;
        int mt = (tile / n_tiles) * CG + (int)rank;
        n0 = nt * p.bn;
        if (mt_out) *mt_out = mt;
        if (mt >= p.m_tiles) { img = p.N; c0 = 0; c1 = 0; return; }   // odd tile count: the peer's phantom tile (TMA zero-fills, nothing is stored)
        if (HALO) {
            const int tf = mt % p.tiles_f; mt /= p.tiles_f;
            const int ts = mt % p.tiles_s; img = mt / p.tiles_s;
            c0 = tf * HALO_F; c1 = ts * HALO_S;            // fast / slow pixel origin
        } else {
            const int tx = mt % p.tiles_w; mt /= p.tiles_w;
            const int ty = mt % p.tiles_h; img = mt / p.tiles_h;
            c0 = tx * p.tw; c1 = ty * p.th;                // ow0 / oh0
        }
    };

    // Ring bookkeeping without divisions: slot index + phase bit, advanced together.  Both the producer and the MMA issuer run
    // their loops WARP-UNIFORMLY (all 32 lanes in lockstep) and guard only the TMA / tcgen05 / expect_tx instructions with
    // elect_one_sync(): inside an `if (lane == 0)` region every UTMALDG / UTCHMMA is wrapped in a vote-and-branch waterfall, and
    // the r2 ncu source page showed both loops at ~250 scalar instructions (~1000 clk) per stage - more than the 512 clk of
    // tensor-pipe time a stage holds at N = 128 (profiles/r2e_conv_issue_loops.txt).
    if (warp == 0 || warp >= 2 + UM_EPI) {
        // ================= TMA producers =================
        // UM_PRODUCERS warps (warp 0, warps 6..) share the load stream.  In a bare ingest loop (tools/tma_rate.cu) the loads one
        // thread issues complete one per ~750-900 clk and only more issuing warps raise the rate (12 KB tiles: 14 bytes/clk/SM from
        // one thread, 57 from four warps); inside this kernel the ring is paced by the MMAs, and more producers change nothing.
        // The stream is cut into units - one operand plane of one stage (the B tile, or A tile + B tile in the generic form) or of
        // one halo patch - dealt round-robin to the producers.  Every producer walks the whole stage sequence (slot + phase
        // bookkeeping is a few instructions) and issues only its units; "full" barriers count one arrive.expect_tx per plane.
        constexpr int PL = SPLIT ? 2 : 1;
        const int pw = warp == 0 ? 0 : warp - (1 + UM_EPI);   // producer index 0 .. UM_PRODUCERS-1
        int sb = 0, sa = 0;                                // next B / A ring slot to fill
        uint32_t pb = 0, pa = 0;                           // its phase bit ("empty" is awaited with parity phase ^ 1: a fresh barrier passes)
        int ub = 0, ua = 0;                                // owner of the next B / A unit
        long long w_bempty = 0, w_aempty = 0;
        // "full" barriers live in the leader CTA (pair: shared::cluster address of rank 0's copy)
        const uint32_t bfull0 = CG == 2 ? mapa_rank(smem_u32(&b_full[0]), 0) : smem_u32(&b_full[0]);
        const uint32_t afull0 = CG == 2 ? mapa_rank(smem_u32(&a_full[0]), 0) : smem_u32(&a_full[0]);
        // which planes of the next stage are mine (bit per plane); advances the owner counter past the stage
        auto my_planes = [&](int& u) {
            uint32_t m = 0;
#pragma unroll
            for (int pl = 0; pl < PL; ++pl) {
                if (u == pw) m |= 1u << pl;
                if (++u == UM_PRODUCERS) u = 0;
            }
            return m;
        };
        for (int tile = cta; tile < total_tiles; tile += nworkers) {
            int img, c0, c1, n0;
            decode(tile, img, c0, c1, n0);
            const int nb0 = n0 + (int)rank * (p.bn / CG);          // this CTA's rows of the B tile
            if (HALO) {
                // one (8+2) x (16+2) pixel patch per 64-channel chunk serves all nine taps.  The patch of the NEXT chunk
                // (possibly of the next tile) is requested after the first few B loads of this chunk: by then the MMAs
                // are done with the ring slot it reuses, so the request never blocks the B stream.
                auto load_a = [&](int t_tile, int kc) {
                    const uint32_t mine = my_planes(ua);
                    if (mine) {
                        int ti, tc0, tc1, tn0;
                        decode(t_tile, ti, tc0, tc1, tn0);
                        WAIT_ACC(w_aempty, &a_empty[sa], pa ^ 1u);
                        if (elect_one_sync()) {
                            unsigned char* d = a_ring + (size_t)sa * a_stage_bytes;
                            const uint32_t fb = afull0 + 8u * (uint32_t)sa;
#pragma unroll
                            for (int pl = 0; pl < PL; ++pl)
                                if (mine & (1u << pl)) {
                                    if (rank == 0) mbar_expect_tx(&a_full[sa], CG * HALO_PATCH_TX);
                                    tma_ld_4d<CG>(d + pl * HALO_PATCH_BYTES, pl ? &tmA_lo : &tmA_hi, fb, kc * UM_BK, tc0 - 1, tc1 - 1, ti);
                                }
                        }
                        __syncwarp();
                    }
                    if (++sa == NA) { sa = 0; pa ^= 1u; }
                };
                if (tile == cta) load_a(tile, 0);                     // the very first patch of this worker
                const int tpre = NB < 8 ? NB : 8;
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    for (int t = 0; t < 9; ++t) {
                        if (t == tpre) {                               // next patch: next chunk, or chunk 0 of the next tile
                            if (kc + 1 < p.kchunks) load_a(tile, kc + 1);
                            else if (tile + nworkers < total_tiles) load_a(tile + nworkers, 0);
                        }
                        const uint32_t mine = my_planes(ub);
                        if (mine && (!p.b_resident || tile == cta)) {       // resident weights: loaded with the first tile only
                            const int df = t / 3, ds = t - df * 3;
                            const int tap = p.transposed ? df * 3 + ds : ds * 3 + df;   // tap = ky*3 + kx
                            WAIT_ACC(w_bempty, &b_empty[sb], pb ^ 1u);
                            if (elect_one_sync()) {
                                unsigned char* d = b_ring + (size_t)sb * b_stage_bytes;
                                const uint32_t fb = bfull0 + 8u * (uint32_t)sb;
#pragma unroll
                                for (int pl = 0; pl < PL; ++pl)
                                    if (mine & (1u << pl)) {
                                        if (rank == 0) mbar_expect_tx(&b_full[sb], CG * b_bytes);
                                        tma_ld_3d<CG>(d + pl * b_bytes, pl ? &tmB_lo : &tmB_hi, fb, kc * UM_BK, tap, nb0);
                                    }
                            }
                            __syncwarp();
                        }
                        if (++sb == NB) { sb = 0; pb ^= 1u; }
                    }
                }
            } else {
                for (int tap = 0; tap < taps; ++tap) {
                    const int ky = tap / p.ks, kx = tap - ky * p.ks;
                    // stride 1: shifted box; stride 2: the tensor is viewed as {2C, W/2, 2, H/2, N}, a tap is again a dense box
                    const int dy = ky - pad, dx = kx - pad;
                    const int hpar = dy & 1, wpar = dx & 1;
                    const int hoff = (dy - hpar) / 2, woff = (dx - wpar) / 2;
                    for (int kc = 0; kc < p.kchunks; ++kc) {
                        const uint32_t mine = my_planes(ub);
                        if (mine) {
                            const int ch0 = kc * UM_BK;
                            WAIT_ACC(w_bempty, &b_empty[sb], pb ^ 1u);
                            if (elect_one_sync()) {
                                unsigned char* sa_ = b_ring + (size_t)sb * b_stage_bytes;
                                const uint32_t fb = bfull0 + 8u * (uint32_t)sb;
#pragma unroll
                                for (int pl = 0; pl < PL; ++pl)
                                    if (mine & (1u << pl)) {
                                        const CUtensorMap* mapA = pl ? &tmA_lo : &tmA_hi;
                                        if (rank == 0) mbar_expect_tx(&b_full[sb], CG * (UM_A_BYTES + b_bytes));
                                        if (p.stride == 1) tma_ld_4d<CG>(sa_ + pl * UM_A_BYTES, mapA, fb, ch0, c0 + dx, c1 + dy, img);
                                        else tma_ld_5d<CG>(sa_ + pl * UM_A_BYTES, mapA, fb, wpar * p.x_cs + p.x_co + ch0, c0 + woff, hpar, c1 + hoff, img);
                                        tma_ld_3d<CG>(sa_ + a_stage_bytes + pl * b_bytes, pl ? &tmB_lo : &tmB_hi, fb, ch0, tap, nb0);
                                    }
                            }
                            __syncwarp();
                        }
                        if (++sb == NB) { sb = 0; pb ^= 1u; }
                    }
                }
            }
        }
        if (warp == 0 && lane == 0) { DBG_STAMP(5); DBG_PUT(11, w_bempty); DBG_PUT(12, w_aempty); }
    } else if (warp == 1 && rank == 0) {
        // ================= MMA issuer (pair: leader CTA only) =================
        // All 32 lanes run this loop in lockstep; only the tcgen05 instructions are issued by one elected lane.
        const uint32_t idesc = umma_idesc_fp16(p.bn, 128 * CG);
        int sb = 0, sa = 0;
        uint32_t pb = 0, pa = 0;
        // descriptors of ring slot 0; a slot / plane / tap offset is an addition to the 14-bit (address >> 4) field
        const uint64_t bdesc0 = umma_desc_sw128(smem_u32(b_ring) + (HALO ? 0u : a_stage_bytes));
        const uint64_t adesc0 = HALO ? umma_desc_sw128(smem_u32(a_ring), HALO_PF * 128) : umma_desc_sw128(smem_u32(b_ring));
        const uint32_t b_step = b_stage_bytes >> 4, a_step = a_stage_bytes >> 4;
        const uint32_t a_lo_off = (HALO ? HALO_PATCH_BYTES : UM_A_BYTES) >> 4, b_lo_off = b_bytes >> 4;
        int lt = 0;                                        // local tile counter
        long long w_accempty = 0, w_afull = 0, w_bfull = 0;
        for (int tile = cta; tile < total_tiles; tile += nworkers, ++lt) {
            const int as = lt & 1;
            WAIT_ACC(w_accempty, &acc_empty[as], (((uint32_t)lt >> 1) & 1u) ^ 1u);   // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t tacc = tmem_base + (uint32_t)as * acc_cols;
            if (HALO) {
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    WAIT_ACC(w_afull, &a_full[sa], pa);
                    const int ksteps = (min(UM_BK, p.Cin - kc * UM_BK) + 15) / 16;
                    const uint64_t a_hi_base = adesc0 + (uint64_t)((uint32_t)sa * a_step);
                    const bool last_chunk = kc == p.kchunks - 1;
                    for (int t = 0; t < 9; ++t) {
                        const int df = t / 3, ds = t - df * 3;
                        if (!p.b_resident || lt == 0) WAIT_ACC(w_bfull, &b_full[sb], pb);
                        tc_fence_after();
                        // tile pixel (s, f) under tap (ds, df) is patch row (s + ds) * 10 + f + df: 8-row groups every 10 rows,
                        // start (ds * 10 + df) rows (128 B each = 8 address units) into the patch
                        const uint64_t a_hi0 = a_hi_base + (uint64_t)((uint32_t)(ds * HALO_PF + df) * 8u);
                        const uint64_t b_hi0 = bdesc0 + (uint64_t)((uint32_t)sb * b_step);
                        if (elect_one_sync()) {
                            mma_chunk<MODE, CG>(tacc, a_hi0, a_hi0 + a_lo_off, b_hi0, b_hi0 + b_lo_off, idesc, (kc | t) ? 1u : 0u, ksteps);
                            if (CG == 2) {
                                umma2_commit(&b_empty[sb]);
                                if (t == 8) umma2_commit(&a_empty[sa]);
                                if (last_chunk && t == 8) umma2_commit(&acc_full[as]);
                            } else {
                                umma_commit(&b_empty[sb]);
                                if (t == 8) umma_commit(&a_empty[sa]);
                                if (last_chunk && t == 8) umma_commit(&acc_full[as]);
                            }
                        }
                        __syncwarp();
                        if (++sb == NB) { sb = 0; pb ^= 1u; }
                    }
                    if (++sa == NA) { sa = 0; pa ^= 1u; }
                }
                if (lt == 0 && lane == 0) DBG_STAMP(1);               // first tile's MMAs issued
            } else {
                for (int tap = 0; tap < taps; ++tap) {
                    for (int kc = 0; kc < p.kchunks; ++kc) {
                        WAIT_ACC(w_bfull, &b_full[sb], pb);
                        tc_fence_after();
                        const int ksteps = (min(UM_BK, p.Cin - kc * UM_BK) + 15) / 16;
                        const uint64_t a_hi0 = adesc0 + (uint64_t)((uint32_t)sb * b_step);
                        const uint64_t b_hi0 = bdesc0 + (uint64_t)((uint32_t)sb * b_step);
                        const bool last = tap == taps - 1 && kc == p.kchunks - 1;
                        if (elect_one_sync()) {
                            mma_chunk<MODE, CG>(tacc, a_hi0, a_hi0 + a_lo_off, b_hi0, b_hi0 + b_lo_off, idesc, (tap | kc) ? 1u : 0u, ksteps);
                            if (CG == 2) {
                                umma2_commit(&b_empty[sb]);                  // frees the smem stage (in both CTAs) when these MMAs retire
                                if (last) umma2_commit(&acc_full[as]);       // accumulator complete
                            } else {
                                umma_commit(&b_empty[sb]);
                                if (last) umma_commit(&acc_full[as]);
                            }
                        }
                        __syncwarp();
                        if (++sb == NB) { sb = 0; pb ^= 1u; }
                    }
                }
            }
        }
        if (lane == 0) { DBG_STAMP(2); DBG_PUT(8, w_accempty); DBG_PUT(9, w_afull); DBG_PUT(10, w_bfull); DBG_PUT(15, lt); }
    } else if (warp >= 2 && warp < 2 + UM_EPI) {
        // ================= epilogue: TMEM -> registers -> global =================
        const int quad = warp & 3;                          // TMEM lane quadrant this warp may access
        const int sub = (warp - 2) >> 2;                    // 8-warp form: which of the quadrant's two warps (alternate column rounds)
        const int m = quad * 32 + lane;                     // row of the tile = pixel
        int bias_n0 = -1;
        int lt = 0;
        long long w_accfull = 0, w_busy = 0;
        const uint32_t acc_empty_leader = CG == 2 ? mapa_rank(smem_u32(&acc_empty[0]), 0) : 0u;
        for (int tile = cta; tile < total_tiles; tile += nworkers, ++lt) {
            const int as = lt & 1;
            int img, c0, c1, n0, mt;
            decode(tile, img, c0, c1, n0, &mt);
            int oh, ow;
            if (HALO) {
                const int f = c0 + (m & (HALO_F - 1)), s = c1 + (m >> 3);
                oh = p.transposed ? f : s; ow = p.transposed ? s : f;
            } else {
                const int hh = m / p.tw;
                oh = c1 + hh; ow = c0 + (m - hh * p.tw);
            }
            unsigned char* wsm = ep_buf + (warp - 2) * EP_WARP_BYTES;
            float* sbias = reinterpret_cast<float*>(wsm + EP_WBUF);
            if (n0 != bias_n0) {                             // (re)load this warp's bias slice; global loads batched, off the critical path
                for (int i = lane; i < p.bn; i += 32) sbias[i] = (p.bias && n0 + i < p.Cout) ? __ldg(p.bias + n0 + i) : 0.f;
                bias_n0 = n0;
                __syncwarp();
            }
            WAIT_ACC(w_accfull, &acc_full[as], ((uint32_t)lt >> 1) & 1u);
            tc_fence_after();
#ifdef FAR3D_CONV_WAITSTATS
            const long long t_busy0 = clock64();
#endif
            if (lt == 0 && threadIdx.x == 64) DBG_STAMP(6);          // first accumulator ready
            const bool pix_ok = (oh < p.Ho) && (ow < p.Wo) && (img < p.N);
            if (img >= p.N) img = 0;                         // phantom tile: keep the address arithmetic in range, nothing is stored
            if (UM_EPI == 8) {
                if ((p.Cout & 7) == 0)
                    epilogue_store_coalesced32(p, tmem_base + (uint32_t)as * acc_cols, quad, sub, lane, img, oh, ow, pix_ok, n0, wsm, mt);
                else
                    epilogue_store(p, tmem_base + (uint32_t)as * acc_cols, quad, img, oh, ow, pix_ok, n0, sub, 2);
            } else if ((p.Cout & 7) == 0)
                epilogue_store_coalesced(p, tmem_base + (uint32_t)as * acc_cols, quad, lane, img, oh, ow,
                                         pix_ok, n0, wsm, mt);
            else
                epilogue_store(p, tmem_base + (uint32_t)as * acc_cols, quad, img, oh, ow, pix_ok, n0);
            tc_fence_before();
            __syncwarp();
#ifdef FAR3D_CONV_WAITSTATS
            w_busy += clock64() - t_busy0;
#endif
            if (lane == 0) {                                 // this warp's quarter of the accumulator is drained
                if (CG == 2) mbar_arrive_cluster_relaxed(acc_empty_leader + 8u * (uint32_t)as);
                else mbar_arrive_relaxed(&acc_empty[as]);
            }
        }
        if (threadIdx.x == 64) { DBG_STAMP(3); DBG_PUT(13, w_accfull); DBG_PUT(14, w_busy); }
    }

    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();     // pair: nobody leaves while the other CTA may still touch its smem / barriers
    if (threadIdx.x == 0) DBG_STAMP(4);
    if (warp == 1) {
        tc_fence_after();
        if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)f;
    }
    return fn;
}

static int encode(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                  const cuuint32_t* box) {
    EncodeTiledFn fn = get_encode();
    if (!fn) return fail(FAR3D_E_CUDA, "%scuTensorMapEncodeTiled unavailable (no driver)", "");
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(FAR3D_E_CUDA, "%scuTensorMapEncodeTiled failed (CUresult %ld, rank %ld)", "", (long)r, rank);
    return FAR3D_OK;
}

// tuning knobs (0 = heuristic): N tile, ring depth, persistent grid size, halo kernel on/off (-1 = off)
static int g_force_bn = 0, g_force_stages = 0, g_force_grid = 0, g_halo = 0, g_smem_reserve = 0;
static long long* g_dbg = nullptr;
static int g_num_sms = 0;
static int g_cg = 0;             // CTA-pair (cta_group::2) kernel: 0 = heuristic, 1 = never, 2 = whenever legal
static int g_pdl = 0;            // launch with cudaLaunchAttributeProgrammaticStreamSerialization (far3d_conv_umma_tune8)

static int num_sms() {
    if (!g_num_sms) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            g_num_sms = n;
        else
            g_num_sms = 148;
    }
    return g_num_sms;
}

template <typename K>
static int launch(K kernel, int grid, int cluster, size_t smem, cudaStream_t st, const CUtensorMap& a0, const CUtensorMap& a1,
                  const CUtensorMap& b0, const CUtensorMap& b1, const ConvParams& p) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(FAR3D_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(UM_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (cluster > 1) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = cluster; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
        ++na;
    }
    if (g_pdl) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = at; cfg.numAttrs = na;
    e = cudaLaunchKernelEx(&cfg, kernel, a0, a1, b0, b1, p);
    if (e != cudaSuccess) return fail(FAR3D_E_CUDA, "cudaLaunchKernelEx (conv): %s", cudaGetErrorString(e));
    return launched("conv_persistent_kernel");
}

}  // namespace far3d

using namespace far3d;

// tuning hooks for experiments (not part of the reference-facing ABI)
extern "C" void far3d_conv_umma_tune(int bn, int stages) { g_force_bn = bn; g_force_stages = stages; }
extern "C" void far3d_conv_umma_tune2(int grid, int halo) { g_force_grid = grid; g_halo = halo; }
extern "C" void far3d_conv_umma_debug(void* buf) { g_dbg = (long long*)buf; }
extern "C" void far3d_conv_umma_tune4(int cg) { g_cg = cg; }
extern "C" void far3d_conv_umma_tune8(int pdl) { g_pdl = pdl != 0; }
extern "C" void far3d_conv_umma_tune7(int smem_reserve_bytes) { g_smem_reserve = smem_reserve_bytes < 0 ? 0 : smem_reserve_bytes; }

// tcgen05.mma adds each instruction's K=16 dot products into the fp32 TMEM accumulator with TRUNCATION (round toward zero), not
// round-to-nearest: every accumulating MMA loses on average half an ulp of the running sum, always toward zero.  Measured
// (tools/conv_bias_probe.py, profiles/r1d_conv_bias_probe.txt): the result of a conv is the exact one times
// (1 - 6.4e-8 * k-steps) for monotonically growing sums and (1 - 4.8e-8 * k-steps) for zero-mean operands, with three
// accumulating MMAs per k-step - i.e. 1.6e-8 ... 2.15e-8 of the final value per MMA (0.5 ulp x E[ulp / value] x the mean
// fill of the running sum) - while the operand split itself is exact to 5e-8.  The loss is a pure scale, so it compounds
// linearly through the 106 convolutions of the image branch (5.6e-4 on feat_flatten at cfg-2, uncorrected).  The epilogue
// multiplies the accumulator by 1 + g_rz_loss_per_mma * (accumulating MMAs of the tile) to take the EXPECTED loss out; what
// remains is the zero-mean part of the truncation (~1e-5 per layer at K = 6912).  The constant is the zero-mean-operand one
// (BN-folded weights): with it feat_flatten at cfg-2 is 1.07e-4 from the reference (scale error +1.6e-6) instead of 5.6e-4
// (scale error -2.0e-4), and 99.6 % instead of 92.8 % of the last-layer class-logit rows are within 1e-3
// (profiles/r1d_conv_bias_probe.txt).  far3d_conv_umma_tune6 sets the constant (0 = no compensation; tools / tests).
static float g_rz_loss_per_mma = 1.6e-8f;
extern "C" void far3d_conv_umma_tune6(float loss_per_mma) { g_rz_loss_per_mma = loss_per_mma; }

// x_fmt: format of the x_lo / w_lo planes - 0 = fp16 residual planes (fp16x3), FAR3D_LO_MX(EA) = e4m3 correction planes (fp16mx;
// w_exp = the weights' pre-scale exponent: w_hi8 = e4m3(w_hi * 2^w_exp), w_lo8 = e4m3(w_lo * 2^(w_exp + 11)), and the fp16
// weight plane holds w_hi * 2^(11 + EA + w_exp)).
// y_fmt: format of the y_lo plane this conv writes (independent of the input format).
static int conv_impl(const void* x_hi, const void* x_lo, int N, int H, int W, int x_cs, int x_co, int Cin,
                     const void* w_hi, const void* w_lo, const float* bias, int Cout, int ksize, int stride,
                     int relu, const float* res, int res_cs, float* y_f32, int yf_cs, int yf_co, int64_t yf_ns, void* y_hi,
                     void* y_lo, int yb_cs, int yb_co, void* stream, float* colsum = nullptr, int* m_tiles_out = nullptr,
                     int x_fmt = 0, int w_exp = 0, int y_fmt = 0) {
    FAR3D_REQUIRE(x_hi && w_hi && (y_f32 || y_hi), "null pointer");
    FAR3D_REQUIRE((x_lo == nullptr) == (w_lo == nullptr), "x_lo and w_lo must both be given (split mode) or both NULL");
    const bool mx = x_fmt != 0;
    FAR3D_REQUIRE(!mx || (x_lo && Cin % 32 == 0 && x_cs % 32 == 0 && x_co % 32 == 0), "fp16mx input: correction planes, Cin / x_cs / x_co %% 32 == 0");
    FAR3D_REQUIRE(!mx || (x_fmt >= 64 - 40 && x_fmt <= 64 + 40 && w_exp >= -60 && w_exp <= 60), "fp16mx exponents out of range");
    FAR3D_REQUIRE(y_fmt == 0 || !y_lo || (Cout % 32 == 0 && yb_cs % 32 == 0 && yb_co % 32 == 0 && y_fmt >= 24 && y_fmt <= 104),
                  "e4m3 correction plane output: Cout / yb_cs / yb_co %% 32 == 0");
    FAR3D_REQUIRE(N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "non-positive size");
    FAR3D_REQUIRE((ksize == 1 || ksize == 3) && (stride == 1 || (stride == 2 && ksize == 3)), "ksize/stride unsupported");
    FAR3D_REQUIRE(Cin % 16 == 0 && x_cs % 8 == 0 && x_co % 8 == 0, "Cin %% 16, x_cs %% 8, x_co %% 8");
    FAR3D_REQUIRE((uintptr_t)x_hi % 16 == 0 && (uintptr_t)w_hi % 16 == 0, "16-byte aligned operands");
    FAR3D_REQUIRE(!y_f32 || (yf_cs % 4 == 0 && yf_co % 4 == 0 && (uintptr_t)y_f32 % 16 == 0), "fp32 output alignment");
    FAR3D_REQUIRE(!y_hi || (yb_cs % 8 == 0 && yb_co % 8 == 0 && (uintptr_t)y_hi % 16 == 0), "fp16 output alignment");
    if (stride == 2) FAR3D_REQUIRE(H % 2 == 0 && W % 2 == 0 && Cin % 64 == 0, "stride 2 needs even H, W and Cin %% 64 == 0");
    const bool split = x_lo != nullptr;
    const int pad = ksize / 2;
    ConvParams p = {};
    p.N = N; p.H = H; p.W = W;
    p.Ho = (H + 2 * pad - ksize) / stride + 1; p.Wo = (W + 2 * pad - ksize) / stride + 1;
    p.Cin = Cin; p.Cout = Cout; p.ks = ksize; p.stride = stride;
    p.x_cs = x_cs; p.x_co = x_co; p.relu = relu; p.bias = bias;
    p.y_f32 = y_f32; p.yf_cs = yf_cs; p.yf_co = yf_co;
    p.yf_ns = yf_ns > 0 ? yf_ns : (long long)p.Ho * p.Wo * yf_cs;
    p.y_hi = (fp16*)y_hi; p.y_lo = (fp16*)y_lo; p.yb_cs = yb_cs; p.yb_co = yb_co;
    p.kchunks = (Cin + UM_BK - 1) / UM_BK;
    {   // accumulating MMAs that carry data: taps x ceil(Cin / 16) k-steps x (3 fp16x3 | 2 fp16mx | 1 plain); zero-padded
        // k-steps add exact zeros
        const int mmas = ksize * ksize * ((Cin + 15) / 16) * (x_lo ? (mx ? 2 : 3) : 1);
        p.acc_scale = 1.f + g_rz_loss_per_mma * (float)mmas;
    }
    p.y_fmt = y_lo ? y_fmt : 0;
    if (mx) {
        // both correction products carry the factor 2^(11 + EA + w_exp) (lo8 = lo * 2^(11+EA) times w_hi8 = w_hi * 2^w_exp;
        // hi8 = hi * 2^EA times w_lo8 = w_lo * 2^(w_exp+11)); the caller's fp16 weight plane carries the same factor (exact:
        // a power of two), so all four MMAs of a 32-channel group add into one accumulator at one scale and the epilogue takes
        // it out together with the truncation compensation - no scale factors in TMEM, whole 256-column accumulator stages
        p.acc_scale *= exp2f((float)-(11 + (x_fmt - 64) + w_exp));
    }
    p.dbg = g_dbg;
    p.cm = 1;
    p.res = res; p.res_cs = res_cs;
    p.colsum = colsum;
    FAR3D_REQUIRE(!colsum || (y_f32 && Cout % 8 == 0), "column sums ride on the coalesced fp32 epilogue (Cout %% 8 == 0)");
    FAR3D_REQUIRE(!res || (res_cs % 4 == 0 && (uintptr_t)res % 16 == 0), "residual alignment");
    cudaStream_t st = (cudaStream_t)stream;
    const int sp = split ? 2 : 1;
    const int sms = num_sms();
    const size_t EP_BYTES = UM_EPI * EP_WARP_BYTES;      // epilogue staging + bias slice + row table, per epilogue warp
    // g_smem_reserve (far3d_conv_umma_tune7): bytes of the SM's shared memory left free, so that CTAs of the other frame's head
    // kernels (aggregation: 25 KB, 64 registers x 256 threads = exactly what the conv CTA's 192 x 255 leave) can be resident
    // NEXT TO a persistent conv CTA instead of waiting for the gap between two conv launches
    const size_t SMEM_BUDGET = 225 * 1024 - EP_BYTES - (size_t)g_smem_reserve;

    auto mapB = [&](CUtensorMap* tm, const void* base, int rows) -> int {
        cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)(ksize * ksize), (cuuint64_t)Cout};
        cuuint64_t str[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)ksize * ksize * Cin * 2};
        cuuint32_t box[3] = {(cuuint32_t)UM_BK, 1, (cuuint32_t)rows};
        return encode(tm, base, 3, dims, str, box);
    };
    CUtensorMap tmA_hi, tmA_lo, tmB_hi, tmB_lo;
    int rc;
    const bool halo = (ksize == 3 && stride == 1 && g_halo >= 0);

    // ---- M tiling.  halo: 8 x 16 pixel tiles, 8-pixel side along W (0) or H (1), whichever needs fewer tiles;
    //      generic: th x tw = 128 output pixels with the fewest tiles
    int Fd = 0, Sd = 0;
    if (halo) {
        const long t0 = (long)((W + HALO_F - 1) / HALO_F) * ((H + HALO_S - 1) / HALO_S);
        const long t1 = (long)((H + HALO_F - 1) / HALO_F) * ((W + HALO_S - 1) / HALO_S);
        p.transposed = t1 < t0 ? 1 : 0;
        Fd = p.transposed ? H : W; Sd = p.transposed ? W : H;
        p.tiles_f = (Fd + HALO_F - 1) / HALO_F; p.tiles_s = (Sd + HALO_S - 1) / HALO_S;
        p.m_tiles = N * p.tiles_f * p.tiles_s;
    } else {
        int best_tw = 128; long best_tiles = -1;
        for (int tw = 8; tw <= 128; tw <<= 1) {
            int th = 128 / tw;
            long t = (long)((p.Wo + tw - 1) / tw) * ((p.Ho + th - 1) / th);
            if (best_tiles < 0 || t < best_tiles || (t == best_tiles && tw > best_tw)) { best_tiles = t; best_tw = tw; }
        }
        p.tw = best_tw; p.th = 128 / best_tw;
        p.tiles_w = (p.Wo + p.tw - 1) / p.tw; p.tiles_h = (p.Ho + p.th - 1) / p.th;
        const long m_tiles = (long)N * p.tiles_w * p.tiles_h;
        if (m_tiles * ((Cout + 15) / 16) > 0x7fffffffL) return fail(FAR3D_E_UNSUPPORTED, "%stoo many tiles", "");
        p.m_tiles = (int)m_tiles;
    }

    // ---- N tile: whole Cout when it fits one MMA (<= 256), else the divisor-friendly size with the fewest tiles; problems
    //      that would leave more than half of the SMs idle (decoder GEMMs, the 20 x 30 maps) split N further
    int bn = g_force_bn;
    // an e4m3 correction plane is written in whole 32-channel groups, so the N tile stays a multiple of 32 then
    const int bn_max = 256;
    const int bn_gran = (y_lo && y_fmt != 0) ? 32 : 16;
    if (bn <= 0) {
        if (Cout <= bn_max) bn = (Cout + 15) / 16 * 16;
        else if (Cout <= 256) bn = 128;
        else {
            // the divisor-friendly size with the fewest wasted columns (r2: choosing by useful columns per operand byte - Cout 1024
            // as 5 x 224 instead of 8 x 128 - lost on the 20 x 30 maps, 63 vs 47 us: fewer, longer work items quantise worse)
            const int cand[] = {256, 224, 192, 160, 128};
            long best = -1;
            for (int c : cand) {
                if (c > bn_max) continue;
                const long waste = (long)((Cout + c - 1) / c) * c - Cout;
                if (best < 0 || waste < best) { best = waste; bn = c; }
            }
        }
        // 256-wide tiles pay (less operand traffic per MMA) when a worker gets several of them; with fewer than two per CTA pair
        // the shorter tiles win: the epilogue of one overlaps the MMAs of the next (c5 / FPN maps of 40 x 60 and below)
        if (bn > 128 && Cout % 128 == 0 && (long)((p.m_tiles + 1) / 2) * ((Cout + bn - 1) / bn) < 2L * (sms / 2)) bn = 128;
        // too few work items for the SMs (20 x 30 maps, decoder GEMMs): split N further (half, rounded up to whole groups)
        while (bn >= 128 && (long)p.m_tiles * ((Cout + bn - 1) / bn) * 2 <= sms) {
            const int nb = (bn / 2 + bn_gran - 1) / bn_gran * bn_gran;
            if (nb >= bn) break;
            bn = nb;
        }
        // single-CTA halo kernel: two patches + two whole-B stages must fit (a CTA pair stages half of B and always fits)
        const int cg_req = (g_cg == 1 || p.m_tiles < 2) ? 1 : 2;
        auto halo_fits = [&](int b) { return 2 * (size_t)sp * HALO_PATCH_BYTES + 2 * (size_t)sp * (b / cg_req) * UM_BK * 2 <= SMEM_BUDGET; };
        while (halo && bn % (2 * bn_gran) == 0 && !halo_fits(bn)) bn /= 2;
        while (halo && !halo_fits(bn) && bn > bn_gran) bn = (bn - 1) / bn_gran * bn_gran;   // e.g. 224 with 32-channel groups: 192
    }
    FAR3D_REQUIRE(bn <= bn_max && bn % bn_gran == 0, "N tile not usable in this operand format");
    FAR3D_REQUIRE(bn >= 16 && bn <= 256 && bn % 16 == 0, "bad N tile");
    p.bn = bn;
    const int n_tiles = (Cout + bn - 1) / bn;
    if (m_tiles_out) *m_tiles_out = p.m_tiles;
    // CTA pair (cta_group::2): two M tiles per MMA stream and half a B tile per CTA - less smem traffic per FLOP and a
    // deeper ring in the same smem; measured faster on every layer class with >= 2 M tiles (r1 profile)
    const int cg = (g_cg == 1 || p.m_tiles < 2) ? 1 : 2;
    size_t smem;

    if (halo) {
        const size_t b_stage = (size_t)sp * (bn / cg) * UM_BK * 2;
        p.a_stages = 2;                                  // one patch in use + the next chunk's in flight (a patch feeds 9 B stages)
        size_t a_bytes = (size_t)p.a_stages * sp * HALO_PATCH_BYTES;
        int nb = (int)((SMEM_BUDGET - a_bytes) / b_stage);
        if (g_force_stages > 0) nb = g_force_stages;
        if (nb > 8) nb = 8;
        // every weight tile of the layer fits next to the patches (one 64-channel chunk: 9 tiles): keep them for all tiles of the
        // worker instead of re-streaming 18 small TMA boxes per tile (stem conv 2 was bound by the TMA operation rate)
        if (g_force_stages <= 0 && p.kchunks == 1 && n_tiles == 1 && (SMEM_BUDGET - a_bytes) / b_stage >= 9) { nb = 9; p.b_resident = 1; }
        if (nb < 2) return fail(FAR3D_E_UNSUPPORTED, "%shalo conv: B ring does not fit (bn %ld)", "", bn);
        p.num_stages = nb;
        smem = a_bytes + nb * b_stage + EP_BYTES + 1024;
        auto mapA = [&](CUtensorMap* tm, const void* base) -> int {
            const cuuint64_t sw = (cuuint64_t)x_cs * 2, sh = (cuuint64_t)W * x_cs * 2, sn = (cuuint64_t)H * W * x_cs * 2;
            cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)Fd, (cuuint64_t)Sd, (cuuint64_t)N};
            cuuint64_t str[3] = {p.transposed ? sh : sw, p.transposed ? sw : sh, sn};
            cuuint32_t box[4] = {(cuuint32_t)UM_BK, (cuuint32_t)HALO_PF, (cuuint32_t)HALO_PS, 1};
            return encode(tm, (const fp16*)base + x_co, 4, dims, str, box);
        };
        if ((rc = mapA(&tmA_hi, x_hi))) return rc;
        if (split && (rc = mapA(&tmA_lo, x_lo))) return rc;
    } else {
        const size_t stage_bytes = (size_t)sp * (UM_A_BYTES + (size_t)(bn / cg) * UM_BK * 2);
        int ns = g_force_stages > 0 ? g_force_stages : (int)(SMEM_BUDGET / stage_bytes);
        if (ns > 8) ns = 8;
        if (ns < 2) return fail(FAR3D_E_UNSUPPORTED, "%sconv: stage ring does not fit (bn %ld)", "", bn);
        p.num_stages = ns; p.a_stages = 0;
        smem = ns * stage_bytes + EP_BYTES + 1024;
        auto mapA = [&](CUtensorMap* tm, const void* base) -> int {
            if (stride == 1) {
                cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
                cuuint64_t str[3] = {(cuuint64_t)x_cs * 2, (cuuint64_t)W * x_cs * 2, (cuuint64_t)H * W * x_cs * 2};
                cuuint32_t box[4] = {(cuuint32_t)UM_BK, (cuuint32_t)p.tw, (cuuint32_t)p.th, 1};
                return encode(tm, (const fp16*)base + x_co, 4, dims, str, box);
            }
            cuuint64_t dims[5] = {(cuuint64_t)2 * x_cs, (cuuint64_t)W / 2, 2, (cuuint64_t)H / 2, (cuuint64_t)N};
            cuuint64_t str[4] = {(cuuint64_t)2 * x_cs * 2, (cuuint64_t)W * x_cs * 2, (cuuint64_t)2 * W * x_cs * 2,
                                 (cuuint64_t)H * W * x_cs * 2};
            cuuint32_t box[5] = {(cuuint32_t)UM_BK, (cuuint32_t)p.tw, 1, (cuuint32_t)p.th, 1};
            return encode(tm, base, 5, dims, str, box);
        };
        if ((rc = mapA(&tmA_hi, x_hi))) return rc;
        if (split && (rc = mapA(&tmA_lo, x_lo))) return rc;
    }
    if ((rc = mapB(&tmB_hi, w_hi, bn / cg))) return rc;
    if (split && (rc = mapB(&tmB_lo, w_lo, bn / cg))) return rc;
    if (!split) { tmA_lo = tmA_hi; tmB_lo = tmB_hi; }
    if (smem > 227 * 1024) return fail(FAR3D_E_UNSUPPORTED, "%sconv smem %ld exceeds 227 KB", "", (long)smem);

    const long total = (long)((p.m_tiles + cg - 1) / cg) * n_tiles;          // work items per persistent worker (CTA or CTA pair)
    int workers = (g_force_grid > 0 ? g_force_grid : sms) / cg;
    if (workers < 1) workers = 1;
    if (workers > total) workers = (int)total;
    const int grid = workers * cg;
#define FAR3D_CONV_LAUNCH(MD, HL)                                                                                        \
    (cg == 2 ? launch(conv_persistent_kernel<MD, HL, 2>, grid, 2, smem, st, tmA_hi, tmA_lo, tmB_hi, tmB_lo, p)            \
             : launch(conv_persistent_kernel<MD, HL, 1>, grid, 1, smem, st, tmA_hi, tmA_lo, tmB_hi, tmB_lo, p))
    if (mx) return halo ? FAR3D_CONV_LAUNCH(2, true) : FAR3D_CONV_LAUNCH(2, false);
    if (split) return halo ? FAR3D_CONV_LAUNCH(1, true) : FAR3D_CONV_LAUNCH(1, false);
    return halo ? FAR3D_CONV_LAUNCH(0, true) : FAR3D_CONV_LAUNCH(0, false);
#undef FAR3D_CONV_LAUNCH
}

extern "C" int far3d_conv2d_umma(const void* x_hi, const void* x_lo, int N, int H, int W, int x_cs, int x_co, int Cin,
                                 const void* w_hi, const void* w_lo, const float* bias, int Cout, int ksize, int stride,
                                 int relu, float* y_f32, int yf_cs, int yf_co, int64_t yf_ns, void* y_hi, void* y_lo,
                                 int yb_cs, int yb_co, void* stream) {
    return conv_impl(x_hi, x_lo, N, H, W, x_cs, x_co, Cin, w_hi, w_lo, bias, Cout, ksize, stride, relu, nullptr, 0, y_f32,
                     yf_cs, yf_co, yf_ns, y_hi, y_lo, yb_cs, yb_co, stream);
}

// fp16mx form: x_c8 / w_c8 / y_c8 are e4m3 correction planes (common.cuh) in place of the fp16 residual planes
extern "C" int far3d_conv2d_umma_mx(const void* x_hi, const void* x_c8, int x_fmt, int N, int H, int W, int x_cs, int x_co, int Cin,
                                    const void* w_hi, const void* w_c8, int w_exp, const float* bias, int Cout, int ksize,
                                    int stride, int relu, float* y_f32, int yf_cs, int yf_co, int64_t yf_ns, void* y_hi,
                                    void* y_lo, int y_fmt, int yb_cs, int yb_co, void* stream) {
    return conv_impl(x_hi, x_c8, N, H, W, x_cs, x_co, Cin, w_hi, w_c8, bias, Cout, ksize, stride, relu, nullptr, 0, y_f32,
                     yf_cs, yf_co, yf_ns, y_hi, y_lo, yb_cs, yb_co, stream, nullptr, nullptr, x_fmt, w_exp, y_fmt);
}

// nn.Linear on the tensor cores: y[M,N] = act(x[M,K] @ w[N,K]^T + bias) (+ residual), operands as split-fp16 planes
// (x_lo / w_lo NULL = plain fp16).  A [rows, K] matrix is a 1 x M "image" with K channels for the implicit-GEMM kernel.
extern "C" int far3d_linear_umma(const void* x_hi, const void* x_lo, int ldx, const void* w_hi, const void* w_lo,
                                 const float* bias, const float* residual, int ldr, float* y, int ldy, int M, int N, int K,
                                 int act, void* stream) {
    FAR3D_REQUIRE(y && M > 0 && N > 0 && K > 0 && ldx >= K && ldy >= N, "bad argument");
    return conv_impl(x_hi, x_lo, 1, 1, M, ldx, 0, K, w_hi, w_lo, bias, N, 1, 1, act, residual, ldr, y, ldy, 0, 0, nullptr,
                     nullptr, 0, 0, stream);
}

// ---------------------------------------------------------------------------------------------- concat conv + eSE pooling
// 1x1 conv (the OSA concat conv, vovnet.py:230-232) that also produces the global average pool of its fp32 output
// (eSEModule, vovnet.py:173-185): every epilogue warp writes the column sums of its 32 in-image pixels to
// `workspace[(m_tile * 4 + warp)][Cout]` (deterministic, no atomics) and a second tiny kernel folds the partials of each
// image into `mean[N][Cout]`.  Saves the separate pass that re-read the whole block output from HBM.
namespace far3d {
// block = 32 columns x 32 partial-row lanes: coalesced reads, 32-way split of the `parts` loop, smem tree over the lanes
__global__ void __launch_bounds__(1024)
colsum_final_kernel(const float* __restrict__ part, float* __restrict__ mean, int N, int parts, int C, float inv_hw) {
    __shared__ float sh[32][33];
    const int n = blockIdx.y, c = blockIdx.x * 32 + threadIdx.x, ty = threadIdx.y;
    float s = 0.f;
    if (c < C) {
        const float* q = part + (size_t)n * parts * C + c;
        for (int t = ty; t < parts; t += 32) s += q[(size_t)t * C];
    }
    sh[ty][threadIdx.x] = s;
    __syncthreads();
    if (ty == 0 && c < C) {
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) a += sh[i][threadIdx.x];
        mean[(size_t)n * C + c] = a * inv_hw;
    }
}
}  // namespace far3d

extern "C" int64_t far3d_conv_pool_workspace_floats(int N, int H, int W, int Cout) {
    // generic 1x1 tiling: th x tw = 128 pixels, fewest tiles (same search as conv_impl)
    long best = -1;
    for (int tw = 8; tw <= 128; tw <<= 1) {
        int th = 128 / tw;
        long t = (long)((W + tw - 1) / tw) * ((H + th - 1) / th);
        if (best < 0 || t < best) best = t;
    }
    return (int64_t)N * best * 4 * Cout;
}

static int conv_pool_impl(const void* x_hi, const void* x_lo, int x_fmt, int N, int H, int W, int x_cs, int x_co, int Cin,
                          const void* w_hi, const void* w_lo, int w_exp, const float* bias, int Cout, int relu,
                          float* y_f32, int yf_cs, int yf_co, float* workspace, float* mean, void* stream) {
    FAR3D_REQUIRE(workspace && mean && y_f32, "null pointer");
    int m_tiles = 0;
    int rc = conv_impl(x_hi, x_lo, N, H, W, x_cs, x_co, Cin, w_hi, w_lo, bias, Cout, 1, 1, relu, nullptr, 0, y_f32, yf_cs,
                       yf_co, 0, nullptr, nullptr, 0, 0, stream, workspace, &m_tiles, x_fmt, w_exp, 0);
    if (rc) return rc;
    const int parts = (m_tiles / N) * 4;                 // tiles never straddle images
    colsum_final_kernel<<<dim3((Cout + 31) / 32, N), dim3(32, 32), 0, (cudaStream_t)stream>>>(workspace, mean, N, parts, Cout,
                                                                                              1.f / (float)((long)H * W));
    return launched("colsum_final_kernel");
}

extern "C" int far3d_conv2d_umma_pool(const void* x_hi, const void* x_lo, int N, int H, int W, int x_cs, int x_co, int Cin,
                                      const void* w_hi, const void* w_lo, const float* bias, int Cout, int relu,
                                      float* y_f32, int yf_cs, int yf_co, float* workspace, float* mean, void* stream) {
    return conv_pool_impl(x_hi, x_lo, 0, N, H, W, x_cs, x_co, Cin, w_hi, w_lo, 0, bias, Cout, relu, y_f32, yf_cs, yf_co, workspace,
                          mean, stream);
}
extern "C" int far3d_conv2d_umma_pool_mx(const void* x_hi, const void* x_c8, int x_fmt, int N, int H, int W, int x_cs, int x_co,
                                         int Cin, const void* w_hi, const void* w_c8, int w_exp, const float* bias, int Cout,
                                         int relu, float* y_f32, int yf_cs, int yf_co, float* workspace, float* mean,
                                         void* stream) {
    return conv_pool_impl(x_hi, x_c8, x_fmt, N, H, W, x_cs, x_co, Cin, w_hi, w_c8, w_exp, bias, Cout, relu, y_f32, yf_cs, yf_co,
                          workspace, mean, stream);
}
