// Persistent, warp-specialised fp16 GEMM for sm_100a on CTA PAIRS (tcgen05 cta_group::2):
//     C[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue)
//
//   * A (activations) and W (nn.Linear weight [out,in]) are K-major fp16 in HBM.
//   * The grid is one 2-CTA cluster per SM pair (TPC), persistent over 256 x BLOCK_N output tiles.
//     Each CTA of the pair TMA-loads ITS 128 rows of A and ITS HALF (BLOCK_N / 2 rows) of W per
//     64-wide k-block (128-byte swizzle) into a shared-memory ring; the leader CTA's elected thread
//     issues tcgen05.mma.cta_group::2 (UMMA 256 x BLOCK_N x 16): both tensor cores read A from their
//     own SM and the two W halves from both SMs, so each SM stages and reads half the W bytes a
//     single-CTA 128 x BLOCK_N tile would need -- the shared-memory bandwidth that capped the 1-CTA
//     kernel (TMA writes + UMMA reads > 128 B/clk/SM) is no longer the limiter.
//   * fp32 accumulators live in TMEM (each CTA holds its 128 rows x BLOCK_N columns), two stages, so
//     the epilogue of tile i overlaps the MMAs of tile i+1.
//   * Epilogue warps drain TMEM with tcgen05.ld, transpose 32 x 32 chunks through XOR-swizzled shared
//     memory and read/write HBM with fully coalesced 128-bit accesses (4 rows x 128 B per warp
//     instruction) while applying bias / quick_gelu / residual / position-embedding fusions.
//
// Barrier protocol (leader = cluster rank 0):
//   full[s]   (leader's)    1 arrival (leader producer, expect_tx = both CTAs' bytes) + TMA bytes of both CTAs
//   empty[s]  (each CTA's)  tcgen05.commit multicast from the leader's MMA thread
//   tmem_full[a]  (each)    tcgen05.commit multicast after the last k-block of a tile
//   tmem_empty[a] (leader's) 16 arrivals: 8 epilogue warps x 2 CTAs (remote arrive from rank 1)
//
// Replaces the cuBLAS SGEMM calls behind nn.Linear in HF CLIP:
//   q/k/v_proj HF:modeling_clip.py:310-312, out_proj :334, fc1/fc2 :348-350, patch conv :148-154.
#pragma once
#include <cuda.h>
#include "gemm_tcgen05.cuh"
#include "ptx.cuh"

namespace mcm {

constexpr int kGemm2TileM = 256;     // rows per cluster tile (128 per CTA)
#ifndef MCM_GEMM_F16_TMA_STORE
#define MCM_GEMM_F16_TMA_STORE 1   // 1: fp16 outputs leave through a shared-memory staging tile + TMA bulk stores (default)
                                   // 0: 32-byte row-per-thread stores straight from registers (A/B builds; measured slower)
#endif
#ifndef MCM_GEMM_F16_CHUNK_STORE
#define MCM_GEMM_F16_CHUNK_STORE 1 // 1 (default): BLOCK_N = 256 fp16 outputs leave as TWO 32-column bulk stores per warp and tile through a
                                   //    2 KB staging tile -- 32 KB of staging instead of 64 KB (one more ring stage), and the first half of
                                   //    a warp's output is on its way while the second is still being computed
                                   // 0: one 64-column store per warp and tile (round 1 / early round 2; A/B builds)
#endif
#ifndef MCM_RESID_BUFS
#define MCM_RESID_BUFS 2
#endif
constexpr int kResidBufs = MCM_RESID_BUFS;   // TMA residual epilogue: residual chunks in flight per warp
constexpr int kResidWarpBytes = kResidBufs * 4096 + 2048;   // + the fp16 staging tile
#ifndef MCM_RESID_H2_BUFS
#define MCM_RESID_H2_BUFS 3
#endif
constexpr int kResidH2Bufs = MCM_RESID_H2_BUFS;   // (hi, lo) TMA residual epilogue: 2 KB + 2 KB chunks per warp, all but one in flight
constexpr int kMaxStatsParts = 8;    // LayerNorm fold: partial row statistics per row (width <= 1024: 2 per 256-column tile)
constexpr int kStgLd = 32;           // fp32 staging row stride in floats; 16-byte chunks are XOR-swizzled by (row & 7)

// What an epilogue kind implies.  fp16-output epilogues are instruction-latency bound (TMEM load ->
// math -> staging transpose -> store is one dependent chain per 32-column chunk), so they run on 16
// warps (four per SM sub-partition, one 64-column slice of the tile each); the fp32 residual epilogues
// are bound by the HBM round trip of the residual rows and run on 8 warps with the NEXT chunk's
// residual already in flight (the register budget of 8 warps allows the double buffer).
template <int EPI>
struct EpiTraits {
    static constexpr bool kLn = (EPI == EPI_LN_F16 || EPI == EPI_LN_QGELU_F16);
    static constexpr bool kGelu = (EPI == EPI_BIAS_QGELU_F16 || EPI == EPI_LN_QGELU_F16);
    static constexpr bool kF16 = (EPI == EPI_BIAS_F16 || EPI == EPI_BIAS_QGELU_F16 || kLn);
    static constexpr bool kResidH2 = (EPI == EPI_BIAS_RESID_H2_LN);          // residual stream as an fp16 (hi, lo) pair, LSU
    static constexpr bool kResid = (EPI == EPI_BIAS_RESID_F32 || EPI == EPI_BIAS_RESID_F32_LN || kResidH2);
    static constexpr bool kTmaResid = (EPI == EPI_BIAS_RESID_F32_LN_TMA);   // residual epilogue through TMA loads / stores
    static constexpr bool kTmaResidH2 = (EPI == EPI_BIAS_RESID_H2_LN_TMA);  // the (hi, lo) pair through TMA loads / stores
    static constexpr bool kStats = (EPI == EPI_BIAS_RESID_F32_LN || kResidH2);
    static constexpr int kWarps = kF16 ? 16 : 8;
    static constexpr int kThreads = 64 + 32 * kWarps;   // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, then epilogue
};

// Shared memory: operand ring + epilogue staging + barriers.  What the epilogue stages decides how many
// ring stages are left of the 227 KB:
//   fp16 outputs          16 warps x (32 rows x 64 B) = 32 KB : one TMA bulk store per warp and 32-column chunk
//                         [MCM_GEMM_F16_CHUNK_STORE=0: 16 warps x (32 rows x 128 B) = 64 KB, one store per warp and tile]
//                         [MCM_GEMM_F16_TMA_STORE=0: none, 32-byte row-per-thread stores straight from registers]
//   fp32, LSU             8 warps x (32 rows x 128 B)  = 32 KB : transpose tile (EPI_BIAS_RESID_F32, EPI_POS_F32, ..._LN)
//   fp32 + fp16, TMA      8 warps x (2 x 4 KB residual in / fp32 out + 2 KB fp16 out) = 80 KB  (EPI_BIAS_RESID_F32_LN_TMA)
template <int BLOCK_N, int EPI>
struct Gemm2Smem {
    using T = EpiTraits<EPI>;
    static constexpr int kABytes = kGemmBlockM * kGemmBlockK * 2;          // 128 x 64 fp16
    static constexpr int kBBytes = (BLOCK_N / 2) * kGemmBlockK * 2;        // this CTA's half of W
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr bool kChunkStore = MCM_GEMM_F16_CHUNK_STORE && T::kF16 && BLOCK_N == 256;
    static constexpr int kStagingBytes = T::kF16 ? (MCM_GEMM_F16_TMA_STORE ? (kChunkStore ? 16 * 2048 : 16 * (BLOCK_N / 4) * 64) : 0)
                                         : T::kTmaResid ? 8 * kResidWarpBytes
                                         : T::kTmaResidH2 ? 8 * kResidH2Bufs * 4096 : 32 * 1024;
    static constexpr int kBarrierBytes = 512;
    static constexpr int kBudget = 232448 - 1024 /* alignment slack */ - kBarrierBytes - kStagingBytes;
    static constexpr int kStages = (kBudget / kStageBytes) < 8 ? (kBudget / kStageBytes) : 8;
    static constexpr int kTotal = kStages * kStageBytes + kStagingBytes + kBarrierBytes + 1024;
    static_assert(kStages >= 3, "too few operand stages");
    static_assert(kTotal <= 232448, "exceeds the 227 KB of shared memory a CTA can opt in to");
    static_assert((2 * kStages + 5 + 8 * (kResidBufs > kResidH2Bufs ? kResidBufs : kResidH2Bufs)) * 8 + 8 <= kBarrierBytes, "barrier area too small");
};

// ---- cluster / 2-CTA PTX ----
// tile index -> (row block, column block).  Tiles are walked row-block-major; with p.m_reverse the row blocks run from the
// LAST to the first, so that a GEMM whose producer walked upwards starts on the rows that are still in L2 (engine.cu alternates
// the direction along the qkv -> attention -> out_proj -> fc1 -> fc2 chain).
__device__ __forceinline__ void gemm2_tile_pos(const GemmParams& p, int tile, int& m_blk, int& n_blk) {
    m_blk = tile / p.n_tiles;
    n_blk = tile - m_blk * p.n_tiles;
    if (p.m_reverse) m_blk = p.m_tiles - 1 - m_blk;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
// Arrive on a barrier of any CTA of the cluster (shared::cluster address).  Default semantics (release at CTA scope): what
// it orders here is a drained TENSOR-MEMORY stage (tcgen05.wait::ld + tcgen05.fence::before_thread_sync in front of it), never
// generic-proxy memory.  Round 1 used `.release.cluster`, which ptxas lowers to MEMBAR.ALL.GPU + ERRBAR in front of the
// arrive: ncu's source view put 23-26 % of ALL warp-stall samples of the q/k/v and fc1 GEMMs (16 epilogue warps, one release
// per warp and tile) on those three instructions (profiles/r02_ncu_gemm_vit_b16.txt).
#ifndef MCM_ARRIVE_RELEASE_CLUSTER
#define MCM_ARRIVE_RELEASE_CLUSTER 0
#endif
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
#if MCM_ARRIVE_RELEASE_CLUSTER
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
#else
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
#endif
}
// 2D tiled load into THIS CTA's smem; the byte count is signalled on `bar_cluster_addr`, a
// shared::cluster mbarrier address that may live in the peer (leader) CTA.
__device__ __forceinline__ void tma_load_2d_cta2(void* smem_dst, const void* tmap, uint32_t bar_cluster_addr, int32_t c0,
                                                 int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
// Same, multicast: the box lands at the same smem offset in every CTA of `cta_mask`; the byte count is
// signalled, per destination CTA, on the barrier at `bar`'s offset in that CTA's PAIR LEADER (the
// barrier address carries an even CTA rank) -- the form CUTLASS' SM100_TMA_2SM_LOAD_MULTICAST uses.
__device__ __forceinline__ void tma_load_2d_cta2_mc(void* smem_dst, const void* tmap, uint32_t bar_cluster_addr, int32_t c0,
                                                    int32_t c1, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster_addr), "r"(c0), "r"(c1),
        "h"(cta_mask)
        : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_cta2(uint32_t* dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_cta2(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void umma_f16_cta2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (once all previously issued MMAs retired) on the barrier at this offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_cta2_mc(uint64_t* bar, uint16_t cta_mask) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"(cta_mask)
        : "memory");
}

// ---- fp32 epilogues (residual add, position add): transposed through shared memory ----
// A warp's 32-row x 32-column chunk of the accumulator leaves through an XOR-swizzled smem transpose so
// that global accesses are 4 rows x 128 B per warp instruction.  Lane layout after the transpose:
// cq = lane & 7 is the 4-column group, rows r0 + 4 i (r0 = lane >> 3, i = 0..7).
//
// The residual / position rows of a chunk ("side" values) are loaded by gemm2_load_side one chunk AHEAD
// of their use (across tiles too), so the HBM round trip overlaps the TMEM drain of the previous chunk.
template <int EPI>
__device__ __forceinline__ void gemm2_load_side(const GemmParams& p, int m_base, int col0, int lane, float4 (&side)[8]) {
    const int col = col0 + 4 * (lane & 7);
    const int r0 = lane >> 3;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m_base + r0 + 4 * i;
        side[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if constexpr (EpiTraits<EPI>::kResidH2) {
            if (m < p.m_valid) {
                const size_t off = static_cast<size_t>(m) * p.ldo + col;
                // RAW bits only: converting here would wait for the loads at the point of the prefetch (measured: the
                // epilogue then stalls a full HBM round trip per chunk); gemm2_epilogue_chunk decodes them a chunk later
                const uint2 hh = *reinterpret_cast<const uint2*>(p.resid16 + off), ll = *reinterpret_cast<const uint2*>(p.resid16_lo + off);
                side[i] = make_float4(__uint_as_float(hh.x), __uint_as_float(hh.y), __uint_as_float(ll.x), __uint_as_float(ll.y));
            }
        } else if constexpr (EpiTraits<EPI>::kResid) {
            if (m < p.m_valid && !MCM_DBG_SKIP(p, 8)) side[i] = *reinterpret_cast<const float4*>(p.resid + static_cast<size_t>(m) * p.ldo + col);
        } else {   // EPI_POS_F32: patch row m of image b carries position 1 + patch
            const int pi = m % p.np;
            if (m < p.m_valid) side[i] = __ldg(reinterpret_cast<const float4*>(p.pos + static_cast<size_t>(1 + pi) * p.ldo + col));
        }
    }
}

// t_addr: TMEM address (lane quadrant + column) of the chunk; m_base: global row of this warp's first
// row; col0: first global column of the chunk.  s1 / s2 accumulate, per lane, the sum and the sum of
// squares of the values written to rows r0 + 4 i (EPI_BIAS_RESID_F32_LN only).
template <int EPI, bool SPLIT>
__device__ __forceinline__ void gemm2_epilogue_chunk(const GemmParams& p, uint32_t stg, uint32_t t_addr, int m_base, int col0,
                                                     int lane, const float4 bias, const float4 (&side)[8], float (&s1)[8],
                                                     float (&s2)[8]) {
    const int cq = lane & 7;
    const int r0 = lane >> 3;
    const int col = col0 + 4 * cq;
    if (MCM_DBG_SKIP(p, 4)) return;
    uint32_t acc[32];
    tmem_ld_32x32b_x32(t_addr, acc);
    tmem_ld_wait();
    const uint32_t srow = stg + lane * (kStgLd * 4);
#pragma unroll
    for (int j = 0; j < 8; ++j)   // row `lane`, 16-byte chunk j -> slot j ^ (lane & 7): conflict-free both ways
        sts_v4(srow + ((j ^ (lane & 7)) << 4), make_float4(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1]),
                                                           __uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3])));
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = r0 + 4 * i;
        const int m = m_base + r;
        float4 v = lds_v4(stg + r * (kStgLd * 4) + ((cq ^ (r & 7)) << 4));
        if (m < p.m_valid && (!MCM_DBG_SKIP(p, 1) || v.x == 123.456f)) {
            if constexpr (EpiTraits<EPI>::kResid) {
                float4 sd = side[i];
                if constexpr (EpiTraits<EPI>::kResidH2) {     // (hi pair 0, hi pair 1, lo pair 0, lo pair 1) raw bits -> hi + lo, exact in fp32
                    const float2 h0 = unpack_op16x2(__float_as_uint(sd.x)), h1 = unpack_op16x2(__float_as_uint(sd.y));
                    const float2 l0 = unpack_op16x2(__float_as_uint(sd.z)), l1 = unpack_op16x2(__float_as_uint(sd.w));
                    sd = make_float4(h0.x + l0.x, h0.y + l0.y, h1.x + l1.x, h1.y + l1.y);
                }
                v.x = sd.x + (v.x + bias.x); v.y = sd.y + (v.y + bias.y);
                v.z = sd.z + (v.z + bias.z); v.w = sd.w + (v.w + bias.w);
                const size_t off = static_cast<size_t>(m) * p.ldo + col;
                if constexpr (!EpiTraits<EPI>::kResidH2) *reinterpret_cast<float4*>(static_cast<float*>(p.out) + off) = v;
                if constexpr (EpiTraits<EPI>::kStats) {
                    if constexpr (SPLIT || EpiTraits<EPI>::kResidH2) {
                        uint32_t h0, h1, l0, l1;
                        split_op16x2(v.x, v.y, h0, l0);
                        split_op16x2(v.z, v.w, h1, l1);
                        *reinterpret_cast<uint2*>(p.out16 + off) = make_uint2(h0, h1);
                        *reinterpret_cast<uint2*>(p.out16_lo + off) = make_uint2(l0, l1);
                    } else {
                        *reinterpret_cast<uint2*>(p.out16 + off) = make_uint2(pack_op16x2(v.x, v.y), pack_op16x2(v.z, v.w));
                    }
                    s1[i] += (v.x + v.y) + (v.z + v.w);
                    s2[i] += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
                }
            } else {  // EPI_POS_F32: patch row m of image b -> token row b * seq + 1 + patch
                const int b = m / p.np;
                const size_t orow = static_cast<size_t>(b) * p.seq + 1 + (m - b * p.np);
                v.x += side[i].x; v.y += side[i].y; v.z += side[i].z; v.w += side[i].w;
                *reinterpret_cast<float4*>(static_cast<float*>(p.out) + orow * p.ldo + col) = v;
            }
        }
    }
    __syncwarp();
}

// ---- fp16-output epilogues (bias, LayerNorm fold, quick_gelu) ----
// The math runs on the TMEM registers (row per thread, 32 independent columns -> deep MUFU pipelining),
// only the packed fp16 result (64 B per row) goes through the staging transpose, and HBM sees
// 8 rows x 64 B per warp instruction.  rstd / nmr (= -mu * rstd) are this thread's row statistics.
// 2D tiled store shared -> global (bulk async group of the issuing thread); rows / columns outside the
// tensor are clipped by the TMA unit
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t smem_src, int32_t c0, int32_t c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the staging tile of every committed store of this thread has been read (it may be overwritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// TMEM chunk (32 rows x 32 columns) -> bias / LayerNorm fold / quick_gelu -> 16 packed fp16 pairs of this thread's row.
// release_bar != 0 (last chunk of the tile): arrive on that cluster-space tmem_empty barrier as soon as the
// accumulator values sit in registers, so the next-but-one tile's MMAs need not wait for the math and the stores.
template <int EPI, bool SPLIT>
__device__ __forceinline__ void gemm2_f16_chunk_math(const GemmParams& p, uint32_t t_addr, int col0, int lane, float rstd, float nmr,
                                                     uint32_t release_bar, uint32_t (&pk)[16], uint32_t (&pkl)[SPLIT ? 16 : 1]) {
    using T = EpiTraits<EPI>;
    static_assert(T::kF16, "fp16-output epilogues only");
    uint32_t acc[32];
    tmem_ld_32x32b_x32(t_addr, acc);
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);   // warp-uniform addresses: broadcast loads
    const float4* c4 = reinterpret_cast<const float4*>(p.colsum + col0);
#pragma unroll
    for (int h = 0; h < 2; ++h) {   // 16 columns at a time keeps the column vectors at 32 registers
        float4 bias[4], cs[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            bias[j] = __ldg(b4 + 4 * h + j);
            if constexpr (T::kLn) cs[j] = __ldg(c4 + 4 * h + j);
        }
        if (h == 0) {
            tmem_ld_wait();
            if (release_bar) {
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(release_bar);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int a = 16 * h + 4 * j;
            float v0, v1, v2, v3;
            if constexpr (T::kLn) {
                v0 = fmaf(rstd, __uint_as_float(acc[a + 0]), fmaf(nmr, cs[j].x, bias[j].x));
                v1 = fmaf(rstd, __uint_as_float(acc[a + 1]), fmaf(nmr, cs[j].y, bias[j].y));
                v2 = fmaf(rstd, __uint_as_float(acc[a + 2]), fmaf(nmr, cs[j].z, bias[j].z));
                v3 = fmaf(rstd, __uint_as_float(acc[a + 3]), fmaf(nmr, cs[j].w, bias[j].w));
            } else {
                v0 = __uint_as_float(acc[a + 0]) + bias[j].x;
                v1 = __uint_as_float(acc[a + 1]) + bias[j].y;
                v2 = __uint_as_float(acc[a + 2]) + bias[j].z;
                v3 = __uint_as_float(acc[a + 3]) + bias[j].w;
            }
            if constexpr (T::kGelu) {
                if constexpr (SPLIT) {
                    v0 = quick_gelu_precise(v0); v1 = quick_gelu_precise(v1); v2 = quick_gelu_precise(v2); v3 = quick_gelu_precise(v3);
                } else {
                    v0 = quick_gelu(v0); v1 = quick_gelu(v1); v2 = quick_gelu(v2); v3 = quick_gelu(v3);
                }
            }
            if constexpr (SPLIT) {
                split_op16x2(v0, v1, pk[8 * h + 2 * j + 0], pkl[8 * h + 2 * j + 0]);
                split_op16x2(v2, v3, pk[8 * h + 2 * j + 1], pkl[8 * h + 2 * j + 1]);
            } else {
                pk[8 * h + 2 * j + 0] = pack_op16x2(v0, v1);
                pk[8 * h + 2 * j + 1] = pack_op16x2(v2, v3);
            }
        }
    }
}

// 256-bit global store (one full 32-byte sector per thread)
__device__ __forceinline__ void stg_256(void* gptr, const uint32_t* r) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(gptr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

// One warp's slice of a tile (32 rows x BLOCK_N / 4 columns) WITHOUT shared memory or the TMA unit
// (MCM_GEMM_F16_TMA_STORE=0 builds): math per 32-column chunk on the row this thread owns, then two 32-byte stores
// of the packed row straight from registers (every store instruction writes 32 whole sectors).
// Why it exists, and why it is not the default (clock64 trace build, K = 768 projections, cycles per 256 x 256 tile):
//   no epilogue at all 6.16 k (= the UMMA rate) | staging writes only 6.17 k | staging + TMA bulk stores 7.4 k
//   (the issuer waits for operand data: the bulk stores queue in the TMA unit in front of the operand loads, which
//   alone already move ~125 B/clk/SM) | direct stores 8.0 k: the operand loads are no longer delayed (data waits
//   drop from 4.0 k to 1.7 k) but the LSU retires about one 32-byte sector per clock, the 16 warps' stores pile up
//   behind each other and the epilogue (7.4 k busy) becomes longer than the main loop.
// A third variant -- staging tiles drained by two dedicated store warps with coalesced 128-bit LSU stores, so that
// the math warps never wait on a store -- measured 7.7-8.5 k: two warps cannot keep 64 KB per tile moving through
// the LSU, four more warps do not fit the register file next to 16 epilogue warps.  Bulk stores stay the default.
template <int EPI, int BLOCK_N>
__device__ __forceinline__ void gemm2_epilogue_slice_f16_direct(const GemmParams& p, uint32_t t_addr, int m_base, int col0, int lane,
                                                                float rstd, float nmr, uint32_t release_bar) {
    constexpr int kChunks = BLOCK_N / 128;   // 32-column chunks per slice
    if (MCM_DBG_SKIP(p, 4)) {
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(release_bar);
        return;
    }
    const int m = m_base + lane;
    op16_t* orow = static_cast<op16_t*>(p.out) + static_cast<size_t>(m) * p.ldo + col0;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
        uint32_t pk[16], pkl[1];
        gemm2_f16_chunk_math<EPI, false>(p, t_addr + c * 32, col0 + c * 32, lane, rstd, nmr, c + 1 == kChunks ? release_bar : 0u, pk, pkl);
        if (m < p.m_valid && !MCM_DBG_SKIP(p, 1)) {
            stg_256(orow + c * 32, pk);
            stg_256(orow + c * 32 + 16, pk + 8);
        }
    }
}

// The same slice through a staging tile + TMA (MCM_GEMM_F16_TMA_STORE builds): the packed rows are
// written into the warp's staging tile in the TMA swizzle layout of the output box (64-byte rows / SWIZZLE_64B for
// 32-column boxes -- conflict-free for a row-per-thread writer) and leave as one bulk store per 32-column chunk; rows
// beyond the tensor are clipped by the TMA unit.  (MCM_GEMM_F16_CHUNK_STORE=0: 128-byte rows / SWIZZLE_128B, ONE
// 64-column store per warp and tile.)  Measured, round 2, clock64 trace builds, K = 768 projections at batch 512, cycles
// the MMA issuer needs per 256 x 256 tile: no global stores 6.16 k | one store per tile 7.49 k | one per chunk 6.57 k
// (q/k/v), 6.81 k (fc1: the quick_gelu epilogue itself is the longer side now).  In TIME the gain is 3-5 %, not 12 %:
// the GPU sits at its power cap during these kernels (SM clock ~1.0-1.1 GHz), and cycles the tensor pipe no longer
// idles are paid for with clock (profiles/r02_gemm_ring_experiments.txt).
template <int EPI, int BLOCK_N, bool SPLIT>
__device__ __forceinline__ void gemm2_epilogue_slice_f16(const GemmParams& p, const CUtensorMap* tmap_out, uint32_t stg,
                                                         uint32_t t_addr, int m_base, int col0, int lane, float rstd, float nmr,
                                                         uint32_t release_bar) {
    constexpr int kChunks = BLOCK_N / 128;   // 32-column chunks per slice
    if (MCM_DBG_SKIP(p, 4)) {
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(release_bar);
        return;
    }
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
        uint32_t pk[16], pkl[SPLIT ? 16 : 1];
        gemm2_f16_chunk_math<EPI, SPLIT>(p, t_addr + c * 32, col0 + c * 32, lane, rstd, nmr, c + 1 == kChunks ? release_bar : 0u, pk, pkl);
        if constexpr (SPLIT) {   // low halves: two 32-byte stores of this thread's row (the k-loop is 3x longer: off the critical path)
            const int m = m_base + lane;
            if (m < p.m_valid) {
                op16_t* orow = p.out_lo + static_cast<size_t>(m) * p.ldo + col0 + c * 32;
                stg_256(orow, pkl);
                stg_256(orow + 16, pkl + 8);
            }
        }
        if constexpr (kChunks == 2 && MCM_GEMM_F16_CHUNK_STORE) {
            // one 32-column chunk at a time through a 2 KB tile (64-byte rows, SWIZZLE_64B); the previous chunk's store has had
            // this chunk's math to read the tile
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
            const uint32_t srow = stg + lane * 64;
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4)
                sts_v4u(srow + ((s4 ^ ((lane >> 1) & 3)) << 4), make_uint4(pk[4 * s4], pk[4 * s4 + 1], pk[4 * s4 + 2], pk[4 * s4 + 3]));
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && !MCM_DBG_SKIP(p, 1)) {
                tma_store_2d(tmap_out, stg, col0 + c * 32, m_base);
                tma_store_commit();
            }
            continue;
        }
        if (c == 0) {   // the previous tile's store must have read the staging tile (issued a whole main loop ago)
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
        }
        if constexpr (kChunks == 2) {
            const uint32_t srow = stg + lane * 128;
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4)
                sts_v4u(srow + (((4 * c + s4) ^ (lane & 7)) << 4), make_uint4(pk[4 * s4], pk[4 * s4 + 1], pk[4 * s4 + 2], pk[4 * s4 + 3]));
        } else {
            const uint32_t srow = stg + lane * 64;
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4)
                sts_v4u(srow + ((s4 ^ ((lane >> 1) & 3)) << 4), make_uint4(pk[4 * s4], pk[4 * s4 + 1], pk[4 * s4 + 2], pk[4 * s4 + 3]));
        }
    }
    if constexpr (kChunks == 2 && MCM_GEMM_F16_CHUNK_STORE) return;   // every chunk went out on its own
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0 && !MCM_DBG_SKIP(p, 1)) {
        tma_store_2d(tmap_out, stg, col0, m_base);
        tma_store_commit();
    }
}

// ---- residual epilogue through TMA (EPI_BIAS_RESID_F32_LN_TMA) ----
// Same result as EPI_BIAS_RESID_F32_LN, but no byte of the epilogue goes through the LSU (measured at about
// one 32-byte sector per clock and SM, which made the LSU version of the out_proj epilogue 3x longer than its
// main loop).  Per warp and 32 x 32 chunk: the residual chunk arrives by TMA in a SWIZZLE_128B tile (loads run
// kResidBufs - 1 chunks ahead, across tiles), each thread adds its accumulator row IN PLACE (row per thread,
// 16-byte slots XOR-swizzled by row & 7: conflict-free), accumulates its row's sum / sum of squares in
// registers (no shuffles), writes the packed fp16 row into a SWIZZLE_64B tile, and two bulk stores write
// the fp32 chunk back to x and the fp16 chunk to xh.  The fp32 tensor map (tmap_out) serves load and store.
__device__ __forceinline__ void tma_load_2d_plain(uint32_t smem_dst, const void* tmap, uint32_t bar, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}

// p.m_tiles counts 256-row pair tiles.  Launched with cudaLaunchKernelEx + cluster dimension 2.
// SPLIT: split-fp16 precision mode (gemm_tcgen05.cuh): tmap_a_lo / tmap_b_lo are the low halves of A / W.
template <int BLOCK_N, int EPI, bool SPLIT>
__global__ void __launch_bounds__(EpiTraits<EPI>::kThreads, 1)
gemm_f16_tn_cta2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                        const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_out16,
                        const __grid_constant__ CUtensorMap tmap_a_lo, const __grid_constant__ CUtensorMap tmap_b_lo,
                        const GemmParams p) {
    static_assert(!(SPLIT && (EpiTraits<EPI>::kTmaResid || EpiTraits<EPI>::kTmaResidH2)), "the split mode uses the LSU residual epilogues");
    using L = Gemm2Smem<BLOCK_N, EPI>;
    using T = EpiTraits<EPI>;
    constexpr int kStages = L::kStages;
    constexpr uint32_t kTmemCols = 2 * BLOCK_N;  // two accumulator stages
    static_assert(BLOCK_N == 128 || BLOCK_N == 256, "BLOCK_N");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* staging = smem + kStages * L::kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * L::kStageBytes + L::kStagingBytes);
    uint64_t* full_bar = bars;                      // [kStages]
    uint64_t* empty_bar = bars + kStages;           // [kStages]
    uint64_t* tmem_full = bars + 2 * kStages;       // [2]
    uint64_t* tmem_empty = bars + 2 * kStages + 2;  // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
    uint64_t* tmem_ptr_bars = bars + 2 * kStages + 5;   // [8 warps][kResidBufs]  (EPI_BIAS_RESID_F32_LN_TMA only)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();           // rank inside the CTA pair (0 = leader)
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;
    const int num_tiles = p.m_tiles * p.n_tiles;
    const int k_total = SPLIT ? 3 * p.k_blocks : p.k_blocks;   // SPLIT: (hi, hi), (lo, hi), (hi, lo) per k-block

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        if constexpr (SPLIT) {
            tma_prefetch_desc(&tmap_a_lo);
            tma_prefetch_desc(&tmap_b_lo);
        }
        if constexpr ((T::kF16 && MCM_GEMM_F16_TMA_STORE) || T::kTmaResid || T::kTmaResidH2) tma_prefetch_desc(&tmap_out);
        if constexpr (T::kTmaResid || T::kTmaResidH2) tma_prefetch_desc(&tmap_out16);
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 2 * T::kWarps);
        }
        if constexpr (T::kTmaResid)
            for (int i = 0; i < 8 * kResidBufs; ++i) mbar_init(&tmem_ptr_bars[i], 1);
        if constexpr (T::kTmaResidH2)
            for (int i = 0; i < 8 * kResidH2Bufs; ++i) mbar_init(&tmem_ptr_bars[i], 1);
        fence_barrier_init();
    }
    pdl_launch_dependents();
    if (warp == 1) tmem_alloc_cta2<kTmemCols>(tmem_ptr);
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();   // everything above overlapped the predecessor's tail; global data is touched only below

    if (warp == 0) {
        if (elect_one()) {
            // ===== TMA producer (both CTAs): own A rows, own half of W; bytes land on the leader's full barrier =====
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                int m_blk, n_blk;
                gemm2_tile_pos(p, tile, m_blk, n_blk);
                const int a_row = m_blk * kGemm2TileM + static_cast<int>(rank) * kGemmBlockM;
                const int b_row = n_blk * BLOCK_N + static_cast<int>(rank) * (BLOCK_N / 2);
                for (int kt = 0; kt < k_total; ++kt) {
                    int kb = kt;
                    const CUtensorMap* ta = &tmap_a;
                    const CUtensorMap* tb = &tmap_b;
                    if constexpr (SPLIT) {
                        kb = kt / 3;
                        const int seg = kt - 3 * kb;
                        if (seg == 1) ta = &tmap_a_lo;
                        if (seg == 2) tb = &tmap_b_lo;
                    }
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * L::kStageBytes;
                    uint8_t* sb = sa + L::kABytes;
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
                    const uint32_t bar = mapa_shared(smem_u32(&full_bar[stage]), 0);
                    tma_load_2d_cta2(sa, ta, bar, kb * kGemmBlockK, a_row);
                    tma_load_2d_cta2(sb, tb, bar, kb * kGemmBlockK, b_row);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0 && elect_one()) {
            // ===== MMA issuer (pair leader only) =====
            constexpr uint32_t idesc = make_idesc_f16(kGemm2TileM, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
#ifdef MCM_GEMM_TRACE
            long long t_acc = 0, t_full = 0, t_all = clock64(), n_t = 0;
#define MCM_TR(var, stmt) { const long long t0__ = clock64(); stmt; var += clock64() - t0__; }
#else
#define MCM_TR(var, stmt) { stmt; }
#endif
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                MCM_TR(t_acc, mbar_wait(&tmem_empty[as], aphase ^ 1));
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + as * BLOCK_N;
                for (int kb = 0; kb < k_total; ++kb) {
                    MCM_TR(t_full, mbar_wait(&full_bar[stage], phase));
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
                    const uint64_t adesc = make_smem_desc_sw128(sa, 16, 1024);
                    const uint64_t bdesc = make_smem_desc_sw128(sa + L::kABytes, 16, 1024);
#pragma unroll
                    for (int k = 0; k < kGemmBlockK / 16; ++k)
                        umma_f16_cta2(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                    umma_commit_cta2_mc(&empty_bar[stage], 3);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                umma_commit_cta2_mc(&tmem_full[as], 3);
                if (++as == 2) { as = 0; aphase ^= 1; }
#ifdef MCM_GEMM_TRACE
                ++n_t;
#endif
            }
#ifdef MCM_GEMM_TRACE
            if (p.trace) {   // MMA issuer: cycles blocked on a free accumulator / on operand data, total, tiles
                atomicAdd(reinterpret_cast<unsigned long long*>(p.trace + 0), (unsigned long long)t_acc);
                atomicAdd(reinterpret_cast<unsigned long long*>(p.trace + 1), (unsigned long long)t_full);
                atomicAdd(reinterpret_cast<unsigned long long*>(p.trace + 2), (unsigned long long)(clock64() - t_all));
                atomicAdd(reinterpret_cast<unsigned long long*>(p.trace + 3), (unsigned long long)n_t);
            }
#endif
        }
    } else {
        // ===== epilogue warps (both CTAs): TMEM -> registers -> smem transpose -> coalesced HBM =====
        const int ew = warp - 2;
        const int quad = warp & 3;          // TMEM lane quadrant this warp may access
        int as = 0;
        uint32_t aphase = 0;
        if constexpr (T::kF16) {
            const int slice = ew >> 2;                  // which quarter of the tile's columns this warp drains
            constexpr int kSliceCols = BLOCK_N / 4;
#if MCM_GEMM_F16_TMA_STORE
            const uint32_t stg = smem_u32(staging + ew * (L::kChunkStore ? 2048 : kSliceCols * 64));
#endif
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                int m_blk, n_blk;
                gemm2_tile_pos(p, tile, m_blk, n_blk);
                const int m_base = m_blk * kGemm2TileM + static_cast<int>(rank) * kGemmBlockM + quad * 32;
                const int col_base = n_blk * BLOCK_N + slice * kSliceCols;
                float rstd = 1.f, nmr = 0.f;
                if constexpr (T::kLn) {
                    // row statistics from the partial sums the producing epilogue (or embed_finish) left
                    // (all loads are issued before the first use: a load-add loop would serialise the L2 round trips)
                    const int m = m_base + lane;
                    float2 t[kMaxStatsParts];
#pragma unroll
                    for (int q = 0; q < kMaxStatsParts; ++q) {
                        t[q] = make_float2(0.f, 0.f);
                        if (q < p.stats_parts && m < p.m_valid) t[q] = __ldg(p.stats_in + static_cast<size_t>(q) * p.stats_ld + m);
                    }
                    float s1 = 0.f, s2 = 0.f;
#pragma unroll
                    for (int q = 0; q < kMaxStatsParts; ++q) {
                        s1 += t[q].x;
                        s2 += t[q].y;
                    }
                    const float mu = s1 * p.inv_k;
                    rstd = rsqrtf(fmaxf(s2 * p.inv_k - mu * mu, 0.f) + p.eps);
                    nmr = -mu * rstd;
                }
#ifdef MCM_GEMM_TRACE
                const long long tw0 = clock64();
#endif
                mbar_wait(&tmem_full[as], aphase);
#ifdef MCM_GEMM_TRACE
                const long long tw1 = clock64();
#endif
                tcgen05_fence_after();
                const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BLOCK_N + slice * kSliceCols;
                const uint32_t release_bar = mapa_shared(smem_u32(&tmem_empty[as]), 0);   // the pair leader's barrier
                if (m_base < p.m_valid) {
#if MCM_GEMM_F16_TMA_STORE
                    gemm2_epilogue_slice_f16<EPI, BLOCK_N, SPLIT>(p, &tmap_out, stg, t_row, m_base, col_base, lane, rstd, nmr, release_bar);
#else
                    gemm2_epilogue_slice_f16_direct<EPI, BLOCK_N>(p, t_row, m_base, col_base, lane, rstd, nmr, release_bar);
#endif
                } else {
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(release_bar);
                }
#ifdef MCM_GEMM_TRACE
                if (p.trace && warp == 2 && lane == 0 && rank == 0) {   // one epilogue warp per cluster: wait / busy cycles
                    atomicAdd(reinterpret_cast<unsigned long long*>(p.trace + 4), (unsigned long long)(tw1 - tw0));
                    atomicAdd(reinterpret_cast<unsigned long long*>(p.trace + 5), (unsigned long long)(clock64() - tw1));
                }
#endif
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        } else if constexpr (T::kTmaResid) {
            const int half = ew >> 2;               // which half of the tile's columns this warp drains
            constexpr int kChunks = BLOCK_N / 64;   // 32-column chunks per warp and tile
            const uint32_t wbase = smem_u32(staging + ew * kResidWarpBytes);   // [kResidBufs][4 KB] residual / fp32 out, then 2 KB fp16 out
            const uint32_t hbuf = wbase + kResidBufs * 4096;
            uint64_t* rbar = tmem_ptr_bars + ew * kResidBufs;                  // this warp's "residual chunk landed" barriers
            const int my_tiles = (num_tiles - cluster_id + num_clusters - 1) / num_clusters;
            const int n_chunks = my_tiles * kChunks;
            // chunk g of this warp -> (global row of the warp's first row, first column)
            auto chunk_pos = [&](int g, int& m0, int& c0) {
                const int tile = cluster_id + (g / kChunks) * num_clusters;
                int m_blk, n_blk;
                gemm2_tile_pos(p, tile, m_blk, n_blk);
                m0 = m_blk * kGemm2TileM + static_cast<int>(rank) * kGemmBlockM + quad * 32;
                c0 = n_blk * BLOCK_N + half * (BLOCK_N / 2) + (g % kChunks) * 32;
            };
            auto issue_load = [&](int g) {   // lane 0 only
                int m0, c0;
                chunk_pos(g, m0, c0);
                const uint32_t bar = smem_u32(&rbar[g % kResidBufs]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(4096) : "memory");
                tma_load_2d_plain(wbase + (g % kResidBufs) * 4096, &tmap_out, bar, c0, m0);
            };
            // (an HBM -> L2 prefetch of the residual one tile ahead was measured slower: the chunk time is set by the
            // queueing of this warp's bulk copies behind the operand loads in the SM's TMA unit, not by HBM latency)
            if (lane == 0)
                for (int g = 0; g < kResidBufs - 1 && g < n_chunks; ++g) issue_load(g);
            float s1 = 0.f, s2 = 0.f;
            for (int g = 0; g < n_chunks; ++g) {
                const int c = g % kChunks;
                int m_base, col0;
                chunk_pos(g, m_base, col0);
                if (c == 0) {
#ifdef MCM_GEMM_TRACE
                    const long long tw0 = clock64();
#endif
                    mbar_wait(&tmem_full[as], aphase);
#ifdef MCM_GEMM_TRACE
                    if (p.trace && warp == 2 && lane == 0 && rank == 0)
                        atomicAdd(reinterpret_cast<unsigned long long*>(p.trace + 4), (unsigned long long)(clock64() - tw0));
#endif
                    tcgen05_fence_after();
                }
#ifdef MCM_GEMM_TRACE
                const long long tb0 = clock64();
#endif
                uint32_t acc[32];
                tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BLOCK_N + half * (BLOCK_N / 2) + c * 32, acc);
                const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);   // warp-uniform: broadcast loads
                // every bulk store committed so far has read its smem source: the fp16 tile and the buffer of chunk g - 1
                // are free; start the load that runs kResidBufs - 1 chunks ahead into the latter
                if (lane == 0) {
                    tma_store_wait_read();
                    if (g + kResidBufs - 1 < n_chunks) issue_load(g + kResidBufs - 1);
                }
                tmem_ld_wait();
                if (c + 1 == kChunks) {   // accumulator stage drained into registers: hand it back to the MMA warp
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[as]), 0));
                    if (++as == 2) { as = 0; aphase ^= 1; }
                }
                mbar_wait(&rbar[g % kResidBufs], (g / kResidBufs) & 1);
                const uint32_t rrow = wbase + (g % kResidBufs) * 4096 + lane * 128;
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t a = rrow + ((j ^ (lane & 7)) << 4);
                    float4 v = lds_v4(a);
                    const float4 b = __ldg(b4 + j);
                    v.x += __uint_as_float(acc[4 * j + 0]) + b.x;
                    v.y += __uint_as_float(acc[4 * j + 1]) + b.y;
                    v.z += __uint_as_float(acc[4 * j + 2]) + b.z;
                    v.w += __uint_as_float(acc[4 * j + 3]) + b.w;
                    sts_v4(a, v);
                    s1 += (v.x + v.y) + (v.z + v.w);
                    s2 += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
                    pk[2 * j + 0] = pack_op16x2(v.x, v.y);
                    pk[2 * j + 1] = pack_op16x2(v.z, v.w);
                }
                __syncwarp();   // lane 0's tma_store_wait_read above covers the fp16 tile for every lane from here on
                {
                    const uint32_t hrow = hbuf + lane * 64;
#pragma unroll
                    for (int s4 = 0; s4 < 4; ++s4)
                        sts_v4u(hrow + ((s4 ^ ((lane >> 1) & 3)) << 4), make_uint4(pk[4 * s4], pk[4 * s4 + 1], pk[4 * s4 + 2], pk[4 * s4 + 3]));
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tmap_out, wbase + (g % kResidBufs) * 4096, col0, m_base);
                    tma_store_2d(&tmap_out16, hbuf, col0, m_base);
                    tma_store_commit();
                }
                if (c + 1 == kChunks) {   // this thread's row is complete for this half tile
                    const int m = m_base + lane;
                    const int n_blk = (col0 - half * (BLOCK_N / 2)) / BLOCK_N;
                    if (m < p.m_valid) p.stats_out[static_cast<size_t>(n_blk * 2 + half) * p.stats_ld + m] = make_float2(s1, s2);
                    s1 = s2 = 0.f;
                }
#ifdef MCM_GEMM_TRACE
                if (p.trace && warp == 2 && lane == 0 && rank == 0)
                    atomicAdd(reinterpret_cast<unsigned long long*>(p.trace + 5), (unsigned long long)(clock64() - tb0));
#endif
            }
        } else if constexpr (T::kTmaResidH2) {
            // ---- residual pair through TMA: per warp and 32 x 32 chunk the (hi, lo) fp16 tiles of the residual arrive by TMA
            // (SWIZZLE_64B, 2 KB each, loads run kResidH2Bufs - 1 chunks ahead, across tiles); each thread owns one row: it adds
            // its accumulator row + bias, re-splits the sum IN PLACE, accumulates the row's sum / sum of squares, and two bulk
            // stores write the pair back.  tmap_out is the hi array (the next projection's A operand), tmap_out16 the lo array.
            const int half = ew >> 2;               // which half of the tile's columns this warp drains
            constexpr int kChunks = BLOCK_N / 64;   // 32-column chunks per warp and tile
            const uint32_t wbase = smem_u32(staging + ew * (kResidH2Bufs * 4096));   // [kResidH2Bufs][hi 2 KB | lo 2 KB]
            uint64_t* rbar = tmem_ptr_bars + ew * kResidH2Bufs;
            const int my_tiles = (num_tiles - cluster_id + num_clusters - 1) / num_clusters;
            const int n_chunks = my_tiles * kChunks;
            auto chunk_pos = [&](int g, int& m0, int& c0) {
                const int tile = cluster_id + (g / kChunks) * num_clusters;
                int m_blk, n_blk;
                gemm2_tile_pos(p, tile, m_blk, n_blk);
                m0 = m_blk * kGemm2TileM + static_cast<int>(rank) * kGemmBlockM + quad * 32;
                c0 = n_blk * BLOCK_N + half * (BLOCK_N / 2) + (g % kChunks) * 32;
            };
            auto issue_load = [&](int g) {   // lane 0 only
                int m0, c0;
                chunk_pos(g, m0, c0);
                const uint32_t bar = smem_u32(&rbar[g % kResidH2Bufs]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(4096) : "memory");
                tma_load_2d_plain(wbase + (g % kResidH2Bufs) * 4096, &tmap_out, bar, c0, m0);
                tma_load_2d_plain(wbase + (g % kResidH2Bufs) * 4096 + 2048, &tmap_out16, bar, c0, m0);
            };
            if (lane == 0)
                for (int g = 0; g < kResidH2Bufs - 1 && g < n_chunks; ++g) issue_load(g);
            float s1 = 0.f, s2 = 0.f;
            for (int g = 0; g < n_chunks; ++g) {
                const int c = g % kChunks;
                int m_base, col0;
                chunk_pos(g, m_base, col0);
                if (c == 0) {
                    mbar_wait(&tmem_full[as], aphase);
                    tcgen05_fence_after();
                }
                uint32_t acc[32];
                tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BLOCK_N + half * (BLOCK_N / 2) + c * 32, acc);
                const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);   // warp-uniform: broadcast loads
                // every bulk store committed so far has read its smem source: the buffer of chunk g - 1 is free;
                // start the load that runs kResidH2Bufs - 1 chunks ahead into it
                if (lane == 0) {
                    tma_store_wait_read();
                    if (g + kResidH2Bufs - 1 < n_chunks) issue_load(g + kResidH2Bufs - 1);
                }
                tmem_ld_wait();
                if (c + 1 == kChunks) {   // accumulator stage drained into registers: hand it back to the MMA warp
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[as]), 0));
                    if (++as == 2) { as = 0; aphase ^= 1; }
                }
                mbar_wait(&rbar[g % kResidH2Bufs], (g / kResidH2Bufs) & 1);
                const uint32_t hrow = wbase + (g % kResidH2Bufs) * 4096 + lane * 64;
                const uint32_t lrow = hrow + 2048;
#pragma unroll
                for (int j = 0; j < 4; ++j) {     // 16-byte slot j = columns 8 j .. 8 j + 7 of this thread's row
                    const uint32_t so = static_cast<uint32_t>((j ^ ((lane >> 1) & 3)) << 4);
                    const uint4 hh = lds_v4u(hrow + so), ll = lds_v4u(lrow + so);
                    const uint32_t hs[4] = {hh.x, hh.y, hh.z, hh.w}, ls[4] = {ll.x, ll.y, ll.z, ll.w};
                    const float4 ba = __ldg(b4 + 2 * j), bb = __ldg(b4 + 2 * j + 1);
                    const float bs[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
                    uint32_t nh[4], nl[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 hf = unpack_op16x2(hs[e]), lf = unpack_op16x2(ls[e]);
                        const float v0 = (hf.x + lf.x) + (__uint_as_float(acc[8 * j + 2 * e]) + bs[2 * e]);
                        const float v1 = (hf.y + lf.y) + (__uint_as_float(acc[8 * j + 2 * e + 1]) + bs[2 * e + 1]);
                        split_op16x2(v0, v1, nh[e], nl[e]);
                        s1 += v0 + v1;
                        s2 += v0 * v0 + v1 * v1;
                    }
                    sts_v4u(hrow + so, make_uint4(nh[0], nh[1], nh[2], nh[3]));
                    sts_v4u(lrow + so, make_uint4(nl[0], nl[1], nl[2], nl[3]));
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tmap_out, wbase + (g % kResidH2Bufs) * 4096, col0, m_base);
                    tma_store_2d(&tmap_out16, wbase + (g % kResidH2Bufs) * 4096 + 2048, col0, m_base);
                    tma_store_commit();
                }
                if (c + 1 == kChunks) {   // this thread's row is complete for this half tile
                    const int m = m_base + lane;
                    const int n_blk = (col0 - half * (BLOCK_N / 2)) / BLOCK_N;
                    if (m < p.m_valid) p.stats_out[static_cast<size_t>(n_blk * 2 + half) * p.stats_ld + m] = make_float2(s1, s2);
                    s1 = s2 = 0.f;
                }
            }
        } else {
            const int half = ew >> 2;               // which half of the tile's columns this warp drains
            constexpr int kChunks = BLOCK_N / 64;   // 32-column chunks per warp
            static_assert(kChunks % 2 == 0, "the side-value double buffer alternates per chunk");
            const uint32_t stg = smem_u32(staging + ew * 4096);
            auto tile_rows = [&](int tile, int& n_blk) {
                int m_blk;
                gemm2_tile_pos(p, tile, m_blk, n_blk);
                return m_blk * kGemm2TileM + static_cast<int>(rank) * kGemmBlockM + quad * 32;
            };
            float4 side[2][8];
            int tile = cluster_id;
            if (tile < num_tiles) {
                int n_blk;
                const int m_base = tile_rows(tile, n_blk);
                gemm2_load_side<EPI>(p, m_base, n_blk * BLOCK_N + half * (BLOCK_N / 2), lane, side[0]);
            }
            for (; tile < num_tiles; tile += num_clusters) {
                int n_blk, n_blk_next = 0;
                const int m_base = tile_rows(tile, n_blk);
                const int col_base = n_blk * BLOCK_N + half * (BLOCK_N / 2);
                const bool has_next = tile + num_clusters < num_tiles;
                const int m_base_next = has_next ? tile_rows(tile + num_clusters, n_blk_next) : 0;
                float s1[8], s2[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.f;
#ifdef MCM_GEMM_TRACE
                const long long tw0 = clock64();
#endif
                mbar_wait(&tmem_full[as], aphase);
#ifdef MCM_GEMM_TRACE
                const long long tw1 = clock64();
#endif
                tcgen05_fence_after();
                const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BLOCK_N + half * (BLOCK_N / 2);
#pragma unroll
                for (int c = 0; c < kChunks; ++c) {
                    if (c + 1 < kChunks)
                        gemm2_load_side<EPI>(p, m_base, col_base + (c + 1) * 32, lane, side[(c + 1) & 1]);
                    else if (has_next)
                        gemm2_load_side<EPI>(p, m_base_next, n_blk_next * BLOCK_N + half * (BLOCK_N / 2), lane, side[0]);
                    if (m_base < p.m_valid) {
                        float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
                        if constexpr (T::kResid)
                            bias = __ldg(reinterpret_cast<const float4*>(p.bias + col_base + c * 32 + 4 * (lane & 7)));
                        gemm2_epilogue_chunk<EPI, SPLIT>(p, stg, t_row + c * 32, m_base, col_base + c * 32, lane, bias, side[c & 1], s1, s2);
                    }
                }
                if constexpr (T::kStats) {
                    // the 8 lanes that share a row (cq = 0..7) fold their partial sums; lane cq == 0 writes them
                    // (transposing butterfly: 7 shuffles per quantity; lane cq ends up with the totals of row r0 + 4 cq)
                    if (m_base < p.m_valid) {
#pragma unroll
                        for (int w = 4; w >= 1; w >>= 1) {
                            const bool up = (lane & w) != 0;
#pragma unroll
                            for (int i = 0; i < w; ++i) {
                                const float a1 = __shfl_xor_sync(0xffffffffu, up ? s1[i] : s1[i + w], w);
                                const float a2 = __shfl_xor_sync(0xffffffffu, up ? s2[i] : s2[i + w], w);
                                s1[i] = (up ? s1[i + w] : s1[i]) + a1;
                                s2[i] = (up ? s2[i + w] : s2[i]) + a2;
                            }
                        }
                        const int m = m_base + (lane >> 3) + 4 * (lane & 7);
                        if (m < p.m_valid)
                            p.stats_out[static_cast<size_t>(n_blk * 2 + half) * p.stats_ld + m] = make_float2(s1[0], s2[0]);
                    }
                }
                tcgen05_fence_before();
                __syncwarp();
#ifdef MCM_GEMM_TRACE
                if (p.trace && warp == 2 && lane == 0 && rank == 0) {
                    atomicAdd(reinterpret_cast<unsigned long long*>(p.trace + 4), (unsigned long long)(tw1 - tw0));
                    atomicAdd(reinterpret_cast<unsigned long long*>(p.trace + 5), (unsigned long long)(clock64() - tw1));
                }
#endif
                if (lane == 0) {
                    if (rank == 0) mbar_arrive(&tmem_empty[as]);
                    else mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[as]), 0));
                }
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    }

    if constexpr ((T::kF16 && MCM_GEMM_F16_TMA_STORE) || T::kTmaResid || T::kTmaResidH2) {
        if (warp >= 2 && lane == 0) tma_store_wait_all();   // bulk stores must have completed before the CTA exits
    }
    __syncwarp();
    tcgen05_fence_before();
    cluster_sync_all();   // no CTA may exit (or free TMEM) while its peer can still signal its barriers
    if (warp == 1) {
        __syncwarp();
        tcgen05_fence_after();
        tmem_dealloc_cta2<kTmemCols>(tmem_base);
    }
}

}  // namespace mcm
