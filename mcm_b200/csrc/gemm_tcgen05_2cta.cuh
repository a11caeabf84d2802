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

constexpr int kGemm2Threads = 320;   // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-9: epilogue
constexpr int kGemm2EpiWarps = 8;    // two per TMEM lane quadrant (= per SM sub-partition): one column half each
constexpr int kGemm2TileM = 256;     // rows per cluster tile (128 per CTA)
constexpr int kStgLd = 32;           // staging row stride in floats; 16-byte chunks are XOR-swizzled by (row & 7)

template <int BLOCK_N>
struct Gemm2Smem {
    static constexpr int kABytes = kGemmBlockM * kGemmBlockK * 2;          // 128 x 64 fp16
    static constexpr int kBBytes = (BLOCK_N / 2) * kGemmBlockK * 2;        // this CTA's half of W
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = (BLOCK_N == 256) ? 6 : 8;
    static constexpr int kStagingBytes = kGemm2EpiWarps * 32 * kStgLd * 4;
    static constexpr int kBarrierBytes = 256;
    static constexpr int kTotal = kStages * kStageBytes + kStagingBytes + kBarrierBytes + 1024 /* alignment slack */;
    static_assert(kTotal <= 232448, "exceeds the 227 KB of shared memory a CTA can opt in to");
    static_assert((2 * kStages + 4) * 8 + 4 <= kBarrierBytes, "barrier area too small");
};

// ---- cluster / 2-CTA PTX ----
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
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
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

// One 32-row x 32-column chunk of the accumulator leaves through a swizzled smem transpose so that
// global accesses are 4 rows x 128 B (fp32) or 4 rows x 64 B (fp16) per warp instruction.
// All global reads of the chunk (residual rows / position rows) are issued BEFORE the TMEM load and
// the transpose so their latency overlaps them.  t_addr: TMEM address (lane quadrant + column) of the
// chunk; m_base: global row of this warp's first row; col0: first global column of the chunk.
template <int EPI>
__device__ __forceinline__ void gemm2_epilogue_chunk(const GemmParams& p, uint32_t stg, uint32_t t_addr, int m_base, int col0,
                                                     int lane, const float4 bias) {
    const int cq = lane & 7;       // this lane's 4-column group
    const int r0 = lane >> 3;      // rows r0, r0 + 4, ..., r0 + 28
    const int col = col0 + 4 * cq;
    float4 side[8];                // residual (EPI 2) or position-embedding (EPI 3) values
    size_t orow[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m_base + r0 + 4 * i;
        side[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        orow[i] = static_cast<size_t>(m);
        if constexpr (EPI == EPI_BIAS_RESID_F32) {
            if (m < p.m_valid) side[i] = *reinterpret_cast<const float4*>(p.resid + static_cast<size_t>(m) * p.ldo + col);
        } else if constexpr (EPI == EPI_POS_F32) {
            // patch row m of image b -> token row b * seq + 1 + patch, plus its position embedding
            const int b = m / p.np;
            const int pi = m - b * p.np;
            orow[i] = static_cast<size_t>(b) * p.seq + 1 + pi;
            if (m < p.m_valid) side[i] = __ldg(reinterpret_cast<const float4*>(p.pos + static_cast<size_t>(1 + pi) * p.ldo + col));
        }
    }
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
        if (m < p.m_valid) {
            if constexpr (EPI == EPI_BIAS_F16 || EPI == EPI_BIAS_QGELU_F16) {
                v.x += bias.x; v.y += bias.y; v.z += bias.z; v.w += bias.w;
                if constexpr (EPI == EPI_BIAS_QGELU_F16) {
                    v.x = __fdividef(v.x, 1.0f + __expf(-1.702f * v.x));
                    v.y = __fdividef(v.y, 1.0f + __expf(-1.702f * v.y));
                    v.z = __fdividef(v.z, 1.0f + __expf(-1.702f * v.z));
                    v.w = __fdividef(v.w, 1.0f + __expf(-1.702f * v.w));
                }
                *reinterpret_cast<uint2*>(static_cast<op16_t*>(p.out) + orow[i] * p.ldo + col) =
                    make_uint2(pack_op16x2(v.x, v.y), pack_op16x2(v.z, v.w));
            } else if constexpr (EPI == EPI_BIAS_RESID_F32) {
                v.x = side[i].x + (v.x + bias.x); v.y = side[i].y + (v.y + bias.y);
                v.z = side[i].z + (v.z + bias.z); v.w = side[i].w + (v.w + bias.w);
                *reinterpret_cast<float4*>(static_cast<float*>(p.out) + orow[i] * p.ldo + col) = v;
            } else {  // EPI_POS_F32
                v.x += side[i].x; v.y += side[i].y; v.z += side[i].z; v.w += side[i].w;
                *reinterpret_cast<float4*>(static_cast<float*>(p.out) + orow[i] * p.ldo + col) = v;
            }
        }
    }
    __syncwarp();
}

// fp16-output epilogues (bias, bias + quick_gelu): the math runs on the TMEM registers (row per
// thread, 32 independent columns -> deep MUFU pipelining), only the packed fp16 result (64 B per
// row) goes through the staging transpose, and HBM sees 8 rows x 64 B per warp instruction.
template <int EPI>
__device__ __forceinline__ void gemm2_epilogue_chunk_f16(const GemmParams& p, uint32_t stg, uint32_t t_addr, int m_base,
                                                         int col0, int lane) {
    static_assert(EPI == EPI_BIAS_F16 || EPI == EPI_BIAS_QGELU_F16, "fp16-output epilogues only");
    uint32_t acc[32];
    tmem_ld_32x32b_x32(t_addr, acc);
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);   // warp-uniform addresses: broadcast loads
    float4 bias[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) bias[j] = __ldg(b4 + j);
    tmem_ld_wait();
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float v0 = __uint_as_float(acc[4 * j + 0]) + bias[j].x;
        float v1 = __uint_as_float(acc[4 * j + 1]) + bias[j].y;
        float v2 = __uint_as_float(acc[4 * j + 2]) + bias[j].z;
        float v3 = __uint_as_float(acc[4 * j + 3]) + bias[j].w;
        if constexpr (EPI == EPI_BIAS_QGELU_F16) {
            v0 = __fdividef(v0, 1.0f + __expf(-1.702f * v0));
            v1 = __fdividef(v1, 1.0f + __expf(-1.702f * v1));
            v2 = __fdividef(v2, 1.0f + __expf(-1.702f * v2));
            v3 = __fdividef(v3, 1.0f + __expf(-1.702f * v3));
        }
        pk[2 * j + 0] = pack_op16x2(v0, v1);
        pk[2 * j + 1] = pack_op16x2(v2, v3);
    }
    // staging tile: 32 rows x 64 B; 16-byte slot s of row r lives at slot s ^ ((r >> 1) & 3)  (conflict-free both ways)
    const uint32_t srow = stg + lane * 64;
    const int sw = (lane >> 1) & 3;
#pragma unroll
    for (int s4 = 0; s4 < 4; ++s4)
        sts_v4u(srow + ((s4 ^ sw) << 4), make_uint4(pk[4 * s4], pk[4 * s4 + 1], pk[4 * s4 + 2], pk[4 * s4 + 3]));
    __syncwarp();
    const int slot = lane & 3;     // this lane's 8-column group
    const int r0 = lane >> 2;      // rows r0, r0 + 8, r0 + 16, r0 + 24
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + 8 * i;
        const int m = m_base + r;
        const uint4 v = lds_v4u(stg + r * 64 + ((slot ^ ((r >> 1) & 3)) << 4));
        if (m < p.m_valid)
            *reinterpret_cast<uint4*>(static_cast<op16_t*>(p.out) + static_cast<size_t>(m) * p.ldo + col0 + slot * 8) = v;
    }
    __syncwarp();
}

// p.m_tiles counts 256-row pair tiles.  PAIRS = 1: cluster of 2 CTAs (one pair).  PAIRS = 2: cluster of
// 4 CTAs = two pairs working on vertically adjacent 256-row tiles of the SAME n-block; every CTA
// fetches only a QUARTER of the W tile and multicasts it to its twin in the other pair, which cuts the
// L2 -> SM operand traffic (the measured limiter of the k-loop) from 64 KB to 48 KB per pair k-block.
// Launched with cudaLaunchKernelEx + cluster dimension 2 * PAIRS.
template <int BLOCK_N, int EPI, int PAIRS>
__global__ void __launch_bounds__(kGemm2Threads, 1)
gemm_f16_tn_cta2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                        const GemmParams p) {
    using L = Gemm2Smem<BLOCK_N>;
    constexpr int kStages = L::kStages;
    constexpr uint32_t kTmemCols = 2 * BLOCK_N;  // two accumulator stages
    static_assert(BLOCK_N == 128 || BLOCK_N == 256, "BLOCK_N");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* staging = reinterpret_cast<float*>(smem + kStages * L::kStageBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * L::kStageBytes + L::kStagingBytes);
    uint64_t* full_bar = bars;                      // [kStages]
    uint64_t* empty_bar = bars + kStages;           // [kStages]
    uint64_t* tmem_full = bars + 2 * kStages;       // [2]
    uint64_t* tmem_empty = bars + 2 * kStages + 2;  // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    static_assert(PAIRS == 1 || PAIRS == 2, "PAIRS");
    const uint32_t crank = cluster_ctarank();          // 0 .. 2 * PAIRS - 1
    const uint32_t rank = crank & 1;                   // rank inside the CTA pair (0 = leader)
    const uint32_t pair = crank >> 1;
    const uint32_t leader = crank & ~1u;               // cluster rank of this pair's leader
    const int cluster_id = blockIdx.x / (2 * PAIRS);
    const int num_clusters = gridDim.x / (2 * PAIRS);
    const int m_ctiles = (p.m_tiles + PAIRS - 1) / PAIRS;      // cluster tiles along M
    const int num_tiles = m_ctiles * p.n_tiles;
    constexpr uint16_t kAllMask = (1u << (2 * PAIRS)) - 1;
    const uint16_t pair_mask = static_cast<uint16_t>(3u << (2 * pair));

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], PAIRS);   // every pair that reads this stage (all write into it) must release it
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 2 * kGemm2EpiWarps);
        }
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
            // ===== TMA producer (every CTA): own A rows, own share of W; bytes land on the pair leader's full barrier =====
            int stage = 0;
            uint32_t phase = 0;
            constexpr int kWRows = BLOCK_N / (2 * PAIRS);     // W rows this CTA fetches per k-block
            const uint16_t w_mask = static_cast<uint16_t>((1u << rank) | (PAIRS == 2 ? (1u << (rank + 2)) : 0u));
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                const int mc = tile / p.n_tiles;
                const int n_blk = tile - mc * p.n_tiles;
                const int m_blk = mc * PAIRS + static_cast<int>(pair);
                const int a_row = m_blk * kGemm2TileM + static_cast<int>(rank) * kGemmBlockM;
                const int b_row = n_blk * BLOCK_N + static_cast<int>(rank) * (BLOCK_N / 2) + static_cast<int>(pair) * kWRows;
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * L::kStageBytes;
                    uint8_t* sb = sa + L::kABytes + static_cast<int>(pair) * (kWRows * kGemmBlockK * 2);
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
                    const uint32_t bar = mapa_shared(smem_u32(&full_bar[stage]), leader);
                    tma_load_2d_cta2(sa, &tmap_a, bar, kb * kGemmBlockK, a_row);
                    if constexpr (PAIRS == 1)
                        tma_load_2d_cta2(sb, &tmap_b, bar, kb * kGemmBlockK, b_row);
                    else
                        tma_load_2d_cta2_mc(sb, &tmap_b, bar, kb * kGemmBlockK, b_row, w_mask);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0 && elect_one()) {
            // ===== MMA issuer (pair leaders only) =====
            constexpr uint32_t idesc = make_idesc_f16(kGemm2TileM, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                mbar_wait(&tmem_empty[as], aphase ^ 1);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + as * BLOCK_N;
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
                    const uint64_t adesc = make_smem_desc_sw128(sa, 16, 1024);
                    const uint64_t bdesc = make_smem_desc_sw128(sa + L::kABytes, 16, 1024);
#pragma unroll
                    for (int k = 0; k < kGemmBlockK / 16; ++k)
                        umma_f16_cta2(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                    umma_commit_cta2_mc(&empty_bar[stage], kAllMask);   // the stage is written by CTAs of every pair
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                umma_commit_cta2_mc(&tmem_full[as], pair_mask);
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else {
        // ===== epilogue warps (both CTAs): TMEM -> registers -> smem transpose -> coalesced HBM =====
        const int ew = warp - 2;
        const int quad = warp & 3;          // TMEM lane quadrant this warp may access
        const int half = ew >> 2;           // which half of the tile's columns this warp drains
        constexpr int kChunks = BLOCK_N / 64;   // 32-column chunks per warp
        const uint32_t stg = smem_u32(staging + ew * 32 * kStgLd);   // shared-space address of this warp's staging tile
        int as = 0;
        uint32_t aphase = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
            const int mc = tile / p.n_tiles;
            const int n_blk = tile - mc * p.n_tiles;
            const int m_blk = mc * PAIRS + static_cast<int>(pair);
            const int m_base = m_blk * kGemm2TileM + static_cast<int>(rank) * kGemmBlockM + quad * 32;
            const int col_base = n_blk * BLOCK_N + half * (BLOCK_N / 2);
            float4 bias[kChunks];
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
                bias[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                if constexpr (EPI == EPI_BIAS_RESID_F32)
                    bias[c] = __ldg(reinterpret_cast<const float4*>(p.bias + col_base + c * 32 + 4 * (lane & 7)));
            }
            mbar_wait(&tmem_full[as], aphase);
            tcgen05_fence_after();
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BLOCK_N + half * (BLOCK_N / 2);
            if (m_base < p.m_valid) {
#pragma unroll
                for (int c = 0; c < kChunks; ++c) {
                    if constexpr (EPI == EPI_BIAS_F16 || EPI == EPI_BIAS_QGELU_F16)
                        gemm2_epilogue_chunk_f16<EPI>(p, stg, t_row + c * 32, m_base, col_base + c * 32, lane);
                    else
                        gemm2_epilogue_chunk<EPI>(p, stg, t_row + c * 32, m_base, col_base + c * 32, lane, bias[c]);
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (rank == 0) mbar_arrive(&tmem_empty[as]);
                else mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[as]), leader));
            }
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
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
