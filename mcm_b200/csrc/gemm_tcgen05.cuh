// Persistent, warp-specialised fp16 GEMM for sm_100a:   C[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue)
//
//   * A (activations) and W (nn.Linear weight, [out,in]) are both K-major fp16 in HBM.
//   * TMA (cp.async.bulk.tensor, 128-byte swizzle) stages 128 x 64 A tiles and BLOCK_N x 64 W
//     tiles into a shared-memory ring; one elected thread issues tcgen05.mma (UMMA 128 x BLOCK_N x 16,
//     fp32 accumulation in TMEM); four epilogue warps drain TMEM with tcgen05.ld and apply the fused
//     epilogue.  TMEM holds two accumulator stages so the epilogue of tile i overlaps the MMAs of
//     tile i+1.  The grid is one CTA per SM, tiles are strided over CTAs with N fastest so CTAs that
//     run concurrently share the same A rows in L2.
//
// Replaces the cuBLAS SGEMM calls behind nn.Linear in HF CLIP:
//   q/k/v_proj HF:modeling_clip.py:310-312, out_proj :334, fc1/fc2 :348-350, patch conv :148-154.
#pragma once
#include <cuda.h>
#include "ptx.cuh"

namespace mcm {

enum GemmEpilogue : int {
    EPI_BIAS_F16 = 0,        // out_f16 = acc + bias                               (QKV projection)
    EPI_BIAS_QGELU_F16 = 1,  // out_f16 = quick_gelu(acc + bias)                    (fc1, HF activations.py:117-123)
    EPI_BIAS_RESID_F32 = 2,   // out_f32  = resid_f32 + acc + bias  (in place ok)     (out_proj / fc2 + residual)
    EPI_POS_F32 = 3,          // out_f32[b*S + 1 + p] = acc + pos[1 + p]               (patch embedding, no bias)
};

struct GemmParams {
    int m_tiles;        // M_pad / 128
    int n_tiles;        // N / BLOCK_N
    int k_blocks;       // K / 64
    int m_valid;        // rows >= m_valid are computed but never stored
    int ldo;            // leading dimension (elements) of out / resid
    const float* bias;  // [N] (unused for EPI_POS_F32)
    void* out;          // fp16 or fp32
    const float* resid; // EPI_BIAS_RESID_F32
    const float* pos;   // EPI_POS_F32: position embedding [S, N]
    int np;             // EPI_POS_F32: patches per image
    int seq;            // EPI_POS_F32: tokens per image (np + 1)
};

constexpr int kGemmBlockM = 128;
constexpr int kGemmBlockK = 64;    // 64 fp16 = one 128-byte swizzle row
constexpr int kGemmThreads = 192;  // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue

template <int BLOCK_N>
struct GemmSmem {
    static constexpr int kABytes = kGemmBlockM * kGemmBlockK * 2;
    static constexpr int kBBytes = BLOCK_N * kGemmBlockK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = (BLOCK_N == 256) ? 4 : 6;
    static constexpr int kBarrierBytes = 1024;
    static constexpr int kTotal = kStages * kStageBytes + kBarrierBytes + 1024 /* alignment slack */;
};

template <int EPI>
__device__ __forceinline__ void gemm_epilogue_chunk(const GemmParams& p, const uint32_t (&acc)[32], int m, int n) {
    // this thread owns row m, columns [n, n+32)
    if constexpr (EPI == EPI_BIAS_F16 || EPI == EPI_BIAS_QGELU_F16) {
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + n);
        uint32_t packed[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float4 b = __ldg(b4 + j);
            float v0 = __uint_as_float(acc[4 * j + 0]) + b.x;
            float v1 = __uint_as_float(acc[4 * j + 1]) + b.y;
            float v2 = __uint_as_float(acc[4 * j + 2]) + b.z;
            float v3 = __uint_as_float(acc[4 * j + 3]) + b.w;
            if constexpr (EPI == EPI_BIAS_QGELU_F16) {
                v0 = __fdividef(v0, 1.0f + __expf(-1.702f * v0));
                v1 = __fdividef(v1, 1.0f + __expf(-1.702f * v1));
                v2 = __fdividef(v2, 1.0f + __expf(-1.702f * v2));
                v3 = __fdividef(v3, 1.0f + __expf(-1.702f * v3));
            }
            packed[2 * j + 0] = pack_op16x2(v0, v1);
            packed[2 * j + 1] = pack_op16x2(v2, v3);
        }
        if (m < p.m_valid) {
            uint4* o = reinterpret_cast<uint4*>(static_cast<op16_t*>(p.out) + static_cast<size_t>(m) * p.ldo + n);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                o[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
        }
    } else if constexpr (EPI == EPI_BIAS_RESID_F32) {
        if (m < p.m_valid) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + n);
            const float4* r4 = reinterpret_cast<const float4*>(p.resid + static_cast<size_t>(m) * p.ldo + n);
            float4* o4 = reinterpret_cast<float4*>(static_cast<float*>(p.out) + static_cast<size_t>(m) * p.ldo + n);
            float4 r[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = r4[j];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 b = __ldg(b4 + j);
                float4 v;
                v.x = r[j].x + (__uint_as_float(acc[4 * j + 0]) + b.x);
                v.y = r[j].y + (__uint_as_float(acc[4 * j + 1]) + b.y);
                v.z = r[j].z + (__uint_as_float(acc[4 * j + 2]) + b.z);
                v.w = r[j].w + (__uint_as_float(acc[4 * j + 3]) + b.w);
                o4[j] = v;
            }
        }
    } else {  // EPI_POS_F32
        if (m < p.m_valid) {
            const int b = m / p.np;
            const int pi = m - b * p.np;
            const size_t orow = static_cast<size_t>(b) * p.seq + 1 + pi;
            const float4* e4 = reinterpret_cast<const float4*>(p.pos + static_cast<size_t>(1 + pi) * p.ldo + n);
            float4* o4 = reinterpret_cast<float4*>(static_cast<float*>(p.out) + orow * p.ldo + n);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 e = __ldg(e4 + j);
                float4 v;
                v.x = __uint_as_float(acc[4 * j + 0]) + e.x;
                v.y = __uint_as_float(acc[4 * j + 1]) + e.y;
                v.z = __uint_as_float(acc[4 * j + 2]) + e.z;
                v.w = __uint_as_float(acc[4 * j + 3]) + e.w;
                o4[j] = v;
            }
        }
    }
}

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_f16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const GemmParams p) {
    using L = GemmSmem<BLOCK_N>;
    constexpr int kStages = L::kStages;
    constexpr uint32_t kTmemCols = 2 * BLOCK_N;  // two accumulator stages
    static_assert(BLOCK_N == 128 || BLOCK_N == 256, "BLOCK_N");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * L::kStageBytes);
    uint64_t* full_bar = bars;                    // [kStages]  TMA -> MMA
    uint64_t* empty_bar = bars + kStages;         // [kStages]  MMA -> TMA
    uint64_t* tmem_full = bars + 2 * kStages;     // [2]        MMA -> epilogue
    uint64_t* tmem_empty = bars + 2 * kStages + 2;  // [2]      epilogue -> MMA
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = p.m_tiles * p.n_tiles;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4);
        }
        fence_barrier_init();
    }
    pdl_launch_dependents();
    if (warp == 1) tmem_alloc<kTmemCols>(tmem_ptr);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();

    if (warp == 0) {
        if (elect_one()) {
            // ===== TMA producer =====
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile / p.n_tiles;
                const int n_blk = tile - m_blk * p.n_tiles;
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * L::kStageBytes;
                    uint8_t* sb = sa + L::kABytes;
                    mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
                    tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * kGemmBlockK, m_blk * kGemmBlockM);
                    tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * kGemmBlockK, n_blk * BLOCK_N);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // ===== MMA issuer =====
            constexpr uint32_t idesc = make_idesc_f16(kGemmBlockM, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
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
                    for (int k = 0; k < kGemmBlockK / 16; ++k) {
                        // +32 bytes (16 fp16) along K inside the swizzle atom == +2 in the 16-byte address field
                        umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tmem_full[as]);
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else {
        // ===== epilogue warps: TMEM -> registers -> fused epilogue -> HBM =====
        const int quad = warp & 3;  // TMEM lane quadrant this warp may access
        int as = 0;
        uint32_t aphase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_blk = tile / p.n_tiles;
            const int n_blk = tile - m_blk * p.n_tiles;
            const int m = m_blk * kGemmBlockM + quad * 32 + lane;
            mbar_wait(&tmem_full[as], aphase);
            tcgen05_fence_after();
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BLOCK_N;
#pragma unroll 1
            for (int c = 0; c < BLOCK_N; c += 32) {
                uint32_t acc[32];
                tmem_ld_32x32b_x32(t_row + c, acc);
                tmem_ld_wait();
                gemm_epilogue_chunk<EPI>(p, acc, m, n_blk * BLOCK_N + c);
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[as]);
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tcgen05_fence_after();
        tmem_dealloc<kTmemCols>(tmem_base);
    }
}

}  // namespace mcm
