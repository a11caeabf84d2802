// Parameters and epilogue kinds shared by the tcgen05 GEMM kernel (gemm_tcgen05_2cta.cuh) and the engine.
//
//     C[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue)
//
// A (activations) and W (nn.Linear weight, [out,in]) are both K-major fp16 in HBM.  The GEMMs replace
// the cuBLAS SGEMM calls behind nn.Linear in HF CLIP:
//   q/k/v_proj HF:modeling_clip.py:310-312, out_proj :334, fc1/fc2 :348-350, patch conv :148-154.
//
// LayerNorm folding (EPI_LN_*): the LayerNorm in front of a projection (layer_norm1 -> q/k/v,
// layer_norm2 -> fc1; HF:371-372,380-381) never materialises.  With  xn = (x - mu) * rstd * g + be:
//     xn @ W^T + b = rstd * (x @ (g o W)^T) - rstd * mu * c + d,   c_n = sum_k (g o W)_nk,  d = be @ W^T + b
// so the GEMM reads the RAW residual row (its fp16 copy), W is pre-multiplied by gamma when the weights
// are finalised, and the epilogue applies the per-row (mu, rstd) and the per-column (c, d).  The row
// statistics come from the epilogue that PRODUCED the residual (EPI_BIAS_RESID_F32_LN): every 128-column
// half tile adds its partial (sum, sum of squares) of the fp32 values it writes.
#pragma once
#include <cuda.h>
#include "ptx.cuh"

namespace mcm {

enum GemmEpilogue : int {
    EPI_BIAS_F16 = 0,          // out_f16 = acc + bias                               (plain projection)
    EPI_BIAS_QGELU_F16 = 1,    // out_f16 = quick_gelu(acc + bias)                    (HF activations.py:117-123)
    EPI_BIAS_RESID_F32 = 2,    // out_f32 = resid_f32 + acc + bias  (in place ok)     (out_proj / fc2 + residual)
    EPI_POS_F32 = 3,           // out_f32[b*S + 1 + p] = acc + pos[1 + p]             (patch embedding, no bias)
    EPI_LN_F16 = 4,            // out_f16 = rstd * acc - rstd * mu * c + d            (layer_norm1 folded into q/k/v)
    EPI_LN_QGELU_F16 = 5,      // out_f16 = quick_gelu(the same)                      (layer_norm2 folded into fc1)
    EPI_BIAS_RESID_F32_LN = 6, // EPI_BIAS_RESID_F32 + fp16 copy of the result + partial row statistics
    EPI_BIAS_RESID_F32_LN_TMA = 7,   // the same through TMA loads / stores (the tensor maps carry resid == out and out16)
    // The residual stream of the forward is an fp16 PAIR, x = hi + lo (hi = fp16(x), lo = fp16(x - hi); ~22 significant bits):
    // hi is the A operand the next projection reads anyway, so the pair costs 4 bytes per element where fp32 + fp16 copy
    // cost 6, and a residual epilogue moves 8 bytes per element (read pair, write pair) instead of 10.
    EPI_BIAS_RESID_H2_LN_TMA = 8,    // (hi, lo) = split(resid_hi + resid_lo + acc + bias) + partial row statistics, through TMA
    EPI_BIAS_RESID_H2_LN = 9,        // the same through the LSU (long-K GEMMs, the split-precision mode)
    EPI_KINDS = 10,
};

struct GemmParams {
    int m_tiles;        // ceil(M / 256)
    int n_tiles;        // N / BLOCK_N
    int k_blocks;       // K / 64
    int m_valid;        // rows >= m_valid are computed but never stored
    int m_reverse;      // walk the row blocks from the last to the first (L2 reuse along a producer -> consumer chain)
    int ldo;            // leading dimension (elements) of out / resid / out16
    const float* bias;  // [N] (unused for EPI_POS_F32); the folded d vector for EPI_LN_*
    void* out;          // fp16 or fp32
    const float* resid; // EPI_BIAS_RESID_F32*
    const op16_t* resid16;     // EPI_BIAS_RESID_H2_*: residual pair (may alias out16 / out16_lo)
    const op16_t* resid16_lo;
    const float* pos;   // EPI_POS_F32: position embedding [S, N]
    int np;             // EPI_POS_F32: patches per image
    int seq;            // EPI_POS_F32: tokens per image (np + 1)
    // ---- LayerNorm folding ----
    const float* colsum;     // EPI_LN_*: c[N]
    const float2* stats_in;  // EPI_LN_*: [stats_parts][stats_ld] partial (sum, sum of squares) of each A row's source
    int stats_parts;
    int stats_ld;            // rows per part of stats_in / stats_out
    float inv_k;             // 1 / (row length the statistics were taken over)
    float eps;
    op16_t* out16;           // EPI_BIAS_RESID_F32_LN: fp16 copy of out (the next projection's A operand); EPI_BIAS_RESID_H2_*: hi
    float2* stats_out;       // EPI_BIAS_RESID_F32_LN: [2 * n_tiles][stats_ld]
    // ---- split-fp16 precision mode (SPLIT kernels only; see "Precision modes" below) ----
    op16_t* out_lo;          // fp16-output epilogues: low half of out  (out + out_lo carry ~22 significant bits)
    op16_t* out16_lo;        // EPI_BIAS_RESID_F32_LN (split mode): low half of out16; EPI_BIAS_RESID_H2_*: lo
    long long* trace;        // -DMCM_GEMM_TRACE builds: [8] cycle counters summed over CTAs (see tools/gemm_trace.py)
#ifdef MCM_DEBUG
    int dbg_skip;            // timing experiments only (env MCM_GEMM_DBG_SKIP, -DMCM_DEBUG builds): 1 no global stores,
                             // 2 no staging either, 4 no TMEM drain at all, 8 no residual loads
#endif
};

#ifdef MCM_DEBUG
#define MCM_DBG_SKIP(p, bit) ((p).dbg_skip & (bit))
#else
#define MCM_DBG_SKIP(p, bit) 0
#endif

// Precision modes.
//   fast (default): every tensor-core operand is ONE fp16 value (11-bit significand); per-GEMM relative error ~7e-4.
//   split (MCM_OPT_PRECISION = 1): every operand is a PAIR of fp16 values, x = hi + lo with hi = fp16(x) and
//     lo = fp16(x - hi) (~22 significant bits), and every product is the three-term sum
//         A W^T  ~=  A_hi W_hi^T + A_lo W_hi^T + A_hi W_lo^T            (the dropped A_lo W_lo^T term is ~2^-22)
//     accumulated in the same fp32 TMEM tile: the k-loop of a SPLIT kernel walks the k-blocks three times over
//     (hi, hi), (lo, hi), (hi, lo) operand tiles.  fp32-class results at 3x the tensor work; this is the mode in which
//     AUROC / FPR95 agree with the fp32 reference to the last counted image.

constexpr int kGemmBlockM = 128;   // rows per CTA (256 per CTA pair)
constexpr int kGemmBlockK = 64;    // 64 fp16 = one 128-byte swizzle row

// quick_gelu(v) = v * sigmoid(1.702 v)  (HF activations.py:117-123).
// Default: one MUFU op per element through  sigmoid(z) = 0.5 + 0.5 tanh(z / 2)  (tanh.approx.f32, relative error
// 2^-11: the absolute error stays below half an fp16 ulp of |v|, i.e. below the rounding of the fp16 output itself).
// The fc1 epilogue is MUFU-bound with the two-op form (ex2 + rcp): 7.5 k cycles per 256 x 256 tile against a
// 6.1 k-cycle main loop.  -DMCM_GELU_EXACT restores ex2 + rcp (A/B builds).
// exact form for the split-precision mode: ex2.approx (2 ulp) and a full-precision division
__device__ __forceinline__ float quick_gelu_precise(float v) { return __fdiv_rn(v, 1.0f + __expf(-1.702f * v)); }

// x -> (hi, lo) fp16 pair of the split-precision mode
__device__ __forceinline__ void split_op16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = pack_op16x2(a, b);
    const float2 h = unpack_op16x2(hi);
    lo = pack_op16x2(a - h.x, b - h.y);
}

__device__ __forceinline__ float quick_gelu(float v) {
#ifdef MCM_GELU_EXACT
    return __fdividef(v, 1.0f + __expf(-1.702f * v));
#else
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * v));
    const float h = 0.5f * v;
    return fmaf(h, t, h);
#endif
}

}  // namespace mcm
