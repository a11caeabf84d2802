// Fused scoring tail (all fp32): CLS pool -> post_layernorm -> visual_projection -> L2 normalise
// -> cosine logits against the prompt bank -> softmax(./T) reductions -> one score per image.
//
// Reference: HF:modeling_clip.py:685-686 (pooled = last_hidden_state[:,0]; post_layernorm),
// :860-861 (visual_projection, no bias); utils/detection_util.py:226 (normalise), :232 (logits),
// :233-248 (score reductions).  The reference ships the whole [B,K] softmax to the host and
// reduces there; here only [B] scores leave the kernel.
//
// Four small launches: pooled LayerNorm (one warp per image), two fp32 CUDA-core GEMMs (projection,
// cosine logits), and a one-warp-per-image reduction kernel (warp-shuffle max / sum-exp).
#pragma once
#include "ptx.cuh"

namespace mcm {


enum ScoreKind : int { SCORE_MCM = 0, SCORE_MAX_LOGIT = 1, SCORE_ENERGY = 2, SCORE_ENTROPY = 3, SCORE_VAR = 4 };

// ---- stage 1: post_layernorm of the pooled rows: row (img * row_stride) of x (fp32), or of the residual pair
//      xh + xl when x is null (the forward keeps the residual stream as fp16 (hi, lo) pairs) -> ln [b, D] ----
__global__ void __launch_bounds__(256)
pooled_layernorm_kernel(const float* __restrict__ x, const op16_t* __restrict__ xh, const op16_t* __restrict__ xl, size_t row_stride,
                        int D, int b, const float* __restrict__ g, const float* __restrict__ be, float eps, float* __restrict__ ln) {
    pdl_launch_dependents();
    pdl_wait();
    const int img = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (img >= b) return;
    const size_t r0 = static_cast<size_t>(img) * row_stride;
    auto at = [&](int c) -> float {
        if (x != nullptr) return x[r0 + c];
#ifdef MCM_OP_BF16
        return __bfloat162float(xh[r0 + c]) + __bfloat162float(xl[r0 + c]);
#else
        return __half2float(xh[r0 + c]) + __half2float(xl[r0 + c]);
#endif
    };
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s += at(c);
    const float mean = warp_sum(s) / D;
    float q = 0.f;
    for (int c = lane; c < D; c += 32) {
        const float d = at(c) - mean;
        q += d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) / D + eps);
    for (int c = lane; c < D; c += 32)
        ln[static_cast<size_t>(img) * D + c] = (at(c) - mean) * rstd * __ldg(g + c) + __ldg(be + c);
}

// ---- stages 2, 3: C[M, N] = A[M, Kd] . B[N, Kd]^T in fp32 on the CUDA cores ----
// (visual_projection: [b, D] x [P, D]^T; cosine logits: [b, P] x [K, P]^T.  0.5 GFLOP per 256 images;
// fp32 keeps the tail exact to the reference's own arithmetic.)  64 x 64 x 16 tiles, 256 threads,
// 4 x 4 outputs per thread, operands transposed through shared memory.
constexpr int kSgemmTile = 64, kSgemmK = 16;

__global__ void __launch_bounds__(256)
sgemm_tn_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int M, int N, int Kd) {
    __shared__ float sA[kSgemmK][kSgemmTile + 4];
    __shared__ float sB[kSgemmK][kSgemmTile + 4];
    pdl_launch_dependents();
    pdl_wait();
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * kSgemmTile, n0 = blockIdx.x * kSgemmTile;
    const int lr = tid >> 2;         // 0..63: tile row this thread loads
    const int lc = (tid & 3) * 4;    // 0,4,8,12: k offset of its float4
    const int ty = tid >> 4, tx = tid & 15;   // 16 x 16 thread grid, 4 x 4 outputs each
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const bool a_ok = (m0 + lr) < M, b_ok = (n0 + lr) < N;
    const float* ap = A + static_cast<size_t>(m0 + lr) * Kd + lc;
    const float* bp = B + static_cast<size_t>(n0 + lr) * Kd + lc;
    for (int k0 = 0; k0 < Kd; k0 += kSgemmK) {
        float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = av;
        if (a_ok) av = *reinterpret_cast<const float4*>(ap + k0);
        if (b_ok) bv = __ldg(reinterpret_cast<const float4*>(bp + k0));
        __syncthreads();
        sA[lc + 0][lr] = av.x; sA[lc + 1][lr] = av.y; sA[lc + 2][lr] = av.z; sA[lc + 3][lr] = av.w;
        sB[lc + 0][lr] = bv.x; sB[lc + 1][lr] = bv.y; sB[lc + 2][lr] = bv.z; sB[lc + 3][lr] = bv.w;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kSgemmK; ++k) {
            const float4 a4 = *reinterpret_cast<const float4*>(&sA[k][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&sB[k][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) C[static_cast<size_t>(m) * N + n] = acc[i][j];
        }
    }
}

// ---- stage 4: one warp per image: 1 / ||feat||, z = logits / ||feat||, score reductions (:226, :233-248) ----
__global__ void __launch_bounds__(256)
score_rows_kernel(const float* __restrict__ feats, const float* __restrict__ logits, int P, int K, int b, float T,
                  int kind, float* __restrict__ scores) {
    pdl_launch_dependents();
    pdl_wait();
    const int img = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (img >= b) return;
    const float* f = feats + static_cast<size_t>(img) * P;
    float q = 0.f;
    for (int c = lane; c < P; c += 32) q += f[c] * f[c];
    const float inv = 1.0f / sqrtf(warp_sum(q));
    const float* z = logits + static_cast<size_t>(img) * K;
    float m = -INFINITY;
    for (int k = lane; k < K; k += 32) m = fmaxf(m, z[k] * inv);
    m = warp_max(m);
    float out;
    if (kind == SCORE_MAX_LOGIT) {
        out = -m;
    } else {
        const float invT = 1.0f / T;
        float se = 0.f, sez = 0.f;
        for (int k = lane; k < K; k += 32) {
            const float d = (z[k] * inv - m) * invT;
            const float e = expf(d);
            se += e;
            sez += e * d;
        }
        se = warp_sum(se);
        if (kind == SCORE_MCM) {
            out = -1.0f / se;  // max_k softmax = exp(0) / sum
        } else if (kind == SCORE_ENERGY) {
            out = -T * (m * invT + logf(se));
        } else if (kind == SCORE_ENTROPY) {
            out = logf(se) - warp_sum(sez) / se;  // -sum p log p
        } else {  // SCORE_VAR: -mean_k (p_k - 1/K)^2  (np.var; second pass avoids cancellation)
            const float mp = 1.0f / K, inv_se = 1.0f / se;
            float dv = 0.f;
            for (int k = lane; k < K; k += 32) {
                const float pk = expf((z[k] * inv - m) * invT) * inv_se - mp;
                dv += pk * pk;
            }
            out = -warp_sum(dv) / K;
        }
    }
    if (lane == 0) scores[img] = out;
}

// ---- Mahalanobis baseline (reference `--score maha`, utils/detection_util.py:182-207) ----
//   score_i = -max_k( -0.5 (f_i - mu_k)^T P (f_i - mu_k) ) = 0.5 min_k |g_i - c_k|^2   with P = L L^T, g = f L, c_k = mu_k L:
// the K quadratic forms of the reference's Python loop (two [b,P]x[P,P] GEMMs per class) become ONE whitening GEMM
// (g = f L, sgemm_tn_kernel) and K squared distances per image, accumulated as differences (no cancellation).
// One warp per image; the whitened feature row stays in registers (P <= 1024), the class centres stream from L2.
__global__ void __launch_bounds__(256)
maha_min_dist_kernel(const float* __restrict__ g, const float* __restrict__ centres, int P, int K, int b, float* __restrict__ scores) {
    pdl_launch_dependents();
    pdl_wait();
    const int img = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (img >= b) return;
    float gr[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) gr[i] = (lane + 32 * i < P) ? g[static_cast<size_t>(img) * P + lane + 32 * i] : 0.f;
    float best = INFINITY;
    for (int k = 0; k < K; ++k) {
        const float* c = centres + static_cast<size_t>(k) * P;
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            if (lane + 32 * i < P) {
                const float z = gr[i] - __ldg(c + lane + 32 * i);
                d = fmaf(z, z, d);
            }
        }
        d += __shfl_xor_sync(0xffffffffu, d, 16);
        d += __shfl_xor_sync(0xffffffffu, d, 8);
        d += __shfl_xor_sync(0xffffffffu, d, 4);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        best = fminf(best, d);
    }
    if (lane == 0) scores[img] = 0.5f * best;
}


// bank rows /= ||row||   (utils/detection_util.py:231); one warp per row
__global__ void normalize_rows_kernel(float* __restrict__ bank, int K, int P) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= K) return;
    float* r = bank + static_cast<size_t>(row) * P;
    float q = 0.f;
    for (int c = lane; c < P; c += 32) q += r[c] * r[c];
    const float inv = 1.0f / sqrtf(warp_sum(q));
    for (int c = lane; c < P; c += 32) r[c] *= inv;
}

}  // namespace mcm
