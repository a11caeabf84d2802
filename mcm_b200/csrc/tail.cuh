// Fused scoring tail (all fp32): CLS pool -> post_layernorm -> visual_projection -> L2 normalise
// -> cosine logits against the prompt bank -> softmax(./T) reductions -> one score per image.
//
// Reference: HF:modeling_clip.py:685-686 (pooled = last_hidden_state[:,0]; post_layernorm),
// :860-861 (visual_projection, no bias); utils/detection_util.py:226 (normalise), :232 (logits),
// :233-248 (score reductions).  The reference ships the whole [B,K] softmax to the host and
// reduces there; here only [B] scores leave the kernel.
//
// One CTA scores kTailImgs images so every projection / bank row read from L2 is reused kTailImgs
// times; warps stride over output rows, lanes over the contraction dimension (float4, coalesced),
// warp-shuffle reductions finish each dot product.
#pragma once
#include "ptx.cuh"

namespace mcm {

constexpr int kTailImgs = 4;
constexpr int kTailThreads = 256;

enum ScoreKind : int { SCORE_MCM = 0, SCORE_MAX_LOGIT = 1, SCORE_ENERGY = 2, SCORE_ENTROPY = 3, SCORE_VAR = 4 };

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
    // red: >= 8 floats of scratch; returns the reduction to every thread
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    v = is_max ? warp_max(v) : warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int i = 1; i < kTailThreads / 32; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
    return r;
}

// smem layout (floats): ln[kTailImgs][D] | feat[kTailImgs][P] | z[kTailImgs][K] | red[8] | inv_norm[kTailImgs]
__global__ void __launch_bounds__(kTailThreads)
tail_kernel(const float* __restrict__ x, int S, int D, int P, int K, int b, const float* __restrict__ post_g,
            const float* __restrict__ post_b, float eps, const float* __restrict__ wproj /*[P,D]*/,
            const float* __restrict__ bank /*[K,P] unit rows*/, float T, int kind, float* __restrict__ feats,
            float* __restrict__ scores) {
    extern __shared__ float tsm[];
    float* s_ln = tsm;
    float* s_feat = s_ln + kTailImgs * D;
    float* s_z = s_feat + kTailImgs * P;
    float* s_red = s_z + kTailImgs * K;
    float* s_inv = s_red + 8;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NW = kTailThreads / 32;
    const int img0 = blockIdx.x * kTailImgs;
    const int nimg = min(kTailImgs, b - img0);

    // ---- post_layernorm of the CLS rows (one warp per image) ----
    for (int i = warp; i < kTailImgs; i += NW) {
        float* dst = s_ln + i * D;
        if (i < nimg) {
            const float* row = x + static_cast<size_t>(img0 + i) * S * D;
            float s = 0.f;
            for (int c = lane; c < D; c += 32) s += row[c];
            const float mean = warp_sum(s) / D;
            float q = 0.f;
            for (int c = lane; c < D; c += 32) {
                const float d = row[c] - mean;
                q += d * d;
            }
            const float rstd = rsqrtf(warp_sum(q) / D + eps);
            for (int c = lane; c < D; c += 32) dst[c] = (row[c] - mean) * rstd * __ldg(post_g + c) + __ldg(post_b + c);
        } else {
            for (int c = lane; c < D; c += 32) dst[c] = 0.f;
        }
    }
    __syncthreads();

    // ---- visual projection: feat[i][p] = sum_d ln[i][d] * W[p][d] ----
    const int D4 = D >> 2;
    for (int p = warp; p < P; p += NW) {
        const float4* w4 = reinterpret_cast<const float4*>(wproj + static_cast<size_t>(p) * D);
        float acc[kTailImgs];
#pragma unroll
        for (int i = 0; i < kTailImgs; ++i) acc[i] = 0.f;
        for (int c = lane; c < D4; c += 32) {
            const float4 w = __ldg(w4 + c);
#pragma unroll
            for (int i = 0; i < kTailImgs; ++i) {
                const float4 a = reinterpret_cast<const float4*>(s_ln + i * D)[c];
                acc[i] += (a.x * w.x + a.y * w.y) + (a.z * w.z + a.w * w.w);
            }
        }
#pragma unroll
        for (int i = 0; i < kTailImgs; ++i) acc[i] = warp_sum(acc[i]);
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < kTailImgs; ++i) s_feat[i * P + p] = acc[i];
        }
    }
    __syncthreads();

    if (feats != nullptr) {
        for (int t = threadIdx.x; t < nimg * P; t += kTailThreads) feats[static_cast<size_t>(img0) * P + t] = s_feat[t];
    }
    if (scores == nullptr) return;

    // ---- 1 / ||feat||  (utils/detection_util.py:226) ----
    for (int i = warp; i < kTailImgs; i += NW) {
        float q = 0.f;
        for (int c = lane; c < P; c += 32) q += s_feat[i * P + c] * s_feat[i * P + c];
        q = warp_sum(q);
        if (lane == 0) s_inv[i] = (i < nimg) ? 1.0f / sqrtf(q) : 0.f;
    }
    __syncthreads();

    // ---- cosine logits z[i][k] = (feat[i] . bank[k]) / ||feat[i]||   (:232) ----
    const int P4 = P >> 2;
    for (int k = warp; k < K; k += NW) {
        const float4* t4 = reinterpret_cast<const float4*>(bank + static_cast<size_t>(k) * P);
        float acc[kTailImgs];
#pragma unroll
        for (int i = 0; i < kTailImgs; ++i) acc[i] = 0.f;
        for (int c = lane; c < P4; c += 32) {
            const float4 w = __ldg(t4 + c);
#pragma unroll
            for (int i = 0; i < kTailImgs; ++i) {
                const float4 a = reinterpret_cast<const float4*>(s_feat + i * P)[c];
                acc[i] += (a.x * w.x + a.y * w.y) + (a.z * w.z + a.w * w.w);
            }
        }
#pragma unroll
        for (int i = 0; i < kTailImgs; ++i) acc[i] = warp_sum(acc[i]);
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < kTailImgs; ++i) s_z[i * K + k] = acc[i] * s_inv[i];
        }
    }
    __syncthreads();

    // ---- score reductions over k (:233-248) ----
    const float invT = 1.0f / T;
    for (int i = 0; i < nimg; ++i) {
        const float* z = s_z + i * K;
        float m = -INFINITY;
        for (int k = threadIdx.x; k < K; k += kTailThreads) m = fmaxf(m, z[k]);
        m = block_reduce(m, s_red, true);
        float out;
        if (kind == SCORE_MAX_LOGIT) {
            out = -m;
        } else {
            float se = 0.f, sez = 0.f;
            for (int k = threadIdx.x; k < K; k += kTailThreads) {
                const float d = (z[k] - m) * invT;
                const float e = expf(d);
                se += e;
                sez += e * d;
            }
            se = block_reduce(se, s_red, false);
            if (kind == SCORE_MCM) {
                out = -1.0f / se;  // max_k softmax = exp(0) / sum
            } else if (kind == SCORE_ENERGY) {
                out = -T * (m * invT + logf(se));
            } else if (kind == SCORE_ENTROPY) {
                sez = block_reduce(sez, s_red, false);
                out = logf(se) - sez / se;  // -sum p log p
            } else {  // SCORE_VAR: -mean_k (p_k - 1/K)^2  (np.var; second pass avoids cancellation)
                const float mp = 1.0f / K, inv_se = 1.0f / se;
                float dv = 0.f;
                for (int k = threadIdx.x; k < K; k += kTailThreads) {
                    const float pk = expf((z[k] - m) * invT) * inv_se - mp;
                    dv += pk * pk;
                }
                dv = block_reduce(dv, s_red, false);
                out = -dv / K;
            }
        }
        if (threadIdx.x == 0) scores[img0 + i] = out;
    }
}

// bank rows /= ||row||   (utils/detection_util.py:231); one warp per row
__global__ void normalize_rows_kernel(float* __restrict__ bank, int K, int P) {
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
