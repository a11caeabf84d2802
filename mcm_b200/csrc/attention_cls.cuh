// Last-layer shortcut: only the CLS token of the last encoder layer is consumed downstream
// (pooled = last_hidden_state[:, 0], HF:modeling_clip.py:685), so after the last layer's QKV
// projection every remaining op of that layer is needed for ONE query row per image.
//
// This kernel does the attention of that single query against all S keys of its image / head
// (HF:261-279 restricted to query 0) and gathers the CLS rows of the fp32 residual stream into a
// compact [b, D] buffer, so out-proj, LN2, fc1, fc2 of the last layer run on b rows instead of b*S.
// One warp per (image, head): lanes over keys for the scores, lanes over head dims for P.V.
#pragma once
#include "ptx.cuh"

namespace mcm {

constexpr int kClsWarps = 4;
constexpr int kClsMaxS = 320;

__global__ void __launch_bounds__(kClsWarps * 32)
attention_cls_kernel(const op16_t* __restrict__ qkv, const float* __restrict__ x, op16_t* __restrict__ o_cls,
                     float* __restrict__ x_cls, int b, int S, int H, float scale) {
    __shared__ float s_q[kClsWarps][64];
    __shared__ float s_p[kClsWarps][kClsMaxS];
    pdl_launch_dependents();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * kClsWarps + warp;
    if (item >= b * H) return;
    const int img = item / H, h = item - img * H;
    const int D = H * 64, ld = 3 * D;
    const op16_t* base = qkv + static_cast<size_t>(img) * S * ld + h * 64;

    // gather this head's 64-column slice of the CLS row of the residual stream
    {
        const float2 v = *reinterpret_cast<const float2*>(x + static_cast<size_t>(img) * S * D + h * 64 + lane * 2);
        *reinterpret_cast<float2*>(x_cls + static_cast<size_t>(img) * D + h * 64 + lane * 2) = v;
    }
    {
        const float2 qf = unpack_op16x2(*reinterpret_cast<const uint32_t*>(base + lane * 2));
        s_q[warp][lane * 2] = qf.x;
        s_q[warp][lane * 2 + 1] = qf.y;
    }
    __syncwarp();

    // scores: lane j handles keys j, j + 32, ...
    float mx = -INFINITY;
    for (int j = lane; j < S; j += 32) {
        const uint4* kr = reinterpret_cast<const uint4*>(base + static_cast<size_t>(j) * ld + D);
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint4 u = __ldg(kr + c);
            const uint32_t* hp = reinterpret_cast<const uint32_t*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 kf = unpack_op16x2(hp[e]);
                acc = fmaf(kf.x, s_q[warp][c * 8 + e * 2], acc);
                acc = fmaf(kf.y, s_q[warp][c * 8 + e * 2 + 1], acc);
            }
        }
        acc *= scale;
        s_p[warp][j] = acc;
        mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < S; j += 32) {
        const float e = __expf(s_p[warp][j] - mx);
        s_p[warp][j] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();

    // o = sum_j p_j v_j: lane owns head dims 2 * lane, 2 * lane + 1
    float o0 = 0.f, o1 = 0.f;
    const op16_t* vbase = base + 2 * D + lane * 2;
    for (int j = 0; j < S; ++j) {
        const float2 vf = unpack_op16x2(*reinterpret_cast<const uint32_t*>(vbase + static_cast<size_t>(j) * ld));
        const float p = s_p[warp][j];
        o0 = fmaf(p, vf.x, o0);
        o1 = fmaf(p, vf.y, o1);
    }
    const float inv = 1.0f / sum;
    *reinterpret_cast<uint32_t*>(o_cls + static_cast<size_t>(img) * D + h * 64 + lane * 2) = pack_op16x2(o0 * inv, o1 * inv);
}

}  // namespace mcm
