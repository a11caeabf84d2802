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

// qkv_lo / o_cls_lo (both null or both non-null): split-fp16 precision mode, values are hi + lo pairs.
// xh / xl: the residual stream (an fp16 (hi, lo) pair in both modes); x_cls receives the CLS rows in fp32.
__global__ void __launch_bounds__(kClsWarps * 32)
attention_cls_kernel(const op16_t* __restrict__ qkv, const op16_t* __restrict__ qkv_lo, const op16_t* __restrict__ xh,
                     const op16_t* __restrict__ xl, op16_t* __restrict__ o_cls, op16_t* __restrict__ o_cls_lo,
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
    const bool split = qkv_lo != nullptr;
    const op16_t* base_lo = split ? qkv_lo + static_cast<size_t>(img) * S * ld + h * 64 : nullptr;

    // gather this head's 64-column slice of the CLS row of the residual stream
    {
        const size_t off = static_cast<size_t>(img) * S * D + h * 64 + lane * 2;
        const float2 vh = unpack_op16x2(*reinterpret_cast<const uint32_t*>(xh + off)), vl = unpack_op16x2(*reinterpret_cast<const uint32_t*>(xl + off));
        *reinterpret_cast<float2*>(x_cls + static_cast<size_t>(img) * D + h * 64 + lane * 2) = make_float2(vh.x + vl.x, vh.y + vl.y);
    }
    {
        float2 qf = unpack_op16x2(*reinterpret_cast<const uint32_t*>(base + lane * 2));
        if (split) {
            const float2 ql = unpack_op16x2(*reinterpret_cast<const uint32_t*>(base_lo + lane * 2));
            qf.x += ql.x;
            qf.y += ql.y;
        }
        s_q[warp][lane * 2] = qf.x;
        s_q[warp][lane * 2 + 1] = qf.y;
    }
    __syncwarp();

    // scores: lane j handles keys j, j + 32, ...
    float mx = -INFINITY;
    for (int j = lane; j < S; j += 32) {
        const uint4* kr = reinterpret_cast<const uint4*>(base + static_cast<size_t>(j) * ld + D);
        const uint4* krl = split ? reinterpret_cast<const uint4*>(base_lo + static_cast<size_t>(j) * ld + D) : nullptr;
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint4 u = __ldg(kr + c);
            const uint4 ul = split ? __ldg(krl + c) : make_uint4(0u, 0u, 0u, 0u);
            const uint32_t* hp = reinterpret_cast<const uint32_t*>(&u);
            const uint32_t* lp = reinterpret_cast<const uint32_t*>(&ul);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float2 kf = unpack_op16x2(hp[e]);
                const float2 kl = unpack_op16x2(lp[e]);      // zero in the fp16 mode
                kf.x += kl.x;
                kf.y += kl.y;
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
    const op16_t* vbase_lo = split ? base_lo + 2 * D + lane * 2 : nullptr;
    for (int j = 0; j < S; ++j) {
        float2 vf = unpack_op16x2(*reinterpret_cast<const uint32_t*>(vbase + static_cast<size_t>(j) * ld));
        if (split) {
            const float2 vl = unpack_op16x2(*reinterpret_cast<const uint32_t*>(vbase_lo + static_cast<size_t>(j) * ld));
            vf.x += vl.x;
            vf.y += vl.y;
        }
        const float p = s_p[warp][j];
        o0 = fmaf(p, vf.x, o0);
        o1 = fmaf(p, vf.y, o1);
    }
    const float inv = 1.0f / sum;
    const uint32_t hi = pack_op16x2(o0 * inv, o1 * inv);
    *reinterpret_cast<uint32_t*>(o_cls + static_cast<size_t>(img) * D + h * 64 + lane * 2) = hi;
    if (split) {
        const float2 hf = unpack_op16x2(hi);
        *reinterpret_cast<uint32_t*>(o_cls_lo + static_cast<size_t>(img) * D + h * 64 + lane * 2) = pack_op16x2(o0 * inv - hf.x, o1 * inv - hf.y);
    }
}

}  // namespace mcm
