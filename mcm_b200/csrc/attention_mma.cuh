// Multi-head self-attention core, first version: softmax(Q K^T / sqrt(64)) V per (image, head),
// non-causal, no mask (HF:modeling_clip.py:261-279 eager == what SDPA computes, :318-331).
//
// 4 % of the tower's FLOPs.  This version keeps the whole K and V of one head in
// shared memory (S <= 272 keys x 64 x fp16 = 34 KB each), gives every warp 16-query-row tiles, and
// runs flash-style online softmax over 64-key chunks with warp-level mma.sync tiles
// (m16n8k16 fp16 -> fp32).  Scores, running max / sum and the output accumulator stay in fp32
// registers; probabilities are rounded to fp16 only as the A operand of P.V.
//
// qkv: fp16 [b * S, 3 * H * 64]  (row = token; [q | k | v], head h at columns h * 64 of each part)
// out: fp16 [b * S, H * 64]      (== attn_output.transpose(1,2).reshape(B,S,D), HF:333)
//
// SPLIT = true is the attention of the split-fp16 precision mode (gemm_tcgen05.cuh, "Precision modes"): q, k, v and
// the output are (hi, lo) fp16 pairs, the probabilities are split the same way, and both products are three-term sums
//     S = Qh Kh^T + Ql Kh^T + Qh Kl^T,      O = Ph Vh + Pl Vh + Ph Vl
// on the same fp32 mma.sync accumulators; exp2f instead of ex2.approx.  3x the tensor work of the fp16 kernel, fp32-class results.
#pragma once
#include "ptx.cuh"
#include "gemm_tcgen05.cuh"   // split_op16x2

namespace mcm {

constexpr int kAttnDh = 64;
constexpr int kAttnLd = 72;      // smem row stride in fp16 (144 B): ldmatrix rows fall in distinct banks
constexpr int kAttnChunk = 64;   // keys per online-softmax step

template <bool SPLIT>
__global__ void __launch_bounds__(288)
attention_mma_kernel(const op16_t* __restrict__ qkv, const op16_t* __restrict__ qkv_lo, op16_t* __restrict__ out,
                     op16_t* __restrict__ out_lo, int S, int H, int keys_pad /* S rounded up to 16 */, float scale_log2e) {
    extern __shared__ __align__(16) uint8_t attn_smem[];
    op16_t* sK = reinterpret_cast<op16_t*>(attn_smem);
    op16_t* sV = sK + static_cast<size_t>(keys_pad) * kAttnLd;
    op16_t* sKl = sV + static_cast<size_t>(keys_pad) * kAttnLd;    // SPLIT only: low halves of K and V
    op16_t* sVl = sKl + static_cast<size_t>(keys_pad) * kAttnLd;

    pdl_launch_dependents();
    pdl_wait();
    const int h = blockIdx.x % H;
    const int img = blockIdx.x / H;
    const int ld = 3 * H * kAttnDh;
    const op16_t* base = qkv + static_cast<size_t>(img) * S * ld + h * kAttnDh;
    const op16_t* base_lo = SPLIT ? qkv_lo + static_cast<size_t>(img) * S * ld + h * kAttnDh : nullptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;

    // ---- stage K and V of this head (zero rows beyond S) ----
    for (int t = threadIdx.x; t < keys_pad * 8; t += blockDim.x) {
        const int r = t >> 3, c = (t & 7) * 8;
        op16_t* dk = sK + r * kAttnLd + c;
        op16_t* dv = sV + r * kAttnLd + c;
        if (r < S) {
            const op16_t* src = base + static_cast<size_t>(r) * ld + c;
            cp_async_16(dk, src + H * kAttnDh);
            cp_async_16(dv, src + 2 * H * kAttnDh);
            if constexpr (SPLIT) {
                const op16_t* srl = base_lo + static_cast<size_t>(r) * ld + c;
                cp_async_16(sKl + r * kAttnLd + c, srl + H * kAttnDh);
                cp_async_16(sVl + r * kAttnLd + c, srl + 2 * H * kAttnDh);
            }
        } else {
            *reinterpret_cast<uint4*>(dk) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(dv) = make_uint4(0, 0, 0, 0);
            if constexpr (SPLIT) {
                *reinterpret_cast<uint4*>(sKl + r * kAttnLd + c) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4*>(sVl + r * kAttnLd + c) = make_uint4(0, 0, 0, 0);
            }
        }
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    const int g = lane >> 2;   // row within the 8-row group
    const int tq = lane & 3;   // column pair
    const int mtiles = (S + 15) >> 4;

    for (int mt = warp; mt < mtiles; mt += nwarps) {
        const int row0 = mt * 16 + g;   // this thread's rows: row0 and row0 + 8
        const int row1 = row0 + 8;

        // Q fragments straight from global: 4 k-steps (dh = 64) x 4 regs
        uint32_t qf[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const int c = ks * 16 + tq * 2;
            const uint32_t* p0 = reinterpret_cast<const uint32_t*>(base + static_cast<size_t>(row0) * ld + c);
            const uint32_t* p1 = reinterpret_cast<const uint32_t*>(base + static_cast<size_t>(row1) * ld + c);
            qf[ks][0] = row0 < S ? __ldg(p0) : 0u;
            qf[ks][1] = row1 < S ? __ldg(p1) : 0u;
            qf[ks][2] = row0 < S ? __ldg(p0 + 4) : 0u;
            qf[ks][3] = row1 < S ? __ldg(p1 + 4) : 0u;
        }
        uint32_t ql[SPLIT ? 4 : 1][4];
        if constexpr (SPLIT) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const int c = ks * 16 + tq * 2;
                const uint32_t* p0 = reinterpret_cast<const uint32_t*>(base_lo + static_cast<size_t>(row0) * ld + c);
                const uint32_t* p1 = reinterpret_cast<const uint32_t*>(base_lo + static_cast<size_t>(row1) * ld + c);
                ql[ks][0] = row0 < S ? __ldg(p0) : 0u;
                ql[ks][1] = row1 < S ? __ldg(p1) : 0u;
                ql[ks][2] = row0 < S ? __ldg(p0 + 4) : 0u;
                ql[ks][3] = row1 < S ? __ldg(p1 + 4) : 0u;
            }
        }

        float o[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
        float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

        for (int kc = 0; kc < keys_pad; kc += kAttnChunk) {
            const int nkt = min(kAttnChunk, keys_pad - kc) >> 3;   // 8-key tiles in this chunk (even)
            float s[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;

            // ---- S = Q K^T ----
#pragma unroll
            for (int np = 0; np < 4; ++np) {       // pairs of key tiles (16 keys)
                if (np * 2 < nkt) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        // ldmatrix x4: matrices (keys 0-7, dh 0-7), (keys 0-7, dh 8-15), (keys 8-15, dh 0-7), (keys 8-15, dh 8-15)
                        const int r = kc + np * 16 + (lane & 7) + ((lane >> 4) << 3);
                        const int c = ks * 16 + (((lane >> 3) & 1) << 3);
                        uint32_t kf[4];
                        ldmatrix_x4(kf, smem_u32(sK + r * kAttnLd + c));
                        mma_op16_16816(s[np * 2], qf[ks], kf[0], kf[1]);
                        mma_op16_16816(s[np * 2 + 1], qf[ks], kf[2], kf[3]);
                        if constexpr (SPLIT) {
                            mma_op16_16816(s[np * 2], ql[ks], kf[0], kf[1]);
                            mma_op16_16816(s[np * 2 + 1], ql[ks], kf[2], kf[3]);
                            uint32_t kl[4];
                            ldmatrix_x4(kl, smem_u32(sKl + r * kAttnLd + c));
                            mma_op16_16816(s[np * 2], qf[ks], kl[0], kl[1]);
                            mma_op16_16816(s[np * 2 + 1], qf[ks], kl[2], kl[3]);
                        }
                    }
                }
            }

            // ---- mask padded keys, online softmax ----
            float cm0 = -INFINITY, cm1 = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const int key = kc + nt * 8 + tq * 2;
                if (nt >= nkt || key >= S) s[nt][0] = s[nt][2] = -INFINITY;
                if (nt >= nkt || key + 1 >= S) s[nt][1] = s[nt][3] = -INFINITY;
                cm0 = fmaxf(cm0, fmaxf(s[nt][0], s[nt][1]));
                cm1 = fmaxf(cm1, fmaxf(s[nt][2], s[nt][3]));
            }
            cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1));
            cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
            cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1));
            cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
            const float mn0 = fmaxf(m0, cm0), mn1 = fmaxf(m1, cm1);   // finite: every chunk holds a valid key
            const float a0 = exp2f((m0 - mn0) * scale_log2e), a1 = exp2f((m1 - mn1) * scale_log2e);
            m0 = mn0;
            m1 = mn1;
            const float ms0 = mn0 * scale_log2e, ms1 = mn1 * scale_log2e;
            float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                s[nt][0] = exp2f(s[nt][0] * scale_log2e - ms0);
                s[nt][1] = exp2f(s[nt][1] * scale_log2e - ms0);
                s[nt][2] = exp2f(s[nt][2] * scale_log2e - ms1);
                s[nt][3] = exp2f(s[nt][3] * scale_log2e - ms1);
                rs0 += s[nt][0] + s[nt][1];
                rs1 += s[nt][2] + s[nt][3];
            }
            l0 = l0 * a0 + rs0;
            l1 = l1 * a1 + rs1;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                o[i][0] *= a0; o[i][1] *= a0;
                o[i][2] *= a1; o[i][3] *= a1;
            }

            // ---- O += P V ----
#pragma unroll
            for (int kp = 0; kp < 4; ++kp) {       // 16-key k-steps
                if (kp * 2 < nkt) {
                    uint32_t pf[4], pl[SPLIT ? 4 : 1];
                    if constexpr (SPLIT) {
                        split_op16x2(s[kp * 2][0], s[kp * 2][1], pf[0], pl[0]);
                        split_op16x2(s[kp * 2][2], s[kp * 2][3], pf[1], pl[1]);
                        split_op16x2(s[kp * 2 + 1][0], s[kp * 2 + 1][1], pf[2], pl[2]);
                        split_op16x2(s[kp * 2 + 1][2], s[kp * 2 + 1][3], pf[3], pl[3]);
                    } else {
                        pf[0] = pack_op16x2(s[kp * 2][0], s[kp * 2][1]);
                        pf[1] = pack_op16x2(s[kp * 2][2], s[kp * 2][3]);
                        pf[2] = pack_op16x2(s[kp * 2 + 1][0], s[kp * 2 + 1][1]);
                        pf[3] = pack_op16x2(s[kp * 2 + 1][2], s[kp * 2 + 1][3]);
                    }
#pragma unroll
                    for (int dp = 0; dp < 4; ++dp) {   // pairs of 8-wide dh tiles
                        // ldmatrix x4 trans: (keys 0-7, dh 0-7), (keys 8-15, dh 0-7), (keys 0-7, dh 8-15), (keys 8-15, dh 8-15)
                        const int r = kc + kp * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
                        const int c = dp * 16 + ((lane >> 4) << 3);
                        uint32_t vf[4];
                        ldmatrix_x4_trans(vf, smem_u32(sV + r * kAttnLd + c));
                        mma_op16_16816(o[dp * 2], pf, vf[0], vf[1]);
                        mma_op16_16816(o[dp * 2 + 1], pf, vf[2], vf[3]);
                        if constexpr (SPLIT) {
                            mma_op16_16816(o[dp * 2], pl, vf[0], vf[1]);
                            mma_op16_16816(o[dp * 2 + 1], pl, vf[2], vf[3]);
                            uint32_t vl[4];
                            ldmatrix_x4_trans(vl, smem_u32(sVl + r * kAttnLd + c));
                            mma_op16_16816(o[dp * 2], pf, vl[0], vl[1]);
                            mma_op16_16816(o[dp * 2 + 1], pf, vl[2], vl[3]);
                        }
                    }
                }
            }
        }

        // ---- normalise and store ----
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float i0 = 1.0f / l0, i1 = 1.0f / l1;
        op16_t* ob = out + static_cast<size_t>(img) * S * (H * kAttnDh) + h * kAttnDh;
        op16_t* obl = SPLIT ? out_lo + static_cast<size_t>(img) * S * (H * kAttnDh) + h * kAttnDh : nullptr;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int c = nt * 8 + tq * 2;
            uint32_t h0, h1, l0w = 0u, l1w = 0u;
            if constexpr (SPLIT) {
                split_op16x2(o[nt][0] * i0, o[nt][1] * i0, h0, l0w);
                split_op16x2(o[nt][2] * i1, o[nt][3] * i1, h1, l1w);
            } else {
                h0 = pack_op16x2(o[nt][0] * i0, o[nt][1] * i0);
                h1 = pack_op16x2(o[nt][2] * i1, o[nt][3] * i1);
            }
            if (row0 < S) {
                *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(row0) * (H * kAttnDh) + c) = h0;
                if constexpr (SPLIT) *reinterpret_cast<uint32_t*>(obl + static_cast<size_t>(row0) * (H * kAttnDh) + c) = l0w;
            }
            if (row1 < S) {
                *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(row1) * (H * kAttnDh) + c) = h1;
                if constexpr (SPLIT) *reinterpret_cast<uint32_t*>(obl + static_cast<size_t>(row1) * (H * kAttnDh) + c) = l1w;
            }
        }
    }
}

}  // namespace mcm
