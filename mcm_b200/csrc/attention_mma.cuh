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
#pragma once
#include "ptx.cuh"

namespace mcm {

constexpr int kAttnDh = 64;
constexpr int kAttnLd = 72;      // smem row stride in fp16 (144 B): ldmatrix rows fall in distinct banks
constexpr int kAttnChunk = 64;   // keys per online-softmax step

__global__ void __launch_bounds__(288)
attention_mma_kernel(const op16_t* __restrict__ qkv, op16_t* __restrict__ out, int S, int H,
                     int keys_pad /* S rounded up to 16 */, float scale_log2e) {
    extern __shared__ __align__(16) uint8_t attn_smem[];
    op16_t* sK = reinterpret_cast<op16_t*>(attn_smem);
    op16_t* sV = sK + static_cast<size_t>(keys_pad) * kAttnLd;

    pdl_launch_dependents();
    pdl_wait();
    const int h = blockIdx.x % H;
    const int img = blockIdx.x / H;
    const int ld = 3 * H * kAttnDh;
    const op16_t* base = qkv + static_cast<size_t>(img) * S * ld + h * kAttnDh;
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
        } else {
            *reinterpret_cast<uint4*>(dk) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(dv) = make_uint4(0, 0, 0, 0);
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
                    uint32_t pf[4];
                    pf[0] = pack_op16x2(s[kp * 2][0], s[kp * 2][1]);
                    pf[1] = pack_op16x2(s[kp * 2][2], s[kp * 2][3]);
                    pf[2] = pack_op16x2(s[kp * 2 + 1][0], s[kp * 2 + 1][1]);
                    pf[3] = pack_op16x2(s[kp * 2 + 1][2], s[kp * 2 + 1][3]);
#pragma unroll
                    for (int dp = 0; dp < 4; ++dp) {   // pairs of 8-wide dh tiles
                        // ldmatrix x4 trans: (keys 0-7, dh 0-7), (keys 8-15, dh 0-7), (keys 0-7, dh 8-15), (keys 8-15, dh 8-15)
                        const int r = kc + kp * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
                        const int c = dp * 16 + ((lane >> 4) << 3);
                        uint32_t vf[4];
                        ldmatrix_x4_trans(vf, smem_u32(sV + r * kAttnLd + c));
                        mma_op16_16816(o[dp * 2], pf, vf[0], vf[1]);
                        mma_op16_16816(o[dp * 2 + 1], pf, vf[2], vf[3]);
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
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int c = nt * 8 + tq * 2;
            if (row0 < S)
                *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(row0) * (H * kAttnDh) + c) =
                    pack_op16x2(o[nt][0] * i0, o[nt][1] * i0);
            if (row1 < S)
                *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(row1) * (H * kAttnDh) + c) =
                    pack_op16x2(o[nt][2] * i1, o[nt][3] * i1);
        }
    }
}

}  // namespace mcm
