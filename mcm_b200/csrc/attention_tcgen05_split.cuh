// Split-row variant of the tcgen05 attention kernel (attention_tcgen05.cuh) for keys_pad <= 208
// (ViT-B/16: 197 tokens, ViT-B/32: 50).  Same roles, pipeline and TMEM buffers; the difference is the
// softmax stage: every query row is shared by TWO threads (two warps on the same TMEM lane
// quadrant), each owning half of the key columns, so the latency chain TMEM-load -> max -> exp2 ->
// TMEM-store per unit is half as long and 16 instead of 8 warps keep the MUFU pipe fed.  Row max and
// row sum are combined through shared memory with a 64-thread named barrier per warp pair.
//
// TMEM map of buffer g (base = g * 256 columns), nb = keys_pad / 16 key blocks, nb0 = ceil(nb / 2):
//   S  fp32 [0, keys_pad)
//   P  fp16 key blocks [0, nb0)  -> columns [0, 8 nb0)         (written by half 0 over columns it has read)
//           key blocks [nb0, nb) -> columns [208, 208 + 8 (nb - nb0))   (the 48 columns S never uses)
//   O  fp32 [128, 192)   (dead S columns by the time P.V runs)
#pragma once
#include "attention_tcgen05.cuh"

namespace mcm {

constexpr int kAtsThreads = 64 + 16 * 32;   // producer + MMA warps, 16 softmax warps
constexpr int kAtsStagingBytes = 16 * 32 * 64;   // 16 warps x 32 rows x 32 fp16

__host__ __device__ inline int ats_smem_bytes(int keys_pad) {
    return kAtcQStages * kAtcQBytes + 2 * 2 * keys_pad * 128 + kAtsStagingBytes + 4096 /*max + sum exchange*/ +
           1024 /*barriers*/ + 1024 /*align*/;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(kAtsThreads, 1)
attention_tcgen05_split_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                               const AtcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int kv_bytes = p.keys_pad * 128;
    uint8_t* s_q = smem;
    uint8_t* s_kv = smem + kAtcQStages * kAtcQBytes;
    uint8_t* s_stage = s_kv + 4 * kv_bytes;
    float* s_xch = reinterpret_cast<float*>(s_stage + kAtsStagingBytes);   // [2 kinds][2 buffers][2 halves][128 rows]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_stage + kAtsStagingBytes + 4096);
    uint64_t* q_full = bars;
    uint64_t* q_empty = bars + kAtcQStages;
    uint64_t* kv_full = bars + 2 * kAtcQStages;
    uint64_t* kv_empty = kv_full + 2;
    uint64_t* s_full = kv_full + 4;
    uint64_t* p_full = kv_full + 6;
    uint64_t* o_full = kv_full + 8;
    uint64_t* s_free = kv_full + 10;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(kv_full + 12);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_items = p.b * p.H;
    const int upi = p.units_per_item;
    const int D = p.H * 64;
    const int nb = p.keys_pad >> 4;       // 16-key blocks
    const int nb0 = (nb + 1) >> 1;        // blocks owned by column half 0

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_kv);
        for (int i = 0; i < kAtcQStages; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1);
            mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 8);
            mbar_init(&o_full[i], 1); mbar_init(&s_free[i], 8);
        }
        fence_barrier_init();
    }
    pdl_launch_dependents();
    if (warp == 1) tmem_alloc<512>(tmem_ptr);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();

    const int my_items = (n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const uint32_t n_units = static_cast<uint32_t>(my_items * upi);

    if (warp == 0) {
        if (elect_one()) {
            // ===== TMA producer (identical to the unsplit kernel) =====
            uint32_t ic = 0, uc = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ic) {
                const int img = item / p.H, h = item - img * p.H;
                const int row0 = img * p.S;
                const int kvs = ic & 1;
                mbar_wait(&kv_empty[kvs], ((ic >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&kv_full[kvs], 2 * kv_bytes);
                uint8_t* sk = s_kv + kvs * 2 * kv_bytes;
                tma_load_2d(sk, &tmap_kv, &kv_full[kvs], D + h * 64, row0);
                tma_load_2d(sk + kv_bytes, &tmap_kv, &kv_full[kvs], 2 * D + h * 64, row0);
                for (int mt = 0; mt < upi; ++mt, ++uc) {
                    const int qs = uc % kAtcQStages;
                    mbar_wait(&q_empty[qs], ((uc / kAtcQStages) & 1) ^ 1);
                    mbar_arrive_expect_tx(&q_full[qs], kAtcQBytes);
                    tma_load_2d(s_q + qs * kAtcQBytes, &tmap_q, &q_full[qs], h * 64, row0 + mt * 128);
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // ===== MMA issuer =====
            const uint32_t idesc_qk = make_idesc_f16(128, static_cast<uint32_t>(p.keys_pad));
            const uint32_t idesc_pv = make_idesc_f16(128, 64, /*a_mn_major=*/0, /*b_mn_major=*/1);
            auto issue_qk = [&](uint32_t v) {
                const uint32_t iv = v / upi;
                const int kvs = iv & 1, qs = v % kAtcQStages, buf = v & 1;
                mbar_wait(&kv_full[kvs], (iv >> 1) & 1);
                mbar_wait(&q_full[qs], (v / kAtcQStages) & 1);
                tcgen05_fence_after();
                const uint64_t adesc = make_smem_desc_sw128(smem_u32(s_q + qs * kAtcQBytes), 16, 1024);
                const uint64_t bdesc = make_smem_desc_sw128(smem_u32(s_kv + kvs * 2 * kv_bytes), 16, 1024);
                const uint32_t d = tmem_base + buf * 256;
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16(d, adesc + 2 * k, bdesc + 2 * k, idesc_qk, k != 0);
                umma_commit(&q_empty[qs]);
                umma_commit(&s_full[buf]);
            };
            if (n_units > 0) issue_qk(0);
            if (n_units > 1) issue_qk(1);
            for (uint32_t u = 0; u < n_units; ++u) {
                const int buf = u & 1;
                const uint32_t iu = u / upi;
                const int kvs = iu & 1;
                mbar_wait(&p_full[buf], (u >> 1) & 1);
                tcgen05_fence_after();
                const uint64_t vdesc = make_smem_desc_sw128(smem_u32(s_kv + kvs * 2 * kv_bytes + kv_bytes), 1024, 1024);
                const uint32_t base = tmem_base + buf * 256;
                for (int k = 0; k < nb; ++k) {
                    const uint32_t a = (k < nb0) ? base + 8 * k : base + 208 + 8 * (k - nb0);
                    umma_f16_ts(base + 128, a, vdesc + 128ull * k, idesc_pv, k != 0);
                }
                umma_commit(&o_full[buf]);
                if ((u + 1) % upi == 0) umma_commit(&kv_empty[kvs]);
                if (u + 2 < n_units) {
                    mbar_wait(&s_free[buf], (u >> 1) & 1);
                    tcgen05_fence_after();
                    issue_qk(u + 2);
                }
            }
        }
    } else {
        // ===== softmax / epilogue: 16 warps = 2 buffers x 2 column halves x 4 lane quadrants =====
        const int sw = warp - 2;
        const int g = sw >> 3;                 // TMEM buffer
        const int hf = (sw >> 2) & 1;          // column half
        const int quad = warp & 3;             // TMEM lane quadrant
        const int pair_bar = 1 + g * 4 + quad; // named barrier shared with the warp owning the other half
        const uint32_t t_lane = static_cast<uint32_t>(quad * 32) << 16;
        const uint32_t t_s = tmem_base + t_lane + g * 256;
        const uint32_t stg = smem_u32(s_stage + sw * 32 * 64);
        const int rloc = quad * 32 + lane;     // row within the 128-row unit
        float* x_max = s_xch + (g * 2) * 128;             // [half][row]
        float* x_sum = s_xch + 512 + (g * 2) * 128;
        const int blk_lo = hf ? nb0 : 0, blk_hi = hf ? nb : nb0;
        const float c = p.scale_log2e;
        // Stagger the two groups by half a period: group 1 starts its first softmax only when group 0 has
        // finished its first one, so from then on one group is in the MUFU-bound softmax while the tensor
        // core serves the other group's P.V / next QK^T (the MMA warp issues in exactly that order).
#ifndef MCM_ATC_NO_STAGGER
        if (g == 1 && n_units > 1) mbar_wait(&p_full[0], 0);
#endif
        for (uint32_t u = g; u < n_units; u += 2) {
            const uint32_t j = u >> 1;
            const uint32_t iu = u / upi;
            const int mt = static_cast<int>(u - iu * upi);
            const int item = static_cast<int>(blockIdx.x) + static_cast<int>(iu) * static_cast<int>(gridDim.x);
            const int img = item / p.H, h = item - img * p.H;
            const int wrow0 = mt * 128 + quad * 32;
            const bool warp_valid = wrow0 < p.S;     // identical for both warps of a pair
            mbar_wait(&s_full[g], j & 1);
            tcgen05_fence_after();
            float inv = 1.f;
            if (warp_valid) {
                // ---- pass 1: partial row max over this half's key blocks (two TMEM loads in flight) ----
                float mx = -INFINITY;
                for (int blk = blk_lo; blk < blk_hi; blk += 2) {
                    uint32_t va[16], vb[16];
                    const bool two = blk + 1 < blk_hi;
                    tmem_ld_32x32b_x16(t_s + blk * 16, va);
                    if (two) tmem_ld_32x32b_x16(t_s + blk * 16 + 16, vb);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        if (blk * 16 + e < p.S) mx = fmaxf(mx, __uint_as_float(va[e]));
                    if (two) {
#pragma unroll
                        for (int e = 0; e < 16; ++e)
                            if (blk * 16 + 16 + e < p.S) mx = fmaxf(mx, __uint_as_float(vb[e]));
                    }
                }
                x_max[hf * 128 + rloc] = mx;
                named_bar_sync(pair_bar, 64);
                mx = fmaxf(mx, x_max[(hf ^ 1) * 128 + rloc]);     // finite: key 0 is always valid and lives in half 0
                const float mc = mx * c;
                // ---- pass 2: p = 2^(s * c - max * c) -> fp16 pairs -> TMEM ----
                float sum0 = 0.f, sum1 = 0.f;
                for (int blk = blk_lo; blk < blk_hi; ++blk) {
                    uint32_t v[16];
                    tmem_ld_32x32b_x16(t_s + blk * 16, v);
                    tmem_ld_wait();
                    uint32_t pk[8];
                    if (blk * 16 + 16 <= p.S) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * e]), c, -mc));
                            const float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * e + 1]), c, -mc));
                            sum0 += p0;
                            sum1 += p1;
                            pk[e] = pack_op16x2(p0, p1);
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int k0 = blk * 16 + 2 * e;
                            const float p0 = (k0 < p.S) ? ex2_approx(fmaf(__uint_as_float(v[2 * e]), c, -mc)) : 0.f;
                            const float p1 = (k0 + 1 < p.S) ? ex2_approx(fmaf(__uint_as_float(v[2 * e + 1]), c, -mc)) : 0.f;
                            sum0 += p0;
                            sum1 += p1;
                            pk[e] = pack_op16x2(p0, p1);
                        }
                    }
                    tmem_st_32x32b_x8(hf ? t_s + 208 + 8 * (blk - nb0) : t_s + 8 * blk, pk);
                }
                tmem_st_wait();
                x_sum[hf * 128 + rloc] = sum0 + sum1;
                named_bar_sync(pair_bar, 64);
                inv = 1.0f / (sum0 + sum1 + x_sum[(hf ^ 1) * 128 + rloc]);
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[g]);

            // ---- O = P V by the tensor core; this warp drains 32 of the 64 head dims ----
            mbar_wait(&o_full[g], j & 1);
            tcgen05_fence_after();
            if (warp_valid) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(t_s + 128 + hf * 32, v);
                tmem_ld_wait();
                const int swz = (lane >> 1) & 3;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4 w;
                    w.x = pack_op16x2(__uint_as_float(v[8 * q + 0]) * inv, __uint_as_float(v[8 * q + 1]) * inv);
                    w.y = pack_op16x2(__uint_as_float(v[8 * q + 2]) * inv, __uint_as_float(v[8 * q + 3]) * inv);
                    w.z = pack_op16x2(__uint_as_float(v[8 * q + 4]) * inv, __uint_as_float(v[8 * q + 5]) * inv);
                    w.w = pack_op16x2(__uint_as_float(v[8 * q + 6]) * inv, __uint_as_float(v[8 * q + 7]) * inv);
                    sts_v4u(stg + lane * 64 + ((q ^ swz) << 4), w);
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[g]);
            if (warp_valid) {
                const int slot = lane & 3;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = (lane >> 2) + 8 * i;
                    const int row = wrow0 + r;
                    if (row < p.S) {
                        const uint4 w = lds_v4u(stg + r * 64 + ((slot ^ ((r >> 1) & 3)) << 4));
                        *reinterpret_cast<uint4*>(p.out + (static_cast<size_t>(img) * p.S + row) * D + h * 64 + hf * 32 + slot * 8) = w;
                    }
                }
            }
            __syncwarp();
        }
    }

    __syncwarp();
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tcgen05_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace mcm
