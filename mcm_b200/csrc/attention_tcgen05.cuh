// Multi-head self-attention core on the 5th-gen tensor cores:  softmax(Q K^T / 8) V  per (image, head)
// for sequences of up to 257 tokens (ViT-B/16: 197, ViT-B/32: 50, ViT-L/14: 257), head dim 64, no mask, not
// causal (HF:modeling_clip.py:261-279 eager == what SDPA computes, :318-331).
//
// 257 = 256 + 1: a UMMA tile is at most 256 keys wide and two fp32 score buffers of more than 256 columns do
// not fit the 512 columns of tensor memory, so the ONE key beyond 256 of ViT-L/14 ("extra key") never enters
// the tensor core: every softmax thread computes its row's score against it with 64 FMAs (its q row from the Q
// tile, the key from a small "x box": the producer also loads rows 256.. of q / k / v, 8 rows x 128 B each, with
// every item), folds it into the row max / row sum, and adds p_extra * v_extra to the O row while draining it.
// (Reading q / k / v of that token from global memory instead put ~5 k cycles of load latency on every unit.)  The 256 other keys take the normal path with keys_pad = 256.  Likewise the one
// QUERY row beyond 256 does not get a (1 / 128 full) third unit: warp 10 computes it with warp-level mma.sync tiles
// from the K / V tiles already in shared memory, hidden behind the two real units.
//
// Persistent CTAs (one per SM) walk over (image, head) items; every item is cut into 128-query-row
// units.  Warp roles:
//   warp 0      TMA producer: K and V of the item ([keys_pad x 64] fp16 boxes, 128-byte swizzle, 2-stage
//               ring; 3-D tensor map over [image][token][3 D], so key rows beyond the image's S tokens are zero-filled)
//               and the Q tile of every unit (3-stage ring), straight out of the fused QKV buffer.
//   warp 1      MMA issuer (one elected thread):  S = Q K^T   as UMMA 128 x keys_pad x 16 (x4, SS mode),
//                                                 O = P V     as UMMA 128 x 64 x 16 (x keys_pad/16, TS mode:
//               P is read from TENSOR MEMORY, V from shared memory as an MN-major operand).
//   warps 2-9   two softmax groups of 4 warps; group g owns TMEM buffer g (256 columns), one thread per
//               query row: row max and exp2 in fp32 straight from TMEM (tcgen05.ld), P written back over
//               the dead S columns as fp16 (tcgen05.st), then O / rowsum -> fp16 -> smem transpose ->
//               coalesced 128-byte row stores.  While one group runs its softmax the tensor core serves
//               the other group's QK^T / PV, so the MUFU (exp2) pipe -- the real bound of this kernel:
//               2 x 128 x keys_pad exponentials per item vs 16 per clock per SM -- stays busy.
// TMEM map of buffer g (base = g * 256 columns):  S fp32 [0, keys_pad)  ->  P fp16 [0, keys_pad/2)
//                                                 O fp32 [128, 192)  (dead S columns by the time PV runs)
// Padded keys (>= S) are masked to probability 0; padded query rows are computed and never stored.
// Sequences of at most 64 tokens (ViT-B/32: 50) would fill 39 % of a 128-row unit: there a unit carries TWO items (pair_mode).
//
// Measured (clock64 phase trace of one CTA, ViT-B/16 shape, tools/attn_sweep.py with the -DMCM_ATC_TRACE
// build): one unit takes ~7.7 k cycles end to end -- row max 1.4 k, exp2 pass 3.1 k (the warp's own
// instruction stream, MUFU 53 % busy), and ~3.2 k of hand-offs (barrier hops, issuing 13 P.V UMMAs,
// draining O) -- with two units in flight per SM (TMEM holds two S buffers).  Tried and rejected: software-
// pipelined TMEM loads (+13 %), one pass per row with a lazily raised reference maximum (
// +6 %), two threads per row, one MMA issuer warp per buffer, forcing the two groups out of phase (all +-5 %), and
// moving the 69 rows beyond the first 128 of ViT-B/16 to four mma.sync warps so that an item needs ONE unit (2.2 x
// slower: a 16-row mma.sync tile over 208 keys takes one warp ~8 k cycles, five of them per item on four warps).
// Two more, both neutral or worse although each removes what a model of the kernel says is its bound: O in a TMEM tile
// outside the score buffers so that the next-but-one Q K^T is issued right behind P.V instead of after the O drain
// (97.7 vs 97.9 us), and keeping 1 / 2 / 3 of the seven 32-key score chunks in registers between the two softmax
// passes to relieve the 64 B/clk TMEM read port (245 KB per unit = 3.8 k cycles, the measured unit time): 106 / 112 /
// 123 us against 103 for the same code keeping none.  96 us per layer call vs 293 us for the
// mma.sync kernel; the next step is a third unit in flight (split the keys, rescale O in TMEM).
//
// qkv: fp16 [b * S, 3 * H * 64]  (row = token; [q | k | v], head h at columns h * 64 of each part)
// out: fp16 [b * S, H * 64]      (== attn_output.transpose(1,2).reshape(B,S,D), HF:333)
#pragma once
#include <cuda.h>
#include "ptx.cuh"

#ifndef MCM_ATC_SHARED_O
#define MCM_ATC_SHARED_O 1      // 0: O inside each score buffer for every shape (A/B builds)
#endif

namespace mcm {

constexpr int kAtcThreads = 352;          // 11 warps: TMA, MMA, 2 x 4 softmax, tail-row warp (idle unless S > 256)
constexpr int kAtcQStages = 3;
constexpr int kAtcQBytes = 128 * 128;          // 128 rows x 64 fp16
constexpr int kAtcStagingBytes = 8 * 32 * 128; // 8 softmax warps x 32 rows x 64 fp16
constexpr int kAtcXBytes = 3 * 1024;           // per K/V stage: q | k | v of tokens 256..263 (8 rows x 128 B each)

struct AtcParams {
    int b, S, H, keys_pad;   // keys_pad: min(S, 256) rounded up to 16
    int n_extra;             // S - 256 if S > 256 (0 or 1): keys handled outside the tensor core
    int pair_mode;           // S <= 64 (ViT-B/32): TWO (image, head) items share a unit -- rows / keys 0..63 item 2w, 64..127
                             // item 2w + 1, keys_pad = 128, a row's probabilities of the other item's keys are zero
    int units_per_item;      // ceil(min(S, 256) / 128)
    float inv_H;             // 1 / H (image = item / H without an integer division, see atc_div_h)
    float scale_log2e;       // dh^-0.5 * log2(e)
    op16_t* out;
    long long* trace;        // debug builds (-DMCM_ATC_TRACE): per-phase clock64 stamps of CTA 0, else unused
};

#ifdef MCM_ATC_TRACE
// trace[role][unit][event]: role 0 = MMA thread, 1 = softmax warp of group 0, 2 = softmax warp of group 1
#define ATC_TRACE(role, unit, ev)                                                              \
    do {                                                                                       \
        if (p.trace && blockIdx.x == 0 && (unit) < 16) p.trace[((role) * 16 + (unit)) * 8 + (ev)] = clock64(); \
    } while (0)
#else
#define ATC_TRACE(role, unit, ev) do {} while (0)
#endif

// item / H for item < 2^22 / H: (item + 0.5) / H lies at least 0.5 / H away from the next integer, float rounding (2^-23
// relative) cannot cross it.  An integer division costs ~40 instructions, and the softmax warps did five per unit.
__device__ __forceinline__ int atc_div_h(int item, float inv_H) { return __float2int_rz((static_cast<float>(item) + 0.5f) * inv_H); }
// unit -> item index: units per item is 1 or 2 (at most 256 tensor-core query rows)
__device__ __forceinline__ uint32_t atc_unit_item(uint32_t u, int upi) { return upi == 2 ? (u >> 1) : u; }
__device__ __forceinline__ bool atc_last_unit_of_item(uint32_t u, int upi) { return upi == 2 ? (u & 1) != 0 : true; }

constexpr int kAtcMaxS = 257;   // 256 tensor-core keys + 1 extra key
// keys the tensor core sees, rounded up to the UMMA N granularity
__host__ __device__ inline int atc_keys_pad(int S) { return ((S < 256 ? S : 256) + 15) / 16 * 16; }

// rows of the K / V TMA box: 64 in pair mode (S <= 64, two boxes make a 128-row tile), else keys_pad
__host__ __device__ inline int atc_kv_box_rows(int S) { return S <= 64 ? 64 : atc_keys_pad(S); }

__host__ __device__ inline int atc_smem_bytes(int keys_pad) {
    return kAtcQStages * kAtcQBytes + 2 * 2 * keys_pad * 128 + 2 * kAtcXBytes + kAtcStagingBytes + 1024 /*barriers*/ + 1024 /*align*/;
}

// 3D tiled store shared -> global (bulk async group of the issuing thread); elements outside the tensor are clipped
__device__ __forceinline__ void tma_store_3d(const void* tmap, uint32_t smem_src, int32_t c0, int32_t c1, int32_t c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void atc_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void atc_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void atc_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- extra tcgen05 PTX ----
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {   // one MUFU op; flushes denormal results to zero
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// dot product of two 64-element fp16 rows in shared memory, fp32 accumulation: row a is row `swz` (mod 8) of a
// 128-byte-swizzled tile (16-byte chunk j sits at j ^ swz), row b is row 0 of its tile (not swizzled)
__device__ __forceinline__ float atc_dot64(uint32_t a, int swz, uint32_t b) {
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint4 x = lds_v4u(a + ((j ^ swz) << 4)), y = lds_v4u(b + (j << 4));
        const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 u = unpack_op16x2(xs[e]), v = unpack_op16x2(ys[e]);
            acc0 = fmaf(u.x, v.x, acc0);
            acc1 = fmaf(u.y, v.y, acc1);
        }
    }
    return acc0 + acc1;
}

// running max over one 32-key chunk of a score row (keys k0 .. k0 + 31; keys >= S are padding)
__device__ __forceinline__ float atc_chunk_max(const uint32_t (&v)[32], int k0, int S, float mx) {
    if (k0 + 32 <= S) {
        // a tree instead of one running maximum: 3-input max instructions, dependency depth 5 instead of 16
        float t[11];
#pragma unroll
        for (int i = 0; i < 10; ++i)
            t[i] = fmaxf(fmaxf(__uint_as_float(v[3 * i]), __uint_as_float(v[3 * i + 1])), __uint_as_float(v[3 * i + 2]));
        t[10] = fmaxf(__uint_as_float(v[30]), __uint_as_float(v[31]));
        const float r0 = fmaxf(fmaxf(t[0], t[1]), t[2]), r1 = fmaxf(fmaxf(t[3], t[4]), t[5]);
        const float r2 = fmaxf(fmaxf(t[6], t[7]), t[8]), r3 = fmaxf(fmaxf(t[9], t[10]), mx);
        mx = fmaxf(fmaxf(r0, r1), fmaxf(r2, r3));
    } else {
#pragma unroll
        for (int e = 0; e < 32; ++e)
            if (k0 + e < S) mx = fmaxf(mx, __uint_as_float(v[e]));
    }
    return mx;
}
// p = 2^(s * c - mc) for one 32-key chunk, accumulated into two partial row sums and written to TMEM
// as 16 packed fp16 pairs (the A operand of P.V) at `t_p`
__device__ __forceinline__ void atc_chunk_exp(const uint32_t (&v)[32], int k0, int S, float c, float mc, float& sum0,
                                              float& sum1, uint32_t t_p) {
    uint32_t pk[16];
    if (k0 + 32 <= S) {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * e]), c, -mc));
            const float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * e + 1]), c, -mc));
            sum0 += p0;
            sum1 += p1;
            pk[e] = pack_op16x2(p0, p1);
        }
    } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int k = k0 + 2 * e;
            const float p0 = (k < S) ? ex2_approx(fmaf(__uint_as_float(v[2 * e]), c, -mc)) : 0.f;
            const float p1 = (k + 1 < S) ? ex2_approx(fmaf(__uint_as_float(v[2 * e + 1]), c, -mc)) : 0.f;
            sum0 += p0;
            sum1 += p1;
            pk[e] = pack_op16x2(p0, p1);
        }
    }
    tmem_st_32x32b_x16(t_p, pk);
}

// The two softmax passes of one warp (one query row per thread) over its score columns in tensor memory:
//   pass 1: row maximum (TMEM loads two chunks at a time so that their latencies overlap),
//   pass 2: p = 2^(s * c - max * c), written back over the first half of the score columns as packed fp16, row sum.
// NFULL >= 0: the number of 32-key chunks is a compile-time constant and every one of them is fully valid (the caller
// checked S_tc >= 32 * NFULL), so the loops unroll and the chunk addresses are immediates; REM16: a trailing 16-key chunk,
// the only place padded keys (>= S_tc) can sit.  NFULL < 0: run-time chunk count, any chunk may hold padded keys.
// s_x: raw score of the extra key (ViT-L/14's 257th token), -INFINITY if there is none; on return the probability of that
// key rounded like the P operand (0 if none).  Returns the row sum.
template <int NFULL, bool REM16>
__device__ __forceinline__ float atc_two_pass(uint32_t t_sr, uint32_t t_pw, int nfull_rt, bool rem16_rt, int S_tc, float c, float& s_x) {
    const int nfull = NFULL >= 0 ? NFULL : nfull_rt;
    const bool rem16 = NFULL >= 0 ? REM16 : rem16_rt;
    const int s_lim = NFULL >= 0 ? (1 << 30) : S_tc;       // compile-time shapes: full chunks hold valid keys only
    float mx = s_x;
#pragma unroll 1
    for (int ch = 0; ch < nfull; ch += 2) {
        uint32_t va[32], vb[32];
        const bool two = ch + 1 < nfull;
        tmem_ld_32x32b_x32(t_sr + ch * 32, va);
        if (two) tmem_ld_32x32b_x32(t_sr + ch * 32 + 32, vb);
        tmem_ld_wait();
        mx = atc_chunk_max(va, ch * 32, s_lim, mx);
        if (two) mx = atc_chunk_max(vb, ch * 32 + 32, s_lim, mx);
    }
    if (rem16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(t_sr + nfull * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e)
            if (nfull * 32 + e < S_tc) mx = fmaxf(mx, __uint_as_float(v[e]));
    }
    const float mc = mx * c;
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll (NFULL > 0 && NFULL % 3 == 0 ? 3 : 2)
    for (int ch = 0; ch < nfull; ++ch) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_sr + ch * 32, v);
        tmem_ld_wait();
        atc_chunk_exp(v, ch * 32, s_lim, c, mc, sum0, sum1, t_pw + ch * 16);
    }
    if (rem16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(t_sr + nfull * 32, v);
        tmem_ld_wait();
        uint32_t pk[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k0 = nfull * 32 + 2 * e;
            const float p0 = (k0 < S_tc) ? ex2_approx(fmaf(__uint_as_float(v[2 * e]), c, -mc)) : 0.f;
            const float p1 = (k0 + 1 < S_tc) ? ex2_approx(fmaf(__uint_as_float(v[2 * e + 1]), c, -mc)) : 0.f;
            sum0 += p0;
            sum1 += p1;
            pk[e] = pack_op16x2(p0, p1);
        }
        tmem_st_32x32b_x8(t_pw + nfull * 16, pk);
    }
    float sum = sum0 + sum1;
    if (s_x != -INFINITY) {   // the extra key: its probability, rounded like the P operand
        const float px = ex2_approx(fmaf(s_x, c, -mc));
        sum += px;
        s_x = unpack_op16x2(pack_op16x2(px, 0.f)).x;
    } else {
        s_x = 0.f;
    }
    return sum;
}

// ===== TMA producer (one elected thread of warp 0): K / V of every item, the Q tile of every unit =====
struct AtcSmem {
    uint8_t *s_q, *s_kv, *s_xbox;
    uint64_t *q_full, *q_empty, *kv_full, *kv_empty;
    int kv_bytes;
};
__device__ __forceinline__ void atc_producer(const CUtensorMap& tmap_q, const CUtensorMap& tmap_kv, const CUtensorMap& tmap_x,
                                             const AtcParams& p, const AtcSmem& sm) {
    uint8_t *const s_q = sm.s_q, *const s_kv = sm.s_kv, *const s_xbox = sm.s_xbox;
    uint64_t *const q_full = sm.q_full, *const q_empty = sm.q_empty, *const kv_full = sm.kv_full, *const kv_empty = sm.kv_empty;
    const int kv_bytes = sm.kv_bytes;
    const int n_items = p.b * p.H;
    const bool pair = p.pair_mode != 0;
    const int n_work = pair ? (n_items + 1) / 2 : n_items;
    const int upi = p.units_per_item;
    const int D = p.H * 64;
    uint32_t ic = 0, uc = 0;
    for (int item = blockIdx.x; item < n_work; item += gridDim.x, ++ic) {
        const int kvs = ic & 1;
        uint8_t* sk = s_kv + kvs * 2 * kv_bytes;
        if (pair) {
            // pair mode: 64-row boxes (tmap_kv) of item 2w and item 2w + 1 (the last item again if n_items is odd)
            // stacked into 128-row Q / K / V tiles; 8 KB per box keeps the 128-byte swizzle phase of the rows
            mbar_wait(&kv_empty[kvs], ((ic >> 1) & 1) ^ 1);
            mbar_arrive_expect_tx(&kv_full[kvs], 2 * kv_bytes);
            const int qs = uc % kAtcQStages;
            mbar_wait(&q_empty[qs], ((uc / kAtcQStages) & 1) ^ 1);
            mbar_arrive_expect_tx(&q_full[qs], kAtcQBytes);
            for (int half = 0; half < 2; ++half) {
                const int it2 = min(2 * item + half, n_items - 1);
                const int img2 = atc_div_h(it2, p.inv_H), h2 = it2 - img2 * p.H;
                tma_load_3d(sk + half * 8192, &tmap_kv, &kv_full[kvs], D + h2 * 64, 0, img2);
                tma_load_3d(sk + kv_bytes + half * 8192, &tmap_kv, &kv_full[kvs], 2 * D + h2 * 64, 0, img2);
                tma_load_3d(s_q + qs * kAtcQBytes + half * 8192, &tmap_kv, &q_full[qs], h2 * 64, 0, img2);
            }
            ++uc;
            continue;
        }
        const int img = atc_div_h(item, p.inv_H), h = item - img * p.H;
        const int row0 = img * p.S;
        mbar_wait(&kv_empty[kvs], ((ic >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[kvs], 2 * kv_bytes + (p.n_extra > 0 ? kAtcXBytes : 0));
        // tmap_kv is 3-D ([image][token][3 D]): the padded key rows S .. keys_pad - 1 lie outside the image's plane and
        // arrive as ZEROS -- never another image's (or a stale batch's) rows, whose Inf / NaN would leak through 0 * V
        tma_load_3d(sk, &tmap_kv, &kv_full[kvs], D + h * 64, 0, img);
        tma_load_3d(sk + kv_bytes, &tmap_kv, &kv_full[kvs], 2 * D + h * 64, 0, img);
        if (p.n_extra > 0) {
            uint8_t* sx = s_xbox + kvs * kAtcXBytes;
#pragma unroll
            for (int part = 0; part < 3; ++part)
                tma_load_2d(sx + part * 1024, &tmap_x, &kv_full[kvs], part * D + h * 64, row0 + 256);
        }
        for (int mt = 0; mt < upi; ++mt, ++uc) {
            const int qs = uc % kAtcQStages;
            mbar_wait(&q_empty[qs], ((uc / kAtcQStages) & 1) ^ 1);
            mbar_arrive_expect_tx(&q_full[qs], kAtcQBytes);
            tma_load_2d(s_q + qs * kAtcQBytes, &tmap_q, &q_full[qs], h * 64, row0 + mt * 128);
        }
    }
}

// ===== tail-row warp: query rows >= 256 (ViT-L/14: the one row 256) against all S keys =====
__device__ __forceinline__ void atc_tail_rows(const AtcParams& p, const AtcSmem& sm, int lane) {
    uint8_t *const s_kv = sm.s_kv, *const s_xbox = sm.s_xbox;
    uint64_t *const kv_full = sm.kv_full, *const kv_empty = sm.kv_empty;
    const int kv_bytes = sm.kv_bytes;
    const int n_items = p.b * p.H;
    const int D = p.H * 64;
    const float c = p.scale_log2e;
    const int tq = lane & 3;
    const bool row_lane = (lane >> 2) == 0;      // lanes 0..3 hold row 0 of the tile
    uint32_t ic = 0;
    auto lds32 = [](uint32_t a) { uint32_t w; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(a)); return w; };
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ic) {
        const int img = atc_div_h(item, p.inv_H), h = item - img * p.H;
        const int kvs = ic & 1;
        const uint32_t sk = smem_u32(s_kv + kvs * 2 * kv_bytes);
        const uint32_t sv = sk + kv_bytes;
        const uint32_t sx = smem_u32(s_xbox + kvs * kAtcXBytes);      // q | k | v of tokens 256.., row e swizzled by e
        mbar_wait(&kv_full[kvs], (ic >> 1) & 1);
        for (int r = 256; r < p.S; ++r) {
            const int e0 = r - 256;
            // A fragments: row 0 = the query row (x box, q part), rows 1..15 = 0
            uint32_t qf[4][4];
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const int col = ks * 16 + tq * 2;
                qf[ks][0] = row_lane ? lds32(sx + e0 * 128 + ((((col >> 3)) ^ e0) << 4) + (col & 7) * 2) : 0u;
                qf[ks][2] = row_lane ? lds32(sx + e0 * 128 + ((((col >> 3) + 1) ^ e0) << 4) + (col & 7) * 2) : 0u;
                qf[ks][1] = qf[ks][3] = 0u;
            }
            float o[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
            float m0 = -INFINITY, l0 = 0.f;
            for (int kc = 0; kc < p.keys_pad; kc += 64) {       // S > 256: keys_pad = 256, every chunk is full and valid
                float sc[8][4];
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
#pragma unroll
                for (int np = 0; np < 4; ++np) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        // ldmatrix x4: (keys 0-7, dh 0-7), (keys 0-7, dh 8-15), (keys 8-15, dh 0-7), (keys 8-15, dh 8-15)
                        const int kr = kc + np * 16 + (lane & 7) + ((lane >> 4) << 3);
                        const int ch = ks * 2 + ((lane >> 3) & 1);                 // 16-byte chunk of the key row
                        uint32_t kf[4];
                        ldmatrix_x4(kf, sk + kr * 128 + ((ch ^ (kr & 7)) << 4));
                        mma_op16_16816(sc[np * 2], qf[ks], kf[0], kf[1]);
                        mma_op16_16816(sc[np * 2 + 1], qf[ks], kf[2], kf[3]);
                    }
                }
                float cm = -INFINITY;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) cm = fmaxf(cm, fmaxf(sc[nt][0], sc[nt][1]));
                cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 1));
                cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 2));
                const float mn = fmaxf(m0, cm);
                const float a = ex2_approx((m0 - mn) * c);                         // 0 on the first chunk
                m0 = mn;
                const float ms = mn * c;
                float rs = 0.f;
                uint32_t pf[4][4];
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    const float p0 = ex2_approx(fmaf(sc[nt][0], c, -ms)), p1 = ex2_approx(fmaf(sc[nt][1], c, -ms));
                    rs += p0 + p1;
                    pf[nt >> 1][(nt & 1) * 2] = pack_op16x2(p0, p1);
                    pf[nt >> 1][(nt & 1) * 2 + 1] = 0u;                            // rows 8..15 of the tile
                }
                l0 = l0 * a + rs;
#pragma unroll
                for (int i = 0; i < 8; ++i) { o[i][0] *= a; o[i][1] *= a; }
#pragma unroll
                for (int kp = 0; kp < 4; ++kp) {
#pragma unroll
                    for (int dp = 0; dp < 4; ++dp) {
                        // ldmatrix x4 trans: (keys 0-7, dh 0-7), (keys 8-15, dh 0-7), (keys 0-7, dh 8-15), (keys 8-15, dh 8-15)
                        const int vr = kc + kp * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
                        const int ch = dp * 2 + (lane >> 4);
                        uint32_t vf[4];
                        ldmatrix_x4_trans(vf, sv + vr * 128 + ((ch ^ (vr & 7)) << 4));
                        mma_op16_16816(o[dp * 2], pf[kp], vf[0], vf[1]);
                        mma_op16_16816(o[dp * 2 + 1], pf[kp], vf[2], vf[3]);
                    }
                }
            }
            // the extra keys (tokens 256 + e, x box k / v parts), one at a time
            for (int e = 0; e < p.n_extra; ++e) {
                float sx_dot = 0.f;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const int col = ks * 16 + tq * 2;
                    const float2 q0 = unpack_op16x2(qf[ks][0]), q1 = unpack_op16x2(qf[ks][2]);
                    const float2 k0 = unpack_op16x2(lds32(sx + 1024 + e * 128 + (((col >> 3) ^ e) << 4) + (col & 7) * 2));
                    const float2 k1 = unpack_op16x2(lds32(sx + 1024 + e * 128 + ((((col >> 3) + 1) ^ e) << 4) + (col & 7) * 2));
                    sx_dot = fmaf(q0.x, k0.x, fmaf(q0.y, k0.y, fmaf(q1.x, k1.x, fmaf(q1.y, k1.y, sx_dot))));
                }
                sx_dot += __shfl_xor_sync(0xffffffffu, sx_dot, 1);
                sx_dot += __shfl_xor_sync(0xffffffffu, sx_dot, 2);
                const float mn = fmaxf(m0, sx_dot);
                const float a = ex2_approx((m0 - mn) * c);
                const float px = ex2_approx((sx_dot - mn) * c);
                m0 = mn;
                l0 = l0 * a + (tq == 0 ? px : 0.f);                                // l0 is summed over the quad below
                const float pxr = unpack_op16x2(pack_op16x2(px, 0.f)).x;           // rounded like the P operand
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    const int col = nt * 8 + tq * 2;
                    const float2 v = unpack_op16x2(lds32(sx + 2048 + e * 128 + (((col >> 3) ^ e) << 4) + (col & 7) * 2));
                    o[nt][0] = fmaf(pxr, v.x, o[nt][0] * a);
                    o[nt][1] = fmaf(pxr, v.y, o[nt][1] * a);
                }
            }
            l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
            l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
            const float inv = 1.0f / l0;
            if (row_lane) {
                op16_t* orow = p.out + (static_cast<size_t>(img) * p.S + r) * D + h * 64;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
                    *reinterpret_cast<uint32_t*>(orow + nt * 8 + tq * 2) = pack_op16x2(o[nt][0] * inv, o[nt][1] * inv);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&kv_empty[kvs]);     // this warp is done with the K / V stage
    }
}

// NFULL / REM16: compile-time chunk structure of the softmax passes (atc_two_pass): <6, true> ViT-B/16 (keys_pad 208, at
// least 192 valid keys), <8, false> ViT-L/14 (keys_pad 256), <-1, *> any other shape at run time (pair mode, tests).
template <int NFULL, bool REM16>
__global__ void __launch_bounds__(kAtcThreads, 1)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                         const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_o, const AtcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int kv_bytes = p.keys_pad * 128;                       // one K or V tile
    uint8_t* s_q = smem;                                         // [kAtcQStages][16 KB]
    uint8_t* s_kv = smem + kAtcQStages * kAtcQBytes;             // [2][K | V]
    uint8_t* s_xbox = s_kv + 4 * kv_bytes;                          // [2][q | k | v of tokens 256..263]  (S > 256 only)
    uint8_t* s_stage = s_xbox + 2 * kAtcXBytes;                     // [8 warps][32 rows][128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_stage + kAtcStagingBytes);
    uint64_t* q_full = bars;                       // [3]
    uint64_t* q_empty = bars + kAtcQStages;        // [3]
    uint64_t* kv_full = bars + 2 * kAtcQStages;    // [2]
    uint64_t* kv_empty = kv_full + 2;              // [2]
    uint64_t* s_full = kv_full + 4;                // [2]  MMA -> softmax group: S ready
    uint64_t* p_full = kv_full + 6;                // [2]  softmax group -> MMA: P written
    uint64_t* o_full = kv_full + 8;                // [2]  MMA -> softmax group: O ready
    uint64_t* s_free = kv_full + 10;               // [2]  softmax group -> MMA: O drained, buffer reusable
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(kv_full + 12);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_items = p.b * p.H;
    const bool pair = p.pair_mode != 0;
    const int n_work = pair ? (n_items + 1) / 2 : n_items;      // work items of the persistent loops: items, or pairs of items
    const int upi = p.units_per_item;
    const int D = p.H * 64;
    // TMEM map.  keys_pad <= 224: two score buffers back to back and ONE O tile behind them, shared by the two groups
    // (S / P of group g at g * keys_pad, O at 2 * keys_pad): P.V writes outside the score buffer, so the next-but-one
    // Q K^T is issued right behind P.V and the group drains O while the tensor core already computes its next S.
    // Wider score rows (ViT-L/14: 256 keys) leave no room: O stays inside each buffer (columns 128..191) and the next
    // Q K^T has to wait for the drain.
    const bool shared_o = MCM_ATC_SHARED_O && p.keys_pad <= 224;
    const uint32_t buf_stride = shared_o ? static_cast<uint32_t>(p.keys_pad) : 256u;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_kv);
        if (p.n_extra > 0) tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_o);
        // with an extra key the softmax warps read their q rows from the Q tile and k / v of the extra token from the
        // x box: they release the Q stage and the K / V stage together with the MMA thread's commits
        for (int i = 0; i < kAtcQStages; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], p.n_extra > 0 ? 5 : 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], p.n_extra > 0 ? 2 + 4 * upi : 1);   // + tail-row warp + 4 softmax warps per unit
            mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4);
            mbar_init(&o_full[i], 1); mbar_init(&s_free[i], 4);
        }
        fence_barrier_init();
    }
    pdl_launch_dependents();
    if (warp == 1) tmem_alloc<512>(tmem_ptr);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();   // everything above overlapped the predecessor's tail; global data is touched only below

    const AtcSmem sm{s_q, s_kv, s_xbox, q_full, q_empty, kv_full, kv_empty, kv_bytes};
    if (warp == 0) {
        if (elect_one()) atc_producer(tmap_q, tmap_kv, tmap_x, p, sm);
    } else if (warp == 1) {
        if (elect_one()) {
            // ===== MMA issuer =====
            const int my_items = (n_work - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
            const uint32_t n_units = static_cast<uint32_t>(my_items * upi);
            const uint32_t idesc_qk = make_idesc_f16(128, static_cast<uint32_t>(p.keys_pad));
            const uint32_t idesc_pv = make_idesc_f16(128, 64, /*a_mn_major=*/0, /*b_mn_major=*/1);
            const int ksteps = p.keys_pad >> 4;
            auto issue_qk = [&](uint32_t v) {
                const uint32_t iv = atc_unit_item(v, upi);                      // CTA-local item index of unit v
                const int kvs = iv & 1, qs = v % kAtcQStages, buf = v & 1;
                mbar_wait(&kv_full[kvs], (iv >> 1) & 1);
                mbar_wait(&q_full[qs], (v / kAtcQStages) & 1);
                tcgen05_fence_after();
                const uint64_t adesc = make_smem_desc_sw128(smem_u32(s_q + qs * kAtcQBytes), 16, 1024);
                const uint64_t bdesc = make_smem_desc_sw128(smem_u32(s_kv + kvs * 2 * kv_bytes), 16, 1024);
                const uint32_t d = tmem_base + buf * buf_stride;
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16(d, adesc + 2 * k, bdesc + 2 * k, idesc_qk, k != 0);
                umma_commit(&q_empty[qs]);
                umma_commit(&s_full[buf]);
                ATC_TRACE(0, v, 0);                       // QK^T of unit v issued
            };
            if (n_units > 0) issue_qk(0);
            if (n_units > 1) issue_qk(1);
            for (uint32_t u = 0; u < n_units; ++u) {
                const int buf = u & 1;
                const uint32_t iu = atc_unit_item(u, upi);
                const int kvs = iu & 1;
                mbar_wait(&p_full[buf], (u >> 1) & 1);
                ATC_TRACE(0, u, 1);                       // P of unit u ready
                tcgen05_fence_after();
                // V tile: [keys][64 dh] rows of 128 B = MN-major B operand; 16 keys (one UMMA K) = 2048 B
                const uint64_t vdesc = make_smem_desc_sw128(smem_u32(s_kv + kvs * 2 * kv_bytes + kv_bytes), 1024, 1024);
                if (shared_o && u >= 1) {     // the shared O tile: drained by the previous unit's group?
                    mbar_wait(&s_free[buf ^ 1], ((u - 1) >> 1) & 1);
                    ATC_TRACE(0, u - 1, 3);
                    tcgen05_fence_after();
                }
                const uint32_t a = tmem_base + buf * buf_stride;
                const uint32_t d = shared_o ? tmem_base + 2 * buf_stride : a + 128;
                for (int k = 0; k < ksteps; ++k) umma_f16_ts(d, a + 8 * k, vdesc + 128ull * k, idesc_pv, k != 0);
                umma_commit(&o_full[buf]);
                ATC_TRACE(0, u, 2);                       // P.V of unit u issued
                if (atc_last_unit_of_item(u, upi)) umma_commit(&kv_empty[kvs]);   // last unit of the item: K / V stage reusable
                if (u + 2 < n_units) {
                    if (!shared_o) {      // O lives inside the score buffer: wait for the drain
                        mbar_wait(&s_free[buf], (u >> 1) & 1);
                        ATC_TRACE(0, u, 3);                   // buffer of unit u drained
                        tcgen05_fence_after();
                    }
                    issue_qk(u + 2);      // the tensor core runs it behind P.V(u), which is the last reader of P(u)
                }
            }
        }
    } else if (warp == 10) {
        if (p.n_extra > 0) atc_tail_rows(p, sm, lane);
    } else {
        // ===== softmax / epilogue groups =====
        const int g = (warp - 2) >> 2;       // group = TMEM buffer
        const int quad = warp & 3;           // TMEM lane quadrant
        const uint32_t t_lane = static_cast<uint32_t>(quad * 32) << 16;
        const uint32_t t_s = tmem_base + t_lane + g * buf_stride;
        const uint32_t t_o = shared_o ? tmem_base + t_lane + 2 * buf_stride : t_s + 128;
        const uint32_t stg = smem_u32(s_stage + (warp - 2) * 32 * 128);   // shared-space address of this warp's staging tile
        const int my_items = (n_work - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
        const uint32_t n_units = static_cast<uint32_t>(my_items * upi);
        // pair mode: this warp's rows belong to item `half` of the pair and see only that item's 64 key columns
        const int half = pair ? (quad >> 1) : 0;
        const uint32_t t_sr = t_s + (pair ? half * 64 : 0);     // first score column of this warp's keys
        const uint32_t t_pw = t_s + (pair ? half * 32 : 0);     // first (packed fp16) probability column of those keys
        const int nfull = pair ? 2 : p.keys_pad >> 5;
        const bool rem16 = !pair && (p.keys_pad & 16) != 0;
        const float c = p.scale_log2e;
        const int S_tc = p.S - p.n_extra;    // keys that go through the tensor core
        // Stagger the two groups by half a period: group 1 starts its first softmax only when group 0 has
        // finished its first one, so from then on one group is in the MUFU-bound softmax while the tensor
        // core serves the other group's P.V / next QK^T (the MMA warp issues in exactly that order).
#ifndef MCM_ATC_NO_STAGGER
        if (g == 1 && n_units > 1) mbar_wait(&p_full[0], 0);
#endif
        for (uint32_t u = g; u < n_units; u += 2) {
            const uint32_t j = u >> 1;
            const uint32_t iu = atc_unit_item(u, upi);
            const int mt = static_cast<int>(u - iu * upi);
            const int work = static_cast<int>(blockIdx.x) + static_cast<int>(iu) * static_cast<int>(gridDim.x);
            const int item = pair ? 2 * work + half : work;
            const int img = atc_div_h(item, p.inv_H), h = item - img * p.H;
            const int wrow0 = pair ? (quad & 1) * 32 : mt * 128 + quad * 32;       // first query row (within the image) of this warp
            const bool warp_valid = item < n_items && wrow0 < p.S;
            // the extra key (ViT-L/14's 257th token): this row's raw score against it, on the CUDA cores,
            // while the tensor core is still busy with Q K^T of the other 256
            float s_x = 0.f;
            const uint32_t sx = smem_u32(s_xbox + (iu & 1) * kAtcXBytes);     // q | k | v of the extra token: row 0 of each 1 KB box
            mbar_wait(&s_full[g], j & 1);
            if (p.n_extra > 0) {
                // S ready: the MMA thread has seen this unit's Q tile and the item's K / V stage land; observing the
                // same (completed) phases here makes the TMA writes visible to this warp
                mbar_wait(&q_full[u % kAtcQStages], (u / kAtcQStages) & 1);
                mbar_wait(&kv_full[iu & 1], (iu >> 1) & 1);
                if (warp_valid)
                    s_x = atc_dot64(smem_u32(s_q + (u % kAtcQStages) * kAtcQBytes) + (quad * 32 + lane) * 128, lane & 7, sx + 1024);
                __syncwarp();
                if (lane == 0) mbar_arrive(&q_empty[u % kAtcQStages]);
            }
            if (quad == 2 && lane == 0) ATC_TRACE(1 + g, u, 0);       // S ready
            tcgen05_fence_after();
            float row_sum = 1.f;
            if (warp_valid) {
                if (p.n_extra == 0) s_x = -INFINITY;
                row_sum = atc_two_pass<NFULL, REM16>(t_sr, t_pw, nfull, rem16, S_tc, c, s_x);
                if (quad == 2 && lane == 0) ATC_TRACE(1 + g, u, 1);   // both passes issued
                if (pair) {     // probability 0 for the 64 keys of the other item (for half 0 these columns held this row's
                                // own scores 32..63: consumed by now)
                    uint32_t zero[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) zero[e] = 0u;
                    tmem_st_32x32b_x16(t_s + (1 - half) * 32, zero);
                    tmem_st_32x32b_x16(t_s + (1 - half) * 32 + 16, zero);
                }
                tmem_st_wait();
            }
            tcgen05_fence_before();
            __syncwarp();
            if (quad == 2 && lane == 0) ATC_TRACE(1 + g, u, 2);       // pass 2 done
            if (lane == 0) mbar_arrive(&p_full[g]);

            // ---- O = P V is computed by the tensor core; scale by 1 / rowsum and store ----
            mbar_wait(&o_full[g], j & 1);
            if (quad == 2 && lane == 0) ATC_TRACE(1 + g, u, 3);       // O ready
            tcgen05_fence_after();
            if (warp_valid) {
                const float inv = 1.0f / row_sum;
                // the bulk store of this warp's previous unit has read the staging tile (issued a whole unit ago)
                if (lane == 0) atc_store_wait_read();
                __syncwarp();
#pragma unroll
                for (int hc = 0; hc < 2; ++hc) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(t_o + hc * 32, v);
                    tmem_ld_wait();
                    if (p.n_extra > 0) {   // O += p_extra * v_extra (warp-uniform addresses: broadcast loads)
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint4 w = lds_v4u(sx + 2048 + ((hc * 4 + q) << 4));
                            const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 f = unpack_op16x2(ws[e]);
                                v[8 * q + 2 * e] = __float_as_uint(fmaf(s_x, f.x, __uint_as_float(v[8 * q + 2 * e])));
                                v[8 * q + 2 * e + 1] = __float_as_uint(fmaf(s_x, f.y, __uint_as_float(v[8 * q + 2 * e + 1])));
                            }
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {     // 4 x 16-byte chunks (8 fp16) of this half row
                        uint4 w;
                        w.x = pack_op16x2(__uint_as_float(v[8 * q + 0]) * inv, __uint_as_float(v[8 * q + 1]) * inv);
                        w.y = pack_op16x2(__uint_as_float(v[8 * q + 2]) * inv, __uint_as_float(v[8 * q + 3]) * inv);
                        w.z = pack_op16x2(__uint_as_float(v[8 * q + 4]) * inv, __uint_as_float(v[8 * q + 5]) * inv);
                        w.w = pack_op16x2(__uint_as_float(v[8 * q + 6]) * inv, __uint_as_float(v[8 * q + 7]) * inv);
                        const int chunk = hc * 4 + q;
                        sts_v4u(stg + lane * 128 + ((chunk ^ (lane & 7)) << 4), w);      // the SWIZZLE_128B image of the output box
                    }
                }
                fence_proxy_async_smem();
            }
            tcgen05_fence_before();
            __syncwarp();
            if (quad == 2 && lane == 0) ATC_TRACE(1 + g, u, 4);       // O drained
            if (lane == 0) {
                mbar_arrive(&s_free[g]);                            // TMEM buffer may be overwritten by the next QK^T
                if (p.n_extra > 0) mbar_arrive(&kv_empty[iu & 1]);  // this warp is done with the x box of the K / V stage
                // 32 rows x 64 head dims of (image, head) leave as ONE bulk store; the output map is 3-D ([image][token][D]),
                // so rows beyond the image's S tokens (padded query rows) are clipped by the TMA unit
                if (warp_valid) {
                    tma_store_3d(&tmap_o, stg, h * 64, wrow0, img);
                    atc_store_commit();
                }
            }
            __syncwarp();
        }
    }

    if (warp >= 2 && warp < 10 && lane == 0) atc_store_wait_all();   // bulk stores must have completed before the CTA exits
    __syncwarp();
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tcgen05_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}


// =====================================================================================================================
// Cooperative softmax (round 2): ALL eight softmax warps work on the SAME unit, then move to the next one.
//
// What the phase trace of the kernel above showed (profiles/r02_attention_trace.txt): with one 4-warp group per TMEM
// buffer, a group idles for ~3.5 k cycles per unit behind its own hand-off chain (P ready -> P.V -> O ready -> drain ->
// next Q K^T -> S ready, stretched by the in-order MMA thread serving the other group), its softmax passes take 4.3 k
// because a lone warp per SM sub-partition keeps the MUFU only ~55 % busy, and the two groups drift into phase instead
// of alternating: 7.8 k cycles per group and unit, 3.85 k per unit.  Here warp (quad, half) owns the 32 rows of TMEM
// lane quadrant `quad` and one HALF of the key columns, so every sub-partition runs two warps of the same pass (the
// MUFU saturates: the exp2 pass of a unit takes ~1.9 k instead of 3.1 k), and the hand-off chain of one buffer hides
// completely behind the softmax of the other:
//      wait S(u) | row max over own columns, exchange with the partner warp (shared memory + a 64-thread barrier)
//                | drain O(u-1)  (P.V(u-1) was issued when pass 2 of unit u-1 ended)   -> frees buffer (u-1) & 1, the
//                |                                                                       MMA thread issues Q K^T(u+1) into it
//                | exp2 pass over own columns, P in place                              -> MMA thread issues P.V(u)
// TMEM map of buffer b (base = b * 256), h0 / h1 = key columns of half 0 / half 1 (multiples of 16):
//      S fp32 [0, keys_pad)    P fp16: half 0 -> [0, h0/2), half 1 -> [h0, h0 + h1/2)  (each behind its own read pointer)
//      O fp32 [o_col, o_col + 64): 64 if h0 >= 128 (between the two P halves), else the next multiple of 32 behind P of half 1
// =====================================================================================================================
struct AtcSplit {
    int g0, h0, h1, o_col;
};
__host__ __device__ inline AtcSplit atc_split(int keys_pad) {
    AtcSplit c;
    const int groups = keys_pad >> 4;
    c.g0 = groups > 1 ? groups >> 1 : 1;     // odd counts: the 16-column remainder chunk (and with it the padded keys) goes to half 1
    c.h0 = c.g0 << 4;
    c.h1 = keys_pad - c.h0;
    c.o_col = c.h0 >= 128 ? 64 : ((c.h0 + (c.h1 >> 1) + 31) & ~31);
    return c;
}
constexpr int kAtcXchBytes = 2 * 2 * 128 * 16;     // [unit parity][half][row] float4 {partial max, partial sum, p_extra, -}

__host__ __device__ inline int atc_coop_smem_bytes(int keys_pad) {
    return kAtcQStages * kAtcQBytes + 2 * 2 * keys_pad * 128 + 2 * kAtcXBytes + 8 * 2048 /*O staging*/ + kAtcXchBytes + 1024 /*barriers*/ +
           1024 /*align*/;
}

__device__ __forceinline__ void bar_sync_named(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

__global__ void __launch_bounds__(kAtcThreads, 1)
attention_coop_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                      const __grid_constant__ CUtensorMap tmap_x, const AtcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int kv_bytes = p.keys_pad * 128;                       // one K or V tile
    uint8_t* s_q = smem;                                         // [kAtcQStages][16 KB]
    uint8_t* s_kv = smem + kAtcQStages * kAtcQBytes;             // [2][K | V]
    uint8_t* s_xbox = s_kv + 4 * kv_bytes;                       // [2][q | k | v of tokens 256..263]  (S > 256 only)
    uint8_t* s_stage = s_xbox + 2 * kAtcXBytes;                  // [8 warps][32 rows][64 B]
    float4* s_xch = reinterpret_cast<float4*>(s_stage + 8 * 2048);   // [2][2][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_xch) + kAtcXchBytes);
    uint64_t* q_full = bars;                       // [3]
    uint64_t* q_empty = bars + kAtcQStages;        // [3]
    uint64_t* kv_full = bars + 2 * kAtcQStages;    // [2]
    uint64_t* kv_empty = kv_full + 2;              // [2]
    uint64_t* s_full = kv_full + 4;                // [2]  MMA -> softmax warps: S ready
    uint64_t* p_full = kv_full + 6;                // [2]  softmax warps -> MMA: P written
    uint64_t* o_full = kv_full + 8;                // [2]  MMA -> softmax warps: O ready
    uint64_t* s_free = kv_full + 10;               // [2]  softmax warps -> MMA: O drained, buffer reusable
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(kv_full + 12);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_items = p.b * p.H;
    const bool pair = p.pair_mode != 0;
    const int n_work = pair ? (n_items + 1) / 2 : n_items;
    const int upi = p.units_per_item;
    const int D = p.H * 64;
    const AtcSplit sp = atc_split(p.keys_pad);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_kv);
        if (p.n_extra > 0) tma_prefetch_desc(&tmap_x);
        // with an extra key the half-1 softmax warps read their q rows from the Q tile (4 arrivals per unit) and all
        // eight read v of the extra token from the x box while draining O (8 arrivals per unit)
        for (int i = 0; i < kAtcQStages; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], p.n_extra > 0 ? 5 : 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], p.n_extra > 0 ? 2 + 8 * upi : 1);
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

    const AtcSmem sm{s_q, s_kv, s_xbox, q_full, q_empty, kv_full, kv_empty, kv_bytes};
    const int my_items = (n_work - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const uint32_t n_units = my_items > 0 ? static_cast<uint32_t>(my_items * upi) : 0u;
    if (warp == 0) {
        if (elect_one()) atc_producer(tmap_q, tmap_kv, tmap_x, p, sm);
    } else if (warp == 1) {
        if (elect_one()) {
            // ===== MMA issuer =====
            const uint32_t idesc_qk = make_idesc_f16(128, static_cast<uint32_t>(p.keys_pad));
            const uint32_t idesc_pv = make_idesc_f16(128, 64, /*a_mn_major=*/0, /*b_mn_major=*/1);
            const int ksteps = p.keys_pad >> 4;
            auto issue_qk = [&](uint32_t v) {
                const uint32_t iv = atc_unit_item(v, upi);                      // CTA-local item index of unit v
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
                ATC_TRACE(0, v, 0);                       // QK^T of unit v issued
            };
            if (n_units > 0) issue_qk(0);
            if (n_units > 1) issue_qk(1);
            for (uint32_t u = 0; u < n_units; ++u) {
                const int buf = u & 1;
                if (u >= 1 && u + 1 < n_units) {
                    // O(u-1) drained (the softmax warps do that right after the row-max pass of unit u): its buffer takes S(u+1)
                    mbar_wait(&s_free[buf ^ 1], ((u - 1) >> 1) & 1);
                    ATC_TRACE(0, u - 1, 3);
                    tcgen05_fence_after();
                    issue_qk(u + 1);
                }
                const uint32_t iu = atc_unit_item(u, upi);
                const int kvs = iu & 1;
                mbar_wait(&p_full[buf], (u >> 1) & 1);
                ATC_TRACE(0, u, 1);                       // P of unit u ready
                tcgen05_fence_after();
                // V tile: [keys][64 dh] rows of 128 B = MN-major B operand; 16 keys (one UMMA K) = 2048 B
                const uint64_t vdesc = make_smem_desc_sw128(smem_u32(s_kv + kvs * 2 * kv_bytes + kv_bytes), 1024, 1024);
                const uint32_t base = tmem_base + buf * 256;
                const uint32_t d = base + sp.o_col;
                for (int k = 0; k < ksteps; ++k) {
                    const uint32_t a = base + (k < sp.g0 ? 8 * k : sp.h0 + 8 * (k - sp.g0));
                    umma_f16_ts(d, a, vdesc + 128ull * k, idesc_pv, k != 0);
                }
                umma_commit(&o_full[buf]);
                ATC_TRACE(0, u, 2);                       // P.V of unit u issued
                if (atc_last_unit_of_item(u, upi)) umma_commit(&kv_empty[kvs]);   // last unit of the item: K / V stage reusable
            }
        }
    } else if (warp == 10) {
        if (p.n_extra > 0) atc_tail_rows(p, sm, lane);
    } else {
        // ===== softmax / epilogue warps: warp (quad, half) = rows of TMEM lane quadrant `quad`, key columns of `half` =====
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;
        const uint32_t t_lane = static_cast<uint32_t>(quad * 32) << 16;
        const uint32_t stg = smem_u32(s_stage + (warp - 2) * 2048);       // this warp's staging tile: 32 rows x 64 B
        const float c = p.scale_log2e;
        const int S_tc = p.S - p.n_extra;            // keys that go through the tensor core
        const int col_lo = half ? sp.h0 : 0;         // first S column (= key index) of this warp; its P starts at the same column
        const int ncols = half ? sp.h1 : sp.h0;
        // pair mode: rows of quads 0, 1 belong to item 0 of the pair (keys 0..63), quads 2, 3 to item 1 (keys 64..127);
        // h0 = h1 = 64, so a warp either owns its item's keys (half == own item) or the other item's (probability 0)
        const int pitem = pair ? (quad >> 1) : 0;
        const bool own_keys = !pair || half == pitem;
        const int key0 = pair ? 0 : col_lo;          // key index of this warp's first column within its item
        const int nfull = ncols >> 5;
        const bool rem16 = (ncols & 16) != 0;
        const int row_t = quad * 32 + lane;          // row of the 128-row unit tile
        // explicit shared-space accesses (a pointer carved out of the dynamic buffer compiles to generic LD / ST)
        const uint32_t xch_mine = smem_u32(s_xch) + (half * 128 + row_t) * 16;
        const uint32_t xch_other = smem_u32(s_xch) + ((half ^ 1) * 128 + row_t) * 16;

        // drain O of unit v: this warp's 32 head dims (half) of its 32 rows -> fp16 -> staging -> coalesced 64-byte row pieces
        auto drain = [&](uint32_t v) {
            const int vbuf = v & 1, vpar = (v >> 1) & 1;
            const uint32_t iv = atc_unit_item(v, upi);
            const int mt = static_cast<int>(v - iv * upi);
            const int work = static_cast<int>(blockIdx.x) + static_cast<int>(iv) * static_cast<int>(gridDim.x);
            const int item = pair ? 2 * work + pitem : work;
            const int img = atc_div_h(item, p.inv_H), h = item - img * p.H;
            const int wrow0 = pair ? (quad & 1) * 32 : mt * 128 + quad * 32;
            const bool valid = item < n_items && wrow0 < p.S;
            mbar_wait(&o_full[vbuf], vpar);
            if (quad == 2 && lane == 0) ATC_TRACE(1 + half, v, 3);       // O ready
            tcgen05_fence_after();
            if (valid) {
                const uint32_t xo = static_cast<uint32_t>(vbuf) * 2 * 128 * 16;          // exchange slots of unit parity v & 1
                const float4 a = lds_v4(xch_mine + xo), b = lds_v4(xch_other + xo);
                const float inv = 1.0f / (a.y + b.y);
                uint32_t o[32];
                tmem_ld_32x32b_x32(tmem_base + t_lane + vbuf * 256 + sp.o_col + half * 32, o);
                tmem_ld_wait();
                if (p.n_extra > 0) {   // O += p_extra * v_extra (warp-uniform addresses: broadcast loads)
                    const float px = half ? a.z : b.z;
                    const uint32_t sx = smem_u32(s_xbox + (iv & 1) * kAtcXBytes);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint4 w = lds_v4u(sx + 2048 + ((half * 4 + q) << 4));
                        const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 f = unpack_op16x2(ws[e]);
                            o[8 * q + 2 * e] = __float_as_uint(fmaf(px, f.x, __uint_as_float(o[8 * q + 2 * e])));
                            o[8 * q + 2 * e + 1] = __float_as_uint(fmaf(px, f.y, __uint_as_float(o[8 * q + 2 * e + 1])));
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {     // 4 x 16-byte chunks (8 fp16) of this row's 32 head dims
                    uint4 w;
                    w.x = pack_op16x2(__uint_as_float(o[8 * q + 0]) * inv, __uint_as_float(o[8 * q + 1]) * inv);
                    w.y = pack_op16x2(__uint_as_float(o[8 * q + 2]) * inv, __uint_as_float(o[8 * q + 3]) * inv);
                    w.z = pack_op16x2(__uint_as_float(o[8 * q + 4]) * inv, __uint_as_float(o[8 * q + 5]) * inv);
                    w.w = pack_op16x2(__uint_as_float(o[8 * q + 6]) * inv, __uint_as_float(o[8 * q + 7]) * inv);
                    sts_v4u(stg + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4), w);
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (quad == 2 && lane == 0) ATC_TRACE(1 + half, v, 4);       // O drained
            if (lane == 0) {
                mbar_arrive(&s_free[vbuf]);                           // the buffer may take the next S
                if (p.n_extra > 0) mbar_arrive(&kv_empty[iv & 1]);    // this warp is done with the x box of the K / V stage
            }
            if (valid) {
                const int cq = lane & 3;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = (lane >> 2) + 8 * i;
                    const int row = wrow0 + r;
                    if (row < p.S) {
                        const uint4 w = lds_v4u(stg + r * 64 + ((cq ^ ((r >> 1) & 3)) << 4));
                        *reinterpret_cast<uint4*>(p.out + (static_cast<size_t>(img) * p.S + row) * D + h * 64 + half * 32 + cq * 8) = w;
                    }
                }
            }
            __syncwarp();
        };

        for (uint32_t u = 0; u < n_units; ++u) {
            const int buf = u & 1;
            const uint32_t j = u >> 1;
            const uint32_t iu = atc_unit_item(u, upi);
            const int mt = static_cast<int>(u - iu * upi);
            const int work = static_cast<int>(blockIdx.x) + static_cast<int>(iu) * static_cast<int>(gridDim.x);
            const int item = pair ? 2 * work + pitem : work;
            const int wrow0 = pair ? (quad & 1) * 32 : mt * 128 + quad * 32;
            const bool warp_valid = item < n_items && wrow0 < p.S;
            const uint32_t t_s = tmem_base + t_lane + buf * 256 + col_lo;      // first S column of this warp (its P starts here too)
            const uint32_t xo = static_cast<uint32_t>(buf) * 2 * 128 * 16;
            float s_x = -INFINITY;
            const uint32_t sx = smem_u32(s_xbox + (iu & 1) * kAtcXBytes);

            mbar_wait(&s_full[buf], j & 1);
            if (p.n_extra > 0 && half == 1) {
                // the extra key (ViT-L/14's 257th token): this row's raw score against it, on the CUDA cores (half-1 warps).
                // S ready => the MMA thread has seen this unit's Q tile and the item's K / V stage land; observing the same
                // (completed) phases here makes the TMA writes visible to this warp
                mbar_wait(&q_full[u % kAtcQStages], (u / kAtcQStages) & 1);
                mbar_wait(&kv_full[iu & 1], (iu >> 1) & 1);
                if (warp_valid)
                    s_x = atc_dot64(smem_u32(s_q + (u % kAtcQStages) * kAtcQBytes) + (quad * 32 + lane) * 128, lane & 7, sx + 1024);
                __syncwarp();
                if (lane == 0) mbar_arrive(&q_empty[u % kAtcQStages]);
            }
            if (quad == 2 && lane == 0) ATC_TRACE(1 + half, u, 0);       // S ready
            tcgen05_fence_after();

            // ---- pass 1: partial row max over this warp's key columns ----
            float mx = s_x;
            if (warp_valid && own_keys) {
                for (int ch = 0; ch < nfull; ch += 2) {
                    uint32_t va[32], vb[32];
                    const bool two = ch + 1 < nfull;
                    tmem_ld_32x32b_x32(t_s + ch * 32, va);
                    if (two) tmem_ld_32x32b_x32(t_s + ch * 32 + 32, vb);
                    tmem_ld_wait();
                    mx = atc_chunk_max(va, key0 + ch * 32, S_tc, mx);
                    if (two) mx = atc_chunk_max(vb, key0 + ch * 32 + 32, S_tc, mx);
                }
                if (rem16) {
                    uint32_t v[16];
                    tmem_ld_32x32b_x16(t_s + nfull * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        if (key0 + nfull * 32 + e < S_tc) mx = fmaxf(mx, __uint_as_float(v[e]));
                }
            }
            sts_v4(xch_mine + xo, make_float4(mx, 0.f, 0.f, 0.f));
            bar_sync_named(1 + quad, 64);            // the two warps that share these 32 rows
            const float mc = fmaxf(mx, lds_v4(xch_other + xo).x) * c;
            if (quad == 2 && lane == 0) ATC_TRACE(1 + half, u, 1);   // pass 1 done

            // ---- O of the previous unit: its P.V was issued when that unit's pass 2 ended ----
            if (u > 0) drain(u - 1);

            // ---- pass 2: p = 2^(s * c - max * c); P (fp16) goes over the first half of this warp's own S columns ----
            float sum0 = 0.f, sum1 = 0.f, px = 0.f;
            if (warp_valid) {
                if (own_keys) {
                    for (int ch = 0; ch < nfull; ++ch) {
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(t_s + ch * 32, v);
                        tmem_ld_wait();
                        atc_chunk_exp(v, key0 + ch * 32, S_tc, c, mc, sum0, sum1, t_s + ch * 16);
                    }
                    if (rem16) {
                        uint32_t v[16];
                        tmem_ld_32x32b_x16(t_s + nfull * 32, v);
                        tmem_ld_wait();
                        uint32_t pk[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int k0 = key0 + nfull * 32 + 2 * e;
                            const float p0 = (k0 < S_tc) ? ex2_approx(fmaf(__uint_as_float(v[2 * e]), c, -mc)) : 0.f;
                            const float p1 = (k0 + 1 < S_tc) ? ex2_approx(fmaf(__uint_as_float(v[2 * e + 1]), c, -mc)) : 0.f;
                            sum0 += p0;
                            sum1 += p1;
                            pk[e] = pack_op16x2(p0, p1);
                        }
                        tmem_st_32x32b_x8(t_s + nfull * 16, pk);
                    }
                    if (p.n_extra > 0 && half == 1) {   // probability of the extra key, rounded like the P operand
                        const float pe = ex2_approx(fmaf(s_x, c, -mc));
                        sum0 += pe;
                        px = unpack_op16x2(pack_op16x2(pe, 0.f)).x;
                    }
                } else {      // pair mode: this warp's columns are the OTHER item's keys: probability 0
                    uint32_t zero[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) zero[e] = 0u;
                    for (int i = 0; i < (ncols >> 5); ++i) tmem_st_32x32b_x16(t_s + 16 * i, zero);
                }
                tmem_st_wait();
            }
            sts_v4(xch_mine + xo, make_float4(mx, sum0 + sum1, px, 0.f));
            tcgen05_fence_before();
            __syncwarp();
            if (quad == 2 && lane == 0) ATC_TRACE(1 + half, u, 2);       // pass 2 done
            if (lane == 0) mbar_arrive(&p_full[buf]);
        }
        if (n_units > 0) {
            bar_sync_named(1 + quad, 64);        // the partner's sums of the last unit (no later row-max exchange orders them)
            drain(n_units - 1);
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
