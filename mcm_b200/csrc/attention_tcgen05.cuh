// Multi-head self-attention core on the 5th-gen tensor cores:  softmax(Q K^T / 8) V  per (image, head)
// for sequences of up to 257 tokens (ViT-B/16: 197, ViT-B/32: 50, ViT-L/14: 257), head dim 64, no mask, not
// causal (HF:modeling_clip.py:261-279 eager == what SDPA computes, :318-331).
//
// 257 = 256 + 1: a UMMA tile is at most 256 keys wide and two fp32 score buffers of more than 256 columns do
// not fit the 512 columns of tensor memory, so the ONE key beyond 256 of ViT-L/14 ("extra key") never enters
// the tensor core: every softmax thread computes its row's score against it with 64 FMAs (its q row from the Q
// tile, the key from a small "x box": the producer also loads rows 256.. of q / k / v, 8 rows x 128 B each, with
// every item), folds it into the row max / row sum, and adds p_extra * v_extra to the O row while draining it.
// (Reading q / k / v of that token from global memory instead put ~5 k cycles of load latency on every unit.)
// The 256 other keys take the normal path with keys_pad = 256.  Likewise the one QUERY row beyond 256 does not
// get a (1 / 128 full) third unit: warp 10 computes it with warp-level mma.sync tiles from the K / V tiles already
// in shared memory, hidden behind the two real units.
//
// Persistent CTAs (one per SM) walk over (image, head) items; every item is cut into 128-query-row
// units; unit u of a CTA lives in TMEM score buffer u & 1 and is served by softmax group u & 1.  Warp roles (12 warps):
//   warp 0      TMA producer: K and V tiles of the item ([keys_pad x 64] fp16 boxes, 128-byte swizzle, 2 stages, K and V
//               with their own full / empty barrier pairs; 3-D tensor map over [image][token][3 D], so key rows beyond the
//               image's S tokens arrive as zeros) and the Q tile of every unit (3-stage ring), straight out of the fused
//               QKV buffer.  Issue order per item: K, Q of the first unit, V (+ x box), the other Q tiles.
//   warps 1, 11 MMA issuers, one elected thread per TMEM buffer (warp 1: buffer 0, warp 11: buffer 1):
//                 S = Q K^T  as UMMA 128 x keys_pad x 16 (x4, SS mode),
//                 O = P V    as UMMA 128 x 64 x 16 (x keys_pad/16, TS mode: P is read from TENSOR MEMORY, V from shared
//               memory as an MN-major operand), issued in up to three parts of 64 keys while the exp2 pass is still
//               producing the later columns (one mbarrier per part, see atc_two_pass).  Each thread blocks only on its own
//               buffer's barriers; the tensor core executes the two streams in arrival order.
//   warps 2-9   two softmax groups of 4 warps, one thread per query row: row max and exp2 in fp32 straight from TMEM
//               (tcgen05.ld), P written back over the dead S columns as fp16 (tcgen05.st), then O / rowsum -> fp16 -> the
//               SWIZZLE_128B image of a [32 rows x 64] box in shared memory -> ONE 3-D TMA bulk store per warp and unit (rows
//               beyond the image's S tokens are clipped by the TMA unit).  The loops over the 32-key chunks have compile-time
//               trip counts for the two production shapes (template NFULL / REM16).
//   warp 10     tail rows (query rows >= 256, ViT-L/14 only) on mma.sync tiles.
// TMEM map.  keys_pad <= 224 (ViT-B/16: 208): S / P of buffer g at columns g * keys_pad, ONE O tile shared by the two
// groups at 2 * keys_pad (64 columns): P.V writes outside the score buffer, so the next-but-one Q K^T is issued right behind
// P.V and the group drains O while the tensor core already computes its next S.  Wider rows (ViT-L/14: 256) leave no room: O
// stays inside each 256-column buffer (columns 128..191, dead by the time P.V runs) and P.V is issued in one piece.
// Padded keys (>= S) get probability 0; padded query rows are computed and never stored; a warp whose 32 rows are all padding
// skips both passes.  Sequences of at most 64 tokens (ViT-B/32: 50) put TWO items into a unit (pair_mode).
//
// Measured (round 2, clock64 phase trace of one CTA with the -DMCM_ATC_TRACE build, profiles/r02_attention_trace.txt;
// ViT-B/16, b = 512): a group turns a unit around in 5.7 k cycles -- both passes 3.3-3.6 k, wait for the last part of P.V
// 0.8-0.9 k, O drain + store 0.6 k, wait for the next S 0.75 k -- and the two groups interleave, so the SM finishes a unit
// every 2.85 k cycles (round 1: 3.9 k).  Eight softmax warps running NOTHING but the two passes need 2.53 k
// (tools/microbench/softmax_unit.cu); the MUFU bound is 1.66 k.  147 us per launch against 286 us in round 1 and 293 us for
// the mma.sync kernel.  What got it there, in order of effect: constant-trip loops + bulk stores of O (173 -> 154 us),
// one issuer thread per buffer + P.V in parts + separate K / V barriers (154 -> 147), index math without integer divisions
// (179 -> 174), the shared O tile (174 -> 173).  Tried and rejected this round: a cooperative variant with both groups on
// one unit's two key halves (20-30 % slower: the max / sum exchange through shared memory serialises the groups), one
// polling issuer thread using mbarrier.test_wait (~150 cycles per probe; 177 us; with one thread per buffer polling its
// one barrier: 145.9 vs 145.4 us, no difference), P.V in four parts (150 us), part of the
// exponentials on the FMA pipe (ex2_poly below: flat to 12 %, slower beyond), and -- once more, after round 1 -- two softmax
// warps per (group, quadrant) splitting the key chunks of their 32 rows (20 warps, 96 registers, row max / sum exchanged
// through shared memory, each half packing its P at the start of its own score columns): 167 vs 155 us on the same box,
// ViT-L/14 103 vs 85 us.  Four warps per sub-partition do not buy what two cannot hide; the exchange barriers cost more.  Round 1's rejected list (software-pipelined
// TMEM loads, lazy row maximum, two threads per row, score chunks kept in registers between the passes, the 69 rows beyond
// 128 on mma.sync warps) is in DESIGN.md.
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

constexpr int kAtcThreads = 384;          // 12 warps: TMA, MMA issuer of buffer 0, 2 x 4 softmax, tail-row warp (idle unless S > 256), MMA issuer of buffer 1
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
    int reverse;             // walk the items from the last image to the first (L2 reuse: start on the rows the QKV GEMM wrote last)
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
// 2^x for x <= 0 WITHOUT the special-function unit: round-to-nearest split x = j + f (the magic-number add leaves j in the
// low mantissa bits of r), a degree-3 minimax polynomial for 2^f on [-0.5, 0.5] (relative error 7.5e-5, a sixth of the
// fp16 rounding the probability gets next), and j added into the exponent field: 8 FMA/ALU-pipe instructions against ONE
// MUFU.EX2 (16 lanes per clock and SM against 128 FMA lanes).  MEASURED, round 2 (tools/attn_sweep.py, b = 512, S = 197):
// 0 / 12 / 25 / 31 / 37 / 50 % of the keys through the polynomial = 145.3 / 144.3 / 146.2 / 149.5 / 153.6 / 158.3 us -- ptxas
// does interleave the two instruction streams, but the exp2 pass is not paced by the MUFU pipe (math-pipe-throttle stalls:
// 1 % of the samples; the two groups' passes overlap in time and the SM's MUFU lanes are 58 % busy), so the extra issue
// slots buy nothing.  Off by default (mask 0); kept as an A/B build switch.
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -125.f);                                   // keeps the exponent field positive; 2^-125 rounds to P = 0 anyway
    const float r = x + 12582912.f;                         // 1.5 * 2^23
    const float f = x - (r - 12582912.f);
    float p = fmaf(0.0551716685295105f, f, 0.2426111251115799f);
    p = fmaf(p, f, 0.6932609677314758f);
    p = fmaf(p, f, 0.9999280571937561f);
    return __uint_as_float(__float_as_uint(p) + (__float_as_uint(r) << 23));
}
// which of the 32 keys of a full chunk take ex2_poly instead of MUFU.EX2 (bit i = key i of the chunk)
#ifndef MCM_ATC_POLY_MASK
#define MCM_ATC_POLY_MASK 0u
#endif
template <int I>
__device__ __forceinline__ float atc_ex2(float x) {
    if constexpr (((MCM_ATC_POLY_MASK >> I) & 1u) != 0) return ex2_poly(x);
    else return ex2_approx(x);
}
template <int E>
__device__ __forceinline__ void atc_exp_pairs(const uint32_t (&v)[32], float c, float mc, float& sum0, float& sum1, uint32_t (&pk)[16]) {
    if constexpr (E < 16) {
        const float p0 = atc_ex2<2 * E>(fmaf(__uint_as_float(v[2 * E]), c, -mc));
        const float p1 = atc_ex2<2 * E + 1>(fmaf(__uint_as_float(v[2 * E + 1]), c, -mc));
        sum0 += p0;
        sum1 += p1;
        pk[E] = pack_op16x2(p0, p1);
        atc_exp_pairs<E + 1>(v, c, mc, sum0, sum1, pk);
    }
}

// p = 2^(s * c - mc) for one 32-key chunk, accumulated into two partial row sums and written to TMEM
// as 16 packed fp16 pairs (the A operand of P.V) at `t_p`
__device__ __forceinline__ void atc_chunk_exp(const uint32_t (&v)[32], int k0, int S, float c, float mc, float& sum0,
                                              float& sum1, uint32_t t_p) {
    uint32_t pk[16];
    if (k0 + 32 <= S) {
        atc_exp_pairs<0>(v, c, mc, sum0, sum1, pk);
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
// P.V is issued in up to three parts of 64 keys (4 UMMA k-steps) so that most of it runs WHILE pass 2 is still producing
// the later columns: after chunk 2 p + 1 the warp makes its P columns visible (tcgen05.wait::st + fence) and arrives on
// its group's barrier OF THAT PART (p_part[p], p < nparts - 1; the caller signals the last part).  One barrier per part, one
// phase per unit: arrivals carry no phase tag, so a fast warp signalling part p + 1 on a single multi-phase barrier before its
// neighbours signalled part p would be counted towards part p (that deadlocked the first version).  atc_parts() is the split
// both sides agree on.
#ifndef MCM_ATC_MAX_PARTS
#define MCM_ATC_MAX_PARTS 3
#endif
constexpr int kAtcMaxParts = MCM_ATC_MAX_PARTS;
__host__ __device__ inline int atc_parts(int keys_pad, int pair_mode) {
    const int n4 = ((keys_pad >> 4) + 3) >> 2;    // groups of 4 k-steps, the last one possibly short
    return pair_mode ? 1 : (n4 > kAtcMaxParts ? kAtcMaxParts : n4);
}
__device__ __forceinline__ void atc_signal_part(uint64_t* p_part, int lane) {
    tmem_st_wait();
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(p_part);
}

template <int NFULL, bool REM16>
__device__ __forceinline__ float atc_two_pass(uint32_t t_sr, uint32_t t_pw, int nfull_rt, bool rem16_rt, int S_tc, float c, float& s_x,
                                              uint64_t* p_part, int nparts, int lane) {
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
        if (ch == 1 && nparts > 1) atc_signal_part(p_part, lane);          // keys 0..63 are final
        if (ch == 3 && nparts > 2) atc_signal_part(p_part + 1, lane);      // keys 64..127 are final
        if (kAtcMaxParts > 3 && ch == 5 && nparts > 3) atc_signal_part(p_part + 2, lane);   // keys 128..191
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
    uint64_t *q_full, *q_empty, *k_full, *k_empty, *v_full, *v_empty;
    int kv_bytes;
};
// K and V have their own full / empty barrier pairs: the K tile of a stage is free as soon as the last Q K^T of its item
// has run, a whole unit before the last P.V releases the V tile, so the next-but-one item's K is on its way that much earlier
// (with one barrier pair per stage the next Q K^T waited ~2 k cycles for its K tile once P.V stopped being the bottleneck).
__device__ __forceinline__ void atc_producer(const CUtensorMap& tmap_q, const CUtensorMap& tmap_kv, const CUtensorMap& tmap_x,
                                             const AtcParams& p, const AtcSmem& sm) {
    uint8_t *const s_q = sm.s_q, *const s_kv = sm.s_kv, *const s_xbox = sm.s_xbox;
    uint64_t *const q_full = sm.q_full, *const q_empty = sm.q_empty;
    uint64_t *const k_full = sm.k_full, *const k_empty = sm.k_empty, *const v_full = sm.v_full, *const v_empty = sm.v_empty;
    const int kv_bytes = sm.kv_bytes;
    const int n_items = p.b * p.H;
    const bool pair = p.pair_mode != 0;
    const int n_work = pair ? (n_items + 1) / 2 : n_items;
    const int upi = p.units_per_item;
    const int D = p.H * 64;
    uint32_t ic = 0, uc = 0;
    auto load_q = [&](int h, int row0, int mt) {
        const int qs = uc % kAtcQStages;
        mbar_wait(&q_empty[qs], ((uc / kAtcQStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[qs], kAtcQBytes);
        tma_load_2d(s_q + qs * kAtcQBytes, &tmap_q, &q_full[qs], h * 64, row0 + mt * 128);
        ++uc;
    };
    for (int slot = blockIdx.x; slot < n_work; slot += gridDim.x, ++ic) {
        const int item = p.reverse ? n_work - 1 - slot : slot;
        const int kvs = ic & 1;
        const uint32_t ph = ((ic >> 1) & 1) ^ 1;
        uint8_t* sk = s_kv + kvs * 2 * kv_bytes;
        if (pair) {
            // pair mode: 64-row boxes (tmap_kv) of item 2w and item 2w + 1 (the last item again if n_items is odd)
            // stacked into 128-row Q / K / V tiles; 8 KB per box keeps the 128-byte swizzle phase of the rows
            int img2[2], h2[2];
            for (int half = 0; half < 2; ++half) {
                const int it2 = min(2 * item + half, n_items - 1);
                img2[half] = atc_div_h(it2, p.inv_H);
                h2[half] = it2 - img2[half] * p.H;
            }
            mbar_wait(&k_empty[kvs], ph);
            mbar_arrive_expect_tx(&k_full[kvs], kv_bytes);
            for (int half = 0; half < 2; ++half) tma_load_3d(sk + half * 8192, &tmap_kv, &k_full[kvs], D + h2[half] * 64, 0, img2[half]);
            const int qs = uc % kAtcQStages;
            mbar_wait(&q_empty[qs], ((uc / kAtcQStages) & 1) ^ 1);
            mbar_arrive_expect_tx(&q_full[qs], kAtcQBytes);
            for (int half = 0; half < 2; ++half)
                tma_load_3d(s_q + qs * kAtcQBytes + half * 8192, &tmap_kv, &q_full[qs], h2[half] * 64, 0, img2[half]);
            ++uc;
            mbar_wait(&v_empty[kvs], ph);
            mbar_arrive_expect_tx(&v_full[kvs], kv_bytes);
            for (int half = 0; half < 2; ++half)
                tma_load_3d(sk + kv_bytes + half * 8192, &tmap_kv, &v_full[kvs], 2 * D + h2[half] * 64, 0, img2[half]);
            continue;
        }
        const int img = atc_div_h(item, p.inv_H), h = item - img * p.H;
        const int row0 = img * p.S;
        // tmap_kv is 3-D ([image][token][3 D]): the padded key rows S .. keys_pad - 1 lie outside the image's plane and
        // arrive as ZEROS -- never another image's (or a stale batch's) rows, whose Inf / NaN would leak through 0 * V
        mbar_wait(&k_empty[kvs], ph);
        mbar_arrive_expect_tx(&k_full[kvs], kv_bytes);
        tma_load_3d(sk, &tmap_kv, &k_full[kvs], D + h * 64, 0, img);
        load_q(h, row0, 0);
        mbar_wait(&v_empty[kvs], ph);
        mbar_arrive_expect_tx(&v_full[kvs], kv_bytes + (p.n_extra > 0 ? kAtcXBytes : 0));
        tma_load_3d(sk + kv_bytes, &tmap_kv, &v_full[kvs], 2 * D + h * 64, 0, img);
        if (p.n_extra > 0) {     // q | k | v of tokens 256..: travels (and is released) with the V tile
            uint8_t* sx = s_xbox + kvs * kAtcXBytes;
#pragma unroll
            for (int part = 0; part < 3; ++part)
                tma_load_2d(sx + part * 1024, &tmap_x, &v_full[kvs], part * D + h * 64, row0 + 256);
        }
        for (int mt = 1; mt < upi; ++mt) load_q(h, row0, mt);
    }
}

// ===== tail-row warp: query rows >= 256 (ViT-L/14: the one row 256) against all S keys =====
__device__ __forceinline__ void atc_tail_rows(const AtcParams& p, const AtcSmem& sm, int lane) {
    uint8_t *const s_kv = sm.s_kv, *const s_xbox = sm.s_xbox;
    uint64_t *const k_full = sm.k_full, *const k_empty = sm.k_empty, *const v_full = sm.v_full, *const v_empty = sm.v_empty;
    const int kv_bytes = sm.kv_bytes;
    const int n_items = p.b * p.H;
    const int D = p.H * 64;
    const float c = p.scale_log2e;
    const int tq = lane & 3;
    const bool row_lane = (lane >> 2) == 0;      // lanes 0..3 hold row 0 of the tile
    uint32_t ic = 0;
    auto lds32 = [](uint32_t a) { uint32_t w; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(a)); return w; };
    for (int slot = blockIdx.x; slot < n_items; slot += gridDim.x, ++ic) {
        const int item = p.reverse ? n_items - 1 - slot : slot;
        const int img = atc_div_h(item, p.inv_H), h = item - img * p.H;
        const int kvs = ic & 1;
        const uint32_t sk = smem_u32(s_kv + kvs * 2 * kv_bytes);
        const uint32_t sv = sk + kv_bytes;
        const uint32_t sx = smem_u32(s_xbox + kvs * kAtcXBytes);      // q | k | v of tokens 256.., row e swizzled by e
        mbar_wait(&k_full[kvs], (ic >> 1) & 1);
        mbar_wait(&v_full[kvs], (ic >> 1) & 1);
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
        if (lane == 0) {     // this warp is done with the K / V stage
            mbar_arrive(&k_empty[kvs]);
            mbar_arrive(&v_empty[kvs]);
        }
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
    uint64_t* k_full = bars + 2 * kAtcQStages;     // [2]  K tile of an item landed / its last Q K^T has run
    uint64_t* k_empty = k_full + 2;                // [2]
    uint64_t* v_full = k_full + 4;                 // [2]  V tile (+ x box) landed / its last P.V has run (and its readers are done)
    uint64_t* v_empty = k_full + 6;                // [2]
    uint64_t* s_full = k_full + 8;                 // [2]  MMA -> softmax group: S ready
    uint64_t* o_full = k_full + 10;                // [2]  MMA -> softmax group: O ready
    uint64_t* s_free = k_full + 12;                // [2]  softmax group -> MMA: O drained
    uint64_t* p_part = k_full + 14;                // [2][kAtcMaxParts]  softmax group -> MMA: part p (64 keys) of P written
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(k_full + 14 + 2 * kAtcMaxParts);

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
    // P.V is issued in nparts pieces while pass 2 still runs -- possible only when O lies outside the score buffer (inside
    // it, columns 128..191 still hold unread scores while the first keys of P are ready)
    const int nparts = shared_o ? atc_parts(p.keys_pad, p.pair_mode) : 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_kv);
        if (p.n_extra > 0) tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_o);
        // with an extra key the softmax warps read their q rows from the Q tile and k / v of the extra token from the
        // x box (which travels with the V tile): they release the Q stage and the V stage together with the MMA thread's
        // commits; the tail-row warp reads K and V
        for (int i = 0; i < kAtcQStages; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], p.n_extra > 0 ? 5 : 1); }
        for (int i = 0; i < 2; ++i) {
            // every unit of an item commits once on its K tile (after its Q K^T) and once on its V tile (after its P.V): the
            // units of one item are issued by different threads, and a commit only tracks the issuing thread's MMAs
            mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], upi + (p.n_extra > 0 ? 1 : 0));            // + tail-row warp
            mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], upi + (p.n_extra > 0 ? 1 + 4 * upi : 0));  // + tail-row warp + 4 softmax warps per unit
            mbar_init(&s_full[i], 1);
            for (int q = 0; q < kAtcMaxParts; ++q) mbar_init(&p_part[kAtcMaxParts * i + q], 4);
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

    const AtcSmem sm{s_q, s_kv, s_xbox, q_full, q_empty, k_full, k_empty, v_full, v_empty, kv_bytes};
    if (warp == 0) {
        if (elect_one()) atc_producer(tmap_q, tmap_kv, tmap_x, p, sm);
    } else if (warp == 1 || warp == 11) {
        if (elect_one()) {
            // ===== MMA issuers =====
            // Unit u lives in TMEM buffer u & 1:  Q K^T -> [softmax] -> P.V in nparts pieces -> [drain].
            // Shared O tile (keys_pad <= 224): one thread per buffer (warp 1: units 0, 2, ..; warp 11: units 1, 3, ..).  Each
            // blocks only on ITS buffer's barriers, so a group never waits behind the other group's hand-offs (round 1's
            // single in-order thread: ~1.9 k idle cycles per unit, profiles/r02_attention_trace.txt), and the hardware-assisted
            // try_wait wakes it ~60 cycles after the arrive (polling both buffers with test_wait from one thread cost ~150
            // cycles per probe and was slower than the in-order loop).  The tensor core executes the two streams in the order
            // it receives them; every dependency between them goes through a barrier: the shared O tile (s_free of the other
            // group), the K / V stages (k_empty / v_empty count one commit per unit of the item).
            // O inside the score buffers (ViT-L/14): the next Q K^T of a buffer has to wait for the drain anyway, there is
            // nothing for a second thread to overlap, and ONE in-order thread (warp 1) measured 6 % faster (80.7 vs 86 us).
            const uint32_t first = (shared_o && warp == 11) ? 1u : 0u;
            const uint32_t stride = shared_o ? 2u : 1u;
            const int my_items = (n_work - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
            const uint32_t n_units = (!shared_o && warp == 11) ? 0u : static_cast<uint32_t>(my_items * upi);
            const uint32_t idesc_qk = make_idesc_f16(128, static_cast<uint32_t>(p.keys_pad));
            const uint32_t idesc_pv = make_idesc_f16(128, 64, /*a_mn_major=*/0, /*b_mn_major=*/1);
            const int ksteps = p.keys_pad >> 4;
            auto issue_qk = [&](uint32_t v) {
                const uint32_t iv = atc_unit_item(v, upi);                      // CTA-local item index of unit v
                const int kvs = iv & 1, qs = v % kAtcQStages, buf = v & 1;
                mbar_wait(&k_full[kvs], (iv >> 1) & 1);
                mbar_wait(&q_full[qs], (v / kAtcQStages) & 1);
                tcgen05_fence_after();
                const uint64_t adesc = make_smem_desc_sw128(smem_u32(s_q + qs * kAtcQBytes), 16, 1024);
                const uint64_t bdesc = make_smem_desc_sw128(smem_u32(s_kv + kvs * 2 * kv_bytes), 16, 1024);
                const uint32_t a_tmem = tmem_base + buf * buf_stride;
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16(a_tmem, adesc + 2 * k, bdesc + 2 * k, idesc_qk, k != 0);
                umma_commit(&q_empty[qs]);
                umma_commit(&s_full[buf]);
                umma_commit(&k_empty[kvs]);               // this unit is done with the item's K tile
                ATC_TRACE(0, v, 0);                       // QK^T of unit v issued
            };
            for (uint32_t v = first; v < 2u && v < n_units; v += stride) issue_qk(v);
            for (uint32_t u = first; u < n_units; u += stride) {
                const uint32_t iu = atc_unit_item(u, upi);
                const int kvs = iu & 1, buf = u & 1;
                const uint32_t a_tmem = tmem_base + buf * buf_stride;                     // S, then P, of this unit's buffer
                const uint32_t o_tmem = shared_o ? tmem_base + 2 * buf_stride : a_tmem + 128;
                // V tile: [keys][64 dh] rows of 128 B = MN-major B operand; 16 keys (one UMMA K) = 2048 B
                const uint64_t vdesc = make_smem_desc_sw128(smem_u32(s_kv + kvs * 2 * kv_bytes + kv_bytes), 1024, 1024);
                for (int part = 0; part < nparts; ++part) {
                    if (part == 0) {
                        if (shared_o && u >= 1) mbar_wait(&s_free[buf ^ 1], ((u - 1) >> 1) & 1);   // the O tile: drained by unit u - 1's group
                        mbar_wait(&v_full[kvs], (iu >> 1) & 1);
                    }
                    mbar_wait(&p_part[kAtcMaxParts * buf + part], (u >> 1) & 1);
                    if (part == 0) ATC_TRACE(0, u, 1);            // first keys of P of unit u ready
                    if (part + 1 == nparts) ATC_TRACE(0, u, 4);   // last part of P of unit u ready
                    tcgen05_fence_after();
                    const int k0 = 4 * part, k1 = part + 1 == nparts ? ksteps : 4 * (part + 1);
                    for (int k = k0; k < k1; ++k) umma_f16_ts(o_tmem, a_tmem + 8 * k, vdesc + 128ull * k, idesc_pv, k != 0);
                }
                ATC_TRACE(0, u, 5);                       // last P.V instruction of unit u issued
                umma_commit(&o_full[buf]);
                umma_commit(&v_empty[kvs]);               // this unit is done with the item's V tile
                ATC_TRACE(0, u, 2);                       // P.V of unit u issued
                if (u + 2 < n_units) {
                    if (!shared_o) {      // O lives inside the score buffer: wait for the drain
                        mbar_wait(&s_free[buf], (u >> 1) & 1);
                        ATC_TRACE(0, u, 3);
                        tcgen05_fence_after();
                    }
                    issue_qk(u + 2);      // the tensor core runs it behind P.V(u), the last reader of P(u)
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
        if (g == 1 && n_units > 1) mbar_wait(&p_part[nparts - 1], 0);
#endif
        for (uint32_t u = g; u < n_units; u += 2) {
            const uint32_t j = u >> 1;
            const uint32_t iu = atc_unit_item(u, upi);
            const int mt = static_cast<int>(u - iu * upi);
            const int slot = static_cast<int>(blockIdx.x) + static_cast<int>(iu) * static_cast<int>(gridDim.x);
            const int work = p.reverse ? n_work - 1 - slot : slot;
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
                mbar_wait(&v_full[iu & 1], (iu >> 1) & 1);
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
                row_sum = atc_two_pass<NFULL, REM16>(t_sr, t_pw, nfull, rem16, S_tc, c, s_x, &p_part[kAtcMaxParts * g], nparts, lane);
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
            if (quad != 2 && lane == 0) ATC_TRACE(1 + g, u, 5 + (quad == 3 ? 2 : quad));   // ... on the other three quadrants (events 5, 6, 7)
            if (lane == 0) {
                if (!warp_valid)      // a warp without rows still owes the barriers of the earlier parts its arrival
                    for (int i = 0; i + 1 < nparts; ++i) mbar_arrive(&p_part[kAtcMaxParts * g + i]);
                mbar_arrive(&p_part[kAtcMaxParts * g + nparts - 1]);
            }

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
                if (p.n_extra > 0) mbar_arrive(&v_empty[iu & 1]);   // this warp is done with the x box of the V stage
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


}  // namespace mcm
