// HBM-bound row-wise kernels of the vision tower: patch gather, LayerNorm, embedding finish.
// One warp owns one token row; rows live in registers as float4 (lane-strided, fully coalesced
// 512-byte warp transactions); all statistics are fp32 with warp-shuffle reductions.
//
// Reference arithmetic: nn.LayerNorm (biased variance, eps inside the sqrt) as used by HF CLIP,
//   pre_layrnorm HF:modeling_clip.py:659,677; layer_norm1/2 HF:359-361,371,380;
//   embeddings (CLS concat + position add) HF:212-217.
#pragma once
#include "ptx.cuh"

namespace mcm {

constexpr int kRowThreads = 256;  // 8 warps = 8 rows per CTA

// Row of D = 128 * VEC floats held by a warp: element (v, lane, c) <-> column (v * 32 + lane) * 4 + c.
template <int VEC>
struct WarpRow {
    float4 v[VEC];

    __device__ __forceinline__ void load(const float* __restrict__ src, int lane) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
        for (int i = 0; i < VEC; ++i) v[i] = s4[i * 32 + lane];
    }
    __device__ __forceinline__ void load_ro(const float* __restrict__ src, int lane) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
        for (int i = 0; i < VEC; ++i) v[i] = __ldg(s4 + i * 32 + lane);
    }
    __device__ __forceinline__ void add_ro(const float* __restrict__ src, int lane) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            float4 a = __ldg(s4 + i * 32 + lane);
            v[i].x += a.x; v[i].y += a.y; v[i].z += a.z; v[i].w += a.w;
        }
    }
    // (x - mean) * rstd * gamma + beta, in place
    __device__ __forceinline__ void layernorm(const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                              int lane) {
        constexpr float inv_d = 1.0f / (128.0f * VEC);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        const float mean = warp_sum(s) * inv_d;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
        const float rstd = rsqrtf(warp_sum(q) * inv_d + eps);
        const float4* g4 = reinterpret_cast<const float4*>(gamma);
        const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float4 g = __ldg(g4 + i * 32 + lane);
            const float4 b = __ldg(b4 + i * 32 + lane);
            v[i].x = (v[i].x - mean) * rstd * g.x + b.x;
            v[i].y = (v[i].y - mean) * rstd * g.y + b.y;
            v[i].z = (v[i].z - mean) * rstd * g.z + b.z;
            v[i].w = (v[i].w - mean) * rstd * g.w + b.w;
        }
    }
    __device__ __forceinline__ void store_f32(float* __restrict__ dst, int lane) const {
        float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
        for (int i = 0; i < VEC; ++i) d4[i * 32 + lane] = v[i];
    }
    __device__ __forceinline__ void store_f16(op16_t* __restrict__ dst, int lane) const {
        uint2* d2 = reinterpret_cast<uint2*>(dst);
#pragma unroll
        for (int i = 0; i < VEC; ++i) d2[i * 32 + lane] = make_uint2(pack_op16x2(v[i].x, v[i].y), pack_op16x2(v[i].z, v[i].w));
    }
    // split-precision mode: the row as an fp16 (hi, lo) pair, hi = fp16(x), lo = fp16(x - hi)
    __device__ __forceinline__ void store_f16_split(op16_t* __restrict__ dst_hi, op16_t* __restrict__ dst_lo, int lane) const {
        uint2* h2 = reinterpret_cast<uint2*>(dst_hi);
        uint2* l2 = reinterpret_cast<uint2*>(dst_lo);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const uint32_t h0 = pack_op16x2(v[i].x, v[i].y), h1 = pack_op16x2(v[i].z, v[i].w);
            const float2 a = unpack_op16x2(h0), b = unpack_op16x2(h1);
            h2[i * 32 + lane] = make_uint2(h0, h1);
            l2[i * 32 + lane] = make_uint2(pack_op16x2(v[i].x - a.x, v[i].y - a.y), pack_op16x2(v[i].z - b.x, v[i].w - b.y));
        }
    }
};

__device__ __forceinline__ float op16_to_float(op16_t q) {
#ifdef MCM_OP_BF16
    return __bfloat162float(q);
#else
    return __half2float(q);
#endif
}

// x f32 [M, D] -> LayerNorm -> fp16 or f32 [M, D]
template <int VEC, bool OUT_F16>
__global__ void __launch_bounds__(kRowThreads)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 void* __restrict__ out, int M, float eps) {
    constexpr int D = 128 * VEC;
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (kRowThreads / 32) + (threadIdx.x >> 5);
    if (row >= M) return;
    WarpRow<VEC> r;
    r.load(x + static_cast<size_t>(row) * D, lane);
    r.layernorm(gamma, beta, eps, lane);
    if constexpr (OUT_F16)
        r.store_f16(static_cast<op16_t*>(out) + static_cast<size_t>(row) * D, lane);
    else
        r.store_f32(static_cast<float*>(out) + static_cast<size_t>(row) * D, lane);
}

// Finish the embeddings after the patch GEMM has written  patch . W + pos  into the patch rows of x:
//   row b*S      <- class_embedding + pos[0]                    (HF:212-213,217)
//   every row    <- pre_layrnorm(row)                            (HF:677)   -> x   (fp32 residual stream)
// The rows leave as the fp16 (hi, lo) pair the residual stream is kept in (xh, xh_lo; gemm_tcgen05.cuh
// EPI_BIAS_RESID_H2_*) -- hi is what the LayerNorm-folded q/k/v projection of layer 0 reads -- plus their
// (sum, sum of squares) as part 0 of the row statistics.
template <int VEC>
__global__ void __launch_bounds__(kRowThreads)
embed_finish_kernel(float* __restrict__ x, op16_t* __restrict__ xh, op16_t* __restrict__ xh_lo, float2* __restrict__ stats,
                    const float* __restrict__ cls, const float* __restrict__ pos, const float* __restrict__ pre_g,
                    const float* __restrict__ pre_b, int M, int S, float eps, int write_x) {
    constexpr int D = 128 * VEC;
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (kRowThreads / 32) + (threadIdx.x >> 5);
    if (row >= M) return;
    WarpRow<VEC> r;
    if (row % S == 0) {
        r.load_ro(cls, lane);
        r.add_ro(pos, lane);
    } else {
        r.load(x + static_cast<size_t>(row) * D, lane);
    }
    r.layernorm(pre_g, pre_b, eps, lane);
    if (write_x) r.store_f32(x + static_cast<size_t>(row) * D, lane);     // only the per-kernel test hook reads it back
    if (xh != nullptr) {
        // the residual stream from here on: an fp16 (hi, lo) pair; hi is what the q/k/v projection reads
        r.store_f16_split(xh + static_cast<size_t>(row) * D, xh_lo + static_cast<size_t>(row) * D, lane);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            s1 += (r.v[i].x + r.v[i].y) + (r.v[i].z + r.v[i].w);
            s2 += (r.v[i].x * r.v[i].x + r.v[i].y * r.v[i].y) + (r.v[i].z * r.v[i].z + r.v[i].w * r.v[i].w);
        }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        if (lane == 0) stats[row] = make_float2(s1, s2);
    }
}

// LayerNorm fold of one projection (gemm_tcgen05.cuh), one warp per output row n of W [N, K]:
//   w16[n, k] = fp16(gamma[k] * W[n, k]),  c[n] = sum_k w16[n, k] (the ROUNDED operand, so that a constant
//   row cancels exactly),  d[n] = bias[n] + sum_k beta[k] * W[n, k]
// Split-precision mode (w16_lo / c_split non-null): w16_lo = fp16(gamma * W - w16), c_split[n] = sum_k (w16 + w16_lo)[n, k].
__global__ void __launch_bounds__(256)
fold_ln_weight_kernel(const float* __restrict__ w, const float* __restrict__ gamma, const float* __restrict__ beta,
                      const float* __restrict__ bias, op16_t* __restrict__ w16, float* __restrict__ c, float* __restrict__ d,
                      int N, int K, op16_t* __restrict__ w16_lo = nullptr, float* __restrict__ c_split = nullptr) {
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (n >= N) return;
    const float* wr = w + static_cast<size_t>(n) * K;
    op16_t* o = w16 + static_cast<size_t>(n) * K;
    float sc = 0.f, sd = 0.f, sl = 0.f;
    for (int k = lane; k < K; k += 32) {
        const float v = wr[k];
        const float gw = gamma[k] * v;
        const op16_t q = to_op16(gw);
        o[k] = q;
        sc += op16_to_float(q);
        if (w16_lo != nullptr) {
            const op16_t ql = to_op16(gw - op16_to_float(q));
            w16_lo[static_cast<size_t>(n) * K + k] = ql;
            sl += op16_to_float(ql);
        }
        sd = fmaf(beta[k], v, sd);
    }
    sc = warp_sum(sc);
    sd = warp_sum(sd);
    sl = warp_sum(sl);
    if (lane == 0) {
        c[n] = sc;
        d[n] = bias[n] + sd;
        if (c_split != nullptr) c_split[n] = sc + sl;
    }
}

// Patch gather ("im2col" of the stride = kernel = patch conv, HF:148-154,209-210):
//   images f32 [b, 3, H, W] NCHW  ->  patches fp16 [b * G * G, Kp],  column = c * p * p + i * p + j
// (the flattening order of the conv weight [D, 3, p, p]); columns >= 3 p^2 (K padding) are never
// written and stay zero.  One CTA converts the 3 * p image rows that make up one row of G patches:
// 16-byte streaming reads of contiguous image rows, the fp16 patch rows are assembled in shared memory
// (G * 3 p^2 halves, <= 43 KB) and leave as contiguous 8-byte runs (3 p^2 fp16 = 6 p^2 bytes per patch).
__host__ __device__ inline int patchify_smem_bytes(int G, int p) { return G * 3 * p * p * 2; }

// lo_pass != 0 (split-precision mode, second launch): writes fp16(v - fp16(v)), the low halves of the same values.
__device__ __forceinline__ uint32_t patch_pack(float a, float b, int lo_pass) {
    const uint32_t hi = pack_op16x2(a, b);
    if (!lo_pass) return hi;
    const float2 h = unpack_op16x2(hi);
    return pack_op16x2(a - h.x, b - h.y);
}

__global__ void __launch_bounds__(256)
patchify_kernel(const float* __restrict__ img, op16_t* __restrict__ patches, int G, int p, int Kp, int lo_pass) {
    extern __shared__ __align__(16) uint8_t patchify_smem[];
    op16_t* tile = reinterpret_cast<op16_t*>(patchify_smem);      // [G][3 p^2]
    pdl_launch_dependents();
    pdl_wait();
    const int W = G * p;
    const int kpatch = 3 * p * p;
    const int gy = blockIdx.x % G;
    const int b = blockIdx.x / G;
    const int quarter_w = W >> 2;                  // W = 224: a multiple of 4 for every CLIP patch size
    const int n = 3 * p * quarter_w;               // float4 items of this patch row
    const float* src_img = img + static_cast<size_t>(b) * 3 * W * W;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int xq = t % quarter_w;
        const int ci = t / quarter_w;  // c * p + i
        const int i = ci % p;
        const int c = ci / p;
        const int x = xq * 4;
        const float4 v = __ldcs(reinterpret_cast<const float4*>(src_img + (static_cast<size_t>(c) * W + gy * p + i) * W + x));
        const float vs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; e += 2) {           // x is even and p is even: a pixel pair never straddles two patches
            const int gx = (x + e) / p;
            const int j = (x + e) - gx * p;
            *reinterpret_cast<uint32_t*>(tile + gx * kpatch + ci * p + j) = patch_pack(vs[e], vs[e + 1], lo_pass);
        }
    }
    __syncthreads();
    op16_t* dst_row = patches + (static_cast<size_t>(b) * G * G + static_cast<size_t>(gy) * G) * Kp;
    const int run8 = kpatch >> 2;                  // 8-byte items per patch
    const uint2* t2 = reinterpret_cast<const uint2*>(tile);
    for (int t = threadIdx.x; t < G * run8; t += blockDim.x) {
        const int gx = t / run8;
        const int o = t - gx * run8;
        *reinterpret_cast<uint2*>(dst_row + static_cast<size_t>(gx) * Kp + 4 * o) = t2[t];
    }
}

// uint8 ingest: the last two steps of the reference preprocess fused into the patch gather.
//   images u8 [b, H, W, 3] HWC (what PIL / a JPEG decoder holds after Resize(224) + CenterCrop(224),
//   utils/train_eval_util.py:29-34)  ->  ToTensor (x / 255)  ->  Normalize ((x - mean[c]) / std[c], :27-28)
//   ->  patches fp16 [b * G * G, Kp],  column = c * p * p + i * p + j
// fp32 arithmetic in torchvision's order (true divisions), so the fp16 patch rows are bit-identical to
// patchify_kernel run on the fp32 tensor the reference's DataLoader would have produced -- at a quarter of the
// PCIe / HBM bytes.  One CTA handles the p image rows of one row of G patches: reads are contiguous 3 W-byte
// pixel rows (12 bytes = 4 pixels per thread), the patch rows are assembled in shared memory like patchify_kernel.
struct NormConst { float mean[3], std[3]; };

__global__ void __launch_bounds__(256)
patchify_u8_kernel(const uint8_t* __restrict__ img, op16_t* __restrict__ patches, int G, int p, int Kp, const NormConst nc,
                   int lo_pass) {
    extern __shared__ __align__(16) uint8_t patchify_smem[];
    op16_t* tile = reinterpret_cast<op16_t*>(patchify_smem);      // [G][3 p^2]
    pdl_launch_dependents();
    pdl_wait();
    const int W = G * p;
    const int kpatch = 3 * p * p;
    const int gy = blockIdx.x % G;
    const int b = blockIdx.x / G;
    const int row_bytes = 3 * W;                  // one pixel row, HWC
    const uint8_t* src = img + (static_cast<size_t>(b) * W + static_cast<size_t>(gy) * p) * row_bytes;
    // item = 4 horizontally adjacent pixels (12 bytes, 4-byte aligned) of image row i
    const int quarter_w = W >> 2;
    const int n = p * quarter_w;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int xq = t % quarter_w;
        const int i = t / quarter_w;
        const int x = xq * 4;
        const uint32_t* q = reinterpret_cast<const uint32_t*>(src + static_cast<size_t>(i) * row_bytes + 3 * x);
        const uint32_t w[3] = {__ldcs(q), __ldcs(q + 1), __ldcs(q + 2)};
        float v[4][3];                            // [pixel][channel], ToTensor + Normalize in torchvision's order
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            const float u = static_cast<float>((w[k >> 2] >> (8 * (k & 3))) & 0xffu);
            v[k / 3][k % 3] = __fdiv_rn(__fsub_rn(__fdiv_rn(u, 255.0f), nc.mean[k % 3]), nc.std[k % 3]);
        }
#pragma unroll
        for (int e = 0; e < 4; e += 2) {          // x is even and p is even: a pixel pair never straddles two patches
            const int gx = (x + e) / p;
            const int j = (x + e) - gx * p;
            op16_t* d = tile + gx * kpatch + i * p + j;
#pragma unroll
            for (int c = 0; c < 3; ++c) *reinterpret_cast<uint32_t*>(d + c * p * p) = patch_pack(v[e][c], v[e + 1][c], lo_pass);
        }
    }
    __syncthreads();
    op16_t* dst_row = patches + (static_cast<size_t>(b) * G * G + static_cast<size_t>(gy) * G) * Kp;
    const int run8 = kpatch >> 2;
    const uint2* t2 = reinterpret_cast<const uint2*>(tile);
    for (int t = threadIdx.x; t < G * run8; t += blockDim.x) {
        const int gx = t / run8;
        const int o = t - gx * run8;
        *reinterpret_cast<uint2*>(dst_row + static_cast<size_t>(gx) * Kp + 4 * o) = t2[t];
    }
}

// fp32 -> fp16 with row re-striding (weight packing): dst[r * dst_ld + c] = src[r * cols + c];
// dst_lo (nullable): the low halves fp16(src - dst) of the split-precision mode
__global__ void convert_rows_f16_kernel(const float* __restrict__ src, op16_t* __restrict__ dst, op16_t* __restrict__ dst_lo,
                                         int64_t rows, int cols, int dst_ld) {
    const int64_t total = rows * cols;
    for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t r = t / cols;
        const int c = static_cast<int>(t - r * cols);
        const op16_t q = to_op16(src[t]);
        dst[r * dst_ld + c] = q;
        if (dst_lo != nullptr) dst_lo[r * dst_ld + c] = to_op16(src[t] - op16_to_float(q));
    }
}

}  // namespace mcm
