// Resize(size) + CenterCrop(size) of the reference preprocess on the device:
//   val_preprocess = Resize(224) -> CenterCrop(224) -> ToTensor -> Normalize   (utils/train_eval_util.py:29-34)
// The reference runs the first two on PIL images in the DataLoader workers; the arithmetic is Pillow's
// ImagingResample (src/libImaging/Resample.c) driven by torchvision.transforms.functional.resize / center_crop:
//   * shorter edge -> size, longer edge -> int(size * long / short); crop offset = round-half-even((new - size) / 2)
//   * separable two-pass bilinear resampling with antialiasing: horizontal pass first, each pass a triangle
//     filter of support max(scale, 1) around (x + 0.5) * scale whose normalised double coefficients become 22-bit
//     fixed point ((int)(0.5 + k * 2^22)); int32 accumulation from 1 << 21, >> 22, clip to 8 bits AFTER EACH PASS.
// Here the host computes the (bounds, coefficient) tables exactly like precompute_coeffs / normalize_coeffs_8bpc
// (same double operations in the same order; x86-64 gcc does not contract them into FMAs) -- only for the
// size output columns / rows that survive the crop -- and one kernel launch resamples a whole batch of images of
// different sizes: a CTA produces up to 16 output rows of one image, first the horizontal pass of the source rows
// those need into shared memory (uint8, rounded like Pillow's intermediate image), then the vertical pass out of it.
// Output is bit-identical to torchvision on PIL images (oracle/pil_resize_oracle.py, pinned against both).
#pragma once
#include <cmath>
#include <cstdint>
#include <map>
#include <utility>
#include <vector>

#include "ptx.cuh"

namespace mcm {

constexpr int kRcPrecisionBits = 32 - 8 - 2;       // Resample.c PRECISION_BITS
constexpr int kRcTileRows = 16;                    // output rows per CTA (fewer for strong down-scaling)
constexpr int kRcMaxSmem = 96 * 1024;              // horizontal-pass tile: source rows x size x 3 bytes

struct RcImage {
    int64_t src_off;   // byte offset of the image in the packed source buffer
    int32_t h, w;      // source size
    int32_t coef_h;    // index (int32 units) of the horizontal table: per output column [xmin, count, k[ksize_h]]
    int32_t coef_v;    // index of the vertical table: per output row [ymin, count, k[ksize_v]]
    int32_t ksize_h, ksize_v;
};
struct RcTile {
    int32_t img;       // image index
    int32_t r0, nr;    // output rows [r0, r0 + nr)
    int32_t y0, ny;    // source rows [y0, y0 + ny) the vertical pass of these output rows reads
};

// ---- host: Pillow's coefficient tables -------------------------------------------------------------------------
// out_size outputs over in_size inputs; only outputs [first, first + count) are kept.  Appends, per kept output,
// [xmin, n, k[ksize]] to `pool` and returns ksize.
inline int rc_precompute(int in_size, int out_size, int first, int count, std::vector<int32_t>& pool) {
    const double scale = static_cast<double>(static_cast<float>(in_size) - 0.0f) / out_size;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 1.0 * filterscale;          // bilinear filter: support 1.0
    const int ksize = static_cast<int>(std::ceil(support)) * 2 + 1;
    const double ss = 1.0 / filterscale;
    std::vector<double> k(ksize);
    for (int xx = first; xx < first + count; ++xx) {
        const double center = 0.0 + (xx + 0.5) * scale;
        double ww = 0.0;
        int xmin = static_cast<int>(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = static_cast<int>(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        int x = 0;
        for (; x < xmax; ++x) {
            double a = (x + xmin - center + 0.5) * ss;
            if (a < 0.0) a = -a;
            const double w = a < 1.0 ? 1.0 - a : 0.0;
            k[x] = w;
            ww += w;
        }
        for (x = 0; x < xmax; ++x)
            if (ww != 0.0) k[x] /= ww;
        for (; x < ksize; ++x) k[x] = 0;
        pool.push_back(xmin);
        pool.push_back(xmax);
        for (x = 0; x < ksize; ++x) {
            const double v = k[x] * (1 << kRcPrecisionBits);
            pool.push_back(k[x] < 0 ? static_cast<int32_t>(-0.5 + v) : static_cast<int32_t>(0.5 + v));
        }
    }
    return ksize;
}

// Python's round(): to nearest, ties to even (torchvision center_crop uses it on (new - size) / 2.0)
inline int rc_round_half_even(double v) {
    const double f = std::floor(v);
    const double d = v - f;
    if (d > 0.5) return static_cast<int>(f) + 1;
    if (d < 0.5) return static_cast<int>(f);
    return (static_cast<long long>(f) % 2 == 0) ? static_cast<int>(f) : static_cast<int>(f) + 1;
}

struct RcPlan {
    std::vector<RcImage> images;
    std::vector<RcTile> tiles;
    std::vector<int32_t> pool;
    int max_smem = 0;
};

// Plan one batch.  Returns an empty string or an error text.
inline const char* rc_plan(const int64_t* offsets, const int32_t* hs, const int32_t* ws, int n, int size, RcPlan& plan) {
    std::map<std::pair<int, int>, RcImage> by_size;      // images of equal size share their tables
    plan.images.resize(n);
    for (int i = 0; i < n; ++i) {
        const int h = hs[i], w = ws[i];
        if (h <= 0 || w <= 0) return "image with a non-positive size";
        RcImage im{};
        auto it = by_size.find({h, w});
        if (it != by_size.end()) {
            im = it->second;
        } else {
            // torchvision _compute_resized_output_size (int size): shorter edge -> size
            const int shrt = w <= h ? w : h, lng = w <= h ? h : w;
            const int new_long = static_cast<int>(static_cast<double>(size) * lng / shrt);
            const int new_w = w <= h ? size : new_long, new_h = w <= h ? new_long : size;
            const int top = rc_round_half_even((new_h - size) / 2.0), left = rc_round_half_even((new_w - size) / 2.0);
            im.h = h;
            im.w = w;
            im.coef_h = static_cast<int32_t>(plan.pool.size());
            im.ksize_h = rc_precompute(w, new_w, left, size, plan.pool);
            im.coef_v = static_cast<int32_t>(plan.pool.size());
            im.ksize_v = rc_precompute(h, new_h, top, size, plan.pool);
            by_size[{h, w}] = im;
        }
        im.src_off = offsets[i];
        plan.images[i] = im;
        // tiles: as many output rows as fit the shared-memory budget of the horizontal-pass rows
        const int32_t* cv = plan.pool.data() + im.coef_v;
        const int stride = 2 + im.ksize_v;
        for (int r0 = 0; r0 < size;) {
            int nr = size - r0 < kRcTileRows ? size - r0 : kRcTileRows;
            int y0 = cv[r0 * stride], ny = 0;
            for (;; nr = nr / 2) {
                const int last = r0 + nr - 1;
                ny = cv[last * stride] + cv[last * stride + 1] - y0;
                if (ny * size * 3 <= kRcMaxSmem || nr == 1) break;
            }
            if (ny * size * 3 > kRcMaxSmem) return "source image too large for the resampling tile (down-scaling factor above ~70)";
            plan.tiles.push_back(RcTile{i, r0, nr, y0, ny});
            if (ny * size * 3 > plan.max_smem) plan.max_smem = ny * size * 3;
            r0 += nr;
        }
    }
    return nullptr;
}

// ---- device ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint8_t rc_clip8(int acc) {
    const int v = acc >> kRcPrecisionBits;     // arithmetic shift, like Resample.c's clip8 lookup index
    return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// src: packed uint8 HWC images; dst: uint8 [n, size, size, 3]
__global__ void __launch_bounds__(256)
resize_crop_u8_kernel(const uint8_t* __restrict__ src, const RcImage* __restrict__ images, const RcTile* __restrict__ tiles,
                      const int32_t* __restrict__ pool, uint8_t* __restrict__ dst, int size) {
    extern __shared__ __align__(16) uint8_t rc_smem[];      // horizontal pass: [ny][size][3]
    const RcTile t = tiles[blockIdx.x];
    const RcImage im = images[t.img];
    const uint8_t* simg = src + im.src_off;
    // horizontal pass of the source rows this tile needs
    const int sh = 2 + im.ksize_h;
    for (int idx = threadIdx.x; idx < t.ny * size; idx += blockDim.x) {
        const int yl = idx / size, x = idx - yl * size;
        const int32_t* c = pool + im.coef_h + x * sh;
        const int xmin = c[0], cnt = c[1];
        const uint8_t* p = simg + (static_cast<size_t>(t.y0 + yl) * im.w + xmin) * 3;
        int a0 = 1 << (kRcPrecisionBits - 1), a1 = a0, a2 = a0;
        for (int i = 0; i < cnt; ++i) {
            const int k = c[2 + i];
            a0 += p[3 * i] * k;
            a1 += p[3 * i + 1] * k;
            a2 += p[3 * i + 2] * k;
        }
        uint8_t* o = rc_smem + static_cast<size_t>(idx) * 3;
        o[0] = rc_clip8(a0);
        o[1] = rc_clip8(a1);
        o[2] = rc_clip8(a2);
    }
    __syncthreads();
    // vertical pass
    const int sv = 2 + im.ksize_v;
    uint8_t* dimg = dst + static_cast<size_t>(t.img) * size * size * 3;
    for (int idx = threadIdx.x; idx < t.nr * size; idx += blockDim.x) {
        const int rl = idx / size, x = idx - rl * size;
        const int32_t* c = pool + im.coef_v + (t.r0 + rl) * sv;
        const int ymin = c[0] - t.y0, cnt = c[1];
        const uint8_t* p = rc_smem + (static_cast<size_t>(ymin) * size + x) * 3;
        int a0 = 1 << (kRcPrecisionBits - 1), a1 = a0, a2 = a0;
        for (int j = 0; j < cnt; ++j) {
            const int k = c[2 + j];
            a0 += p[0] * k;
            a1 += p[1] * k;
            a2 += p[2] * k;
            p += size * 3;
        }
        uint8_t* o = dimg + (static_cast<size_t>(t.r0 + rl) * size + x) * 3;
        o[0] = rc_clip8(a0);
        o[1] = rc_clip8(a1);
        o[2] = rc_clip8(a2);
    }
}

}  // namespace mcm
