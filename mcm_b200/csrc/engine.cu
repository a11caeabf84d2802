// mcm_b200 engine: the C ABI of include/mcm_b200.h on top of the sm_100a kernels in this directory.
//
// One handle = one CLIP vision tower resident on one B200: fp16 GEMM weights (fused QKV), fp32
// biases / LayerNorm / embeddings / projection / prompt bank, the activation workspace for
// `max_batch` images and the TMA descriptors of every GEMM operand.  A forward is a fixed sequence
// of launches on the caller's stream; nothing is allocated after mcm_create.
//
// Path per batch (reference: utils/detection_util.py:225-248 -> HF modeling_clip.py:829-863):
//   patchify -> patch GEMM(+pos) -> embed_finish(CLS, pre-LN; fp16 copy + row statistics) ->
//   L x [ QKV GEMM (LN1 folded) -> attention -> out-proj GEMM(+residual, fp16 copy, row statistics)
//         -> fc1 GEMM (LN2 folded, +quick_gelu) -> fc2 GEMM(+residual, fp16 copy, row statistics) ]
//   -> tail (post-LN, projection, MCM score)
// layer_norm1 / layer_norm2 never run as kernels: see the LayerNorm-folding note in gemm_tcgen05.cuh.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include <dlfcn.h>

#include "../../include/mcm_b200.h"
#include "attention_cls.cuh"
#include "attention_mma.cuh"
#include "attention_tcgen05.cuh"
#include "gemm_tcgen05.cuh"
#include "gemm_tcgen05_2cta.cuh"
#ifndef MCM_SWEEP_ALTERNATE
#define MCM_SWEEP_ALTERNATE 1         // 1 (default): alternate the direction the streaming kernels walk the token rows (see forward_tower); 0: A/B builds
#endif
#ifndef MCM_RESID_H2_TMA_MAX_K
#define MCM_RESID_H2_TMA_MAX_K 1024   // residual-pair GEMMs with K up to this take the all-TMA epilogue (A/B builds: 0 or 4096)
#endif
#include "resize.cuh"
#include "rowwise.cuh"
#include "tail.cuh"

using namespace mcm;

namespace {

thread_local std::string g_create_error;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

struct LayerWeights {
    op16_t *wqkv = nullptr, *wo = nullptr, *w1 = nullptr, *w2 = nullptr;   // wqkv / w1 hold gamma o W (LayerNorm fold)
    op16_t *wqkv_lo = nullptr, *wo_lo = nullptr, *w1_lo = nullptr, *w2_lo = nullptr;   // low halves (split-precision mode)
    float *wqkv32 = nullptr, *w132 = nullptr;   // fp32 masters of the two folded weights (refolded by every finalize)
    float *cqkv = nullptr, *dqkv = nullptr, *c1 = nullptr, *d1 = nullptr;   // fold vectors c, d (gemm_tcgen05.cuh)
    float *cqkv_s = nullptr, *c1_s = nullptr;   // c of the split-precision operands (row sums of hi + lo)
    float *bqkv = nullptr, *bo = nullptr, *b1 = nullptr, *b2 = nullptr;
    float *ln1g = nullptr, *ln1b = nullptr, *ln2g = nullptr, *ln2b = nullptr;
    CUtensorMap tm_wqkv, tm_wo, tm_w1, tm_w2;
    CUtensorMap tm_wqkv_lo, tm_wo_lo, tm_w1_lo, tm_w2_lo;
};

}  // namespace

struct McmHandle {
    McmConfig cfg{};
    int G = 0, Np = 0, S = 0, D = 0, H = 0, F = 0, P = 0, L = 0, Kp = 0, Kpatch = 0;
    int num_sms = 0;
    int64_t m_pad = 0, mp_pad = 0;  // padded token rows / patch rows for max_batch
    std::string err;

    // weights
    op16_t* wpatch = nullptr;  // [D, Kp]
    op16_t* wpatch_lo = nullptr;
    float *cls = nullptr, *pos = nullptr, *pre_g = nullptr, *pre_b = nullptr, *post_g = nullptr, *post_b = nullptr;
    float* wproj = nullptr;  // [P, D]
    std::vector<LayerWeights> layers;
    CUtensorMap tm_wpatch, tm_wpatch_lo;
    std::vector<uint8_t> loaded;  // one flag per expected tensor
    std::vector<std::string> expected;
    bool finalized = false;
    float* stage = nullptr;  // fp32 staging for one weight tensor
    size_t stage_elems = 0;

    // prompt bank
    float* bank = nullptr;
    int K = 0;

    // Mahalanobis baseline: whitening matrix L^T [P,P], whitened class centres [K,P], scratch [max_batch,P]
    float *maha_lt = nullptr, *maha_c = nullptr, *maha_g = nullptr;
    int maha_K = 0;
    bool maha_normalize = false;

    // workspace
    op16_t *patches = nullptr, *xh = nullptr, *qkv = nullptr, *attn = nullptr, *hid = nullptr;
    float* x = nullptr;                                       // fp32 residual stream; xh = its fp16 copy (GEMM A operand)
    float2* stats = nullptr;                                  // [stats_parts][m_pad] partial (sum, sum of squares) of the rows of x
    int stats_parts = 0;
    float* x_cls = nullptr;                                   // [pad128(max_batch), D] CLS rows of the last layer
    float *t_ln = nullptr, *t_feat = nullptr, *t_logit = nullptr;   // tail scratch: [max_batch, D | P | K]
    CUtensorMap tm_patches, tm_xh, tm_attn, tm_hid;
    // split-precision mode (MCM_OPT_PRECISION = 1): low halves of every fp16 activation buffer, allocated on first use
    int precision = MCM_PRECISION_FP16;
    op16_t *patches_lo = nullptr, *xh_lo = nullptr, *qkv_lo = nullptr, *attn_lo = nullptr, *hid_lo = nullptr;
    CUtensorMap tm_patches_lo, tm_xh_lo, tm_attn_lo, tm_hid_lo;
    CUtensorMap tm_qkv_q, tm_qkv_kv, tm_qkv_x, tm_attn_o;   // attention: 128-row Q boxes / keys_pad-row K,V boxes / 8-row boxes (tokens >= 256) over the fused QKV buffer
    bool attn_mma = false;             // debug A/B switch (env MCM_ATTN_MMA=1): warp-level mma.sync attention
    bool cls_shortcut = true;
    int sweep_desc = 0;            // direction of the NEXT streaming launch over the token rows (forward_tower alternates it)

    // uint8 ingest: Normalize constants of the reference preprocess (utils/train_eval_util.py:27-28)
    NormConst norm{{0.48145466f, 0.4578275f, 0.40821073f}, {0.26862954f, 0.26130258f, 0.27577711f}};

    // host-stream path (img_buf holds fp32 NCHW or uint8 NHWC batches: sized for the larger)
    float* img_buf[2] = {nullptr, nullptr};
    float* scores_buf = nullptr;
    int64_t scores_cap = 0;
    cudaStream_t s_copy = nullptr, s_comp = nullptr;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};

    // resize + crop ingest: double-buffered plan staging (pinned host -> device), one event per slot
    uint8_t* rc_host[2] = {nullptr, nullptr};
    uint8_t* rc_dev[2] = {nullptr, nullptr};
    size_t rc_cap[2] = {0, 0};
    cudaEvent_t rc_ev[2] = {nullptr, nullptr};
    int rc_slot = 0;
    int rc_smem_attr = 0;
    // ragged host stream (mcm_score_stream_host_images): raw packed images of one batch (two slots), resized batch
    uint8_t* raw_buf[2] = {nullptr, nullptr};
    size_t raw_cap[2] = {0, 0};
    uint8_t* rz_buf = nullptr;

    int64_t launches = 0;

    // one workspace per handle: consecutive forwards are ordered by this event, whatever streams they run on
    cudaEvent_t ev_busy = nullptr;
    bool busy_recorded = false;
    cudaStream_t busy_stream = nullptr;

    // cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: remembered per handle (= per device), not per process
    std::vector<std::pair<const void*, int>> smem_attr;
    // epilogue tensor maps (TMA stores / residual loads) by (base, rows, cols, fp32, box columns): encoded once
    std::map<std::tuple<const void*, uint64_t, uint64_t, int, uint32_t>, CUtensorMap> epi_maps;

    // MCM_OPT_CUDA_GRAPH: graphs of the forward by (images pointer, entry kind, batch, T, score kind, option set)
    bool use_graph = false;
    struct GraphKey {
        const void* images; int u8, b, mode, kind, precision, cls, K; uint32_t t_bits;
        bool operator<(const GraphKey& o) const {
            return std::tie(images, u8, b, mode, kind, precision, cls, K, t_bits) <
                   std::tie(o.images, o.u8, o.b, o.mode, o.kind, o.precision, o.cls, o.K, o.t_bits);
        }
    };
    std::map<GraphKey, std::pair<cudaGraphExec_t, int64_t>> graphs;   // (executable graph, kernels in it)
    float* t_scores = nullptr;    // [max_batch] scores of a graph replay (copied to the caller's buffer behind it)
    cudaStream_t s_capture = nullptr;

    // optional per-launch timing (mcm_profile_*): events bracket every launch of a forward
    bool prof_on = false;
    struct ProfRec { int kind; cudaEvent_t a, b; };
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;
    double prof_ms[MCM_PROF_KINDS] = {0};
    int64_t prof_n[MCM_PROF_KINDS] = {0};
};

namespace {

int fail(McmHandle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}

// Selects the handle's device for the duration of an entry point and restores the caller's current device.
struct DeviceGuard {
    int prev = -1;
    bool changed = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) changed = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (changed) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// opt in to `bytes` of dynamic shared memory for `fn` on this handle's device (once per size increase)
cudaError_t ensure_smem(McmHandle* h, const void* fn, int bytes) {
    for (auto& e : h->smem_attr)
        if (e.first == fn) {
            if (e.second >= bytes) return cudaSuccess;
            cudaError_t r = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
            if (r == cudaSuccess) e.second = bytes;
            return r;
        }
    cudaError_t r = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (r == cudaSuccess) h->smem_attr.emplace_back(fn, bytes);
    return r;
}

cudaEvent_t prof_event(McmHandle* h) {
    if (!h->prof_pool.empty()) {
        cudaEvent_t e = h->prof_pool.back();
        h->prof_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

// RAII bracket: records an event pair around one launch when profiling is on
struct ProfScope {
    McmHandle* h;
    cudaStream_t st;
    int kind;
    cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(McmHandle* h_, int kind_, cudaStream_t st_) : h(h_), st(st_), kind(kind_) {
        if (h->prof_on) {
            a = prof_event(h);
            b = prof_event(h);
            cudaEventRecord(a, st);
        }
    }
    ~ProfScope() {
        if (a) {
            cudaEventRecord(b, st);
            h->prof_recs.push_back({kind, a, b});
        }
    }
};

#define MCM_CUDA(h, call)                                                                                  \
    do {                                                                                                   \
        cudaError_t e__ = (call);                                                                          \
        if (e__ != cudaSuccess)                                                                            \
            return fail(h, MCM_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

// Row-major [rows, cols] output (or in-place residual) of a GEMM epilogue: 32-row x box_cols boxes whose smem image
// is the epilogue's staging tile (inner dimension 64 B -> SWIZZLE_64B, 128 B -> SWIZZLE_128B).
int make_tmap_epi(McmHandle* h, CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, bool f32, uint32_t box_cols) {
    const auto key = std::make_tuple(base, rows, cols, f32 ? 1 : 0, box_cols);
    auto it = h->epi_maps.find(key);
    if (it != h->epi_maps.end()) {
        *m = it->second;
        return MCM_OK;
    }
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return fail(h, MCM_ECUDA, "cuTensorMapEncodeTiled is not available from the CUDA driver");
    const uint32_t esz = f32 ? 4 : sizeof(op16_t);
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {cols * esz};
    cuuint32_t box[2] = {box_cols, 32};
    cuuint32_t estr[2] = {1, 1};
#ifdef MCM_OP_BF16
    constexpr CUtensorMapDataType kOpType = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
#else
    constexpr CUtensorMapDataType kOpType = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
#endif
    const uint32_t inner = box_cols * esz;
    if (inner != 64 && inner != 128) return fail(h, MCM_EINVAL, "epilogue box of %u bytes per row unsupported", inner);
    CUresult r = enc(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : kOpType, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, inner == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(h, MCM_ECUDA, "cuTensorMapEncodeTiled (epilogue) failed (%d) rows=%llu cols=%llu", (int)r,
                    (unsigned long long)rows, (unsigned long long)cols);
    if (h->epi_maps.size() > 4096) h->epi_maps.clear();     // callers that keep changing buffers (tests): bounded
    h->epi_maps.emplace(key, *m);
    return MCM_OK;
}

int make_tmap(McmHandle* h, CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return fail(h, MCM_ECUDA, "cuTensorMapEncodeTiled is not available from the CUDA driver");
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {cols * sizeof(op16_t)};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(kGemmBlockK), box_rows};
    cuuint32_t estr[2] = {1, 1};
#ifdef MCM_OP_BF16
    constexpr CUtensorMapDataType kOpType = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
#else
    constexpr CUtensorMapDataType kOpType = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
#endif
    CUresult r = enc(m, kOpType, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(h, MCM_ECUDA, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu box_rows=%u", (int)r,
                    (unsigned long long)rows, (unsigned long long)cols, box_rows);
    return MCM_OK;
}

// K / V boxes of the attention kernel: the fused QKV buffer [images * S, 3 D] viewed as [images][S][3 D] with boxes of
// 64 columns x box_rows tokens of ONE image; tokens >= S are out of bounds and zero-filled.
int make_tmap_kv(McmHandle* h, CUtensorMap* m, const void* base, uint64_t images, uint64_t S, uint64_t cols, uint32_t box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return fail(h, MCM_ECUDA, "cuTensorMapEncodeTiled is not available from the CUDA driver");
    cuuint64_t gdim[3] = {cols, S, images};
    cuuint64_t gstr[2] = {cols * sizeof(op16_t), S * cols * sizeof(op16_t)};
    cuuint32_t box[3] = {static_cast<cuuint32_t>(kGemmBlockK), box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
#ifdef MCM_OP_BF16
    constexpr CUtensorMapDataType kOpType = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
#else
    constexpr CUtensorMapDataType kOpType = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
#endif
    CUresult r = enc(m, kOpType, 3, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(h, MCM_ECUDA, "cuTensorMapEncodeTiled (K/V) failed (%d) images=%llu S=%llu cols=%llu box_rows=%u", (int)r,
                    (unsigned long long)images, (unsigned long long)S, (unsigned long long)cols, box_rows);
    return MCM_OK;
}

// Launch with programmatic dependent launch (every kernel of the forward chain calls pdl_wait()) and an
// optional thread-block cluster.
template <typename... KArgs, typename... Args>
cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int n = 0;
#ifdef MCM_DEBUG
    static const bool no_pdl = [] { const char* e = getenv("MCM_PDL"); return !(e && e[0] == '1'); }();   // A/B switch: measured 2-4 % slower
#else
    constexpr bool no_pdl = true;
#endif
    if (!no_pdl) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (cluster > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = cluster;
        attr[n].val.clusterDim.y = 1;
        attr[n].val.clusterDim.z = 1;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

inline int cuda_rc(McmHandle* h, cudaError_t e) {
    return e == cudaSuccess ? MCM_OK : fail(h, MCM_ECUDA, "kernel launch failed: %s", cudaGetErrorString(e));
}

inline int gemm_block_n(int N) { return (N % 256 == 0) ? 256 : 128; }

template <int BN, int EPI, bool SPLIT>
int launch_gemm2_t(McmHandle* h, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tout, const CUtensorMap& tout16,
                   const CUtensorMap& ta_lo, const CUtensorMap& tb_lo, const GemmParams& p, cudaStream_t st) {
    auto kern = gemm_f16_tn_cta2_kernel<BN, EPI, SPLIT>;
    MCM_CUDA(h, ensure_smem(h, reinterpret_cast<const void*>(kern), Gemm2Smem<BN, EPI>::kTotal));
    const int tiles = p.m_tiles * p.n_tiles;
    const int clusters = std::min(tiles, h->num_sms / 2);   // persistent: one CTA pair per TPC
    MCM_CUDA(h, launch_k(kern, dim3(2 * clusters), dim3(EpiTraits<EPI>::kThreads), Gemm2Smem<BN, EPI>::kTotal, st, 2, ta, tb, tout, tout16,
                         ta_lo, tb_lo, p));
    h->launches++;
    return MCM_OK;
}

// LayerNorm-fold side of a GEMM launch (gemm_tcgen05.cuh): consumer (EPI_LN_*) or producer (EPI_BIAS_RESID_F32_LN) fields
struct GemmLnArgs {
    const float* colsum = nullptr;
    const float2* stats_in = nullptr;
    int stats_parts = 0;
    op16_t* out16 = nullptr;
    float2* stats_out = nullptr;
    int stats_ld = 0;
    int row_len = 1;
    // split-precision mode: low halves of A / W (tensor maps) and of the fp16 outputs
    const CUtensorMap* ta_lo = nullptr;
    const CUtensorMap* tb_lo = nullptr;
    op16_t* out_lo = nullptr;
    op16_t* out16_lo = nullptr;
    // EPI_BIAS_RESID_H2_LN: the residual pair read (may alias out16 / out16_lo)
    const op16_t* resid16 = nullptr;
    const op16_t* resid16_lo = nullptr;
};

// C[M, N] = A[M, K] W[N, K]^T with fused epilogue.  M rows valid; A's tensor map covers >= ceil(M/128)*128 rows.
int launch_gemm(McmHandle* h, int prof_kind, const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, int epi,
                const float* bias, void* out, const float* resid, const float* pos, int np, int seq, cudaStream_t st,
                const GemmLnArgs& ln = GemmLnArgs()) {
    if (M <= 0) return MCM_OK;
    if (N % 128 != 0 || K % kGemmBlockK != 0)
        return fail(h, MCM_EUNSUPPORTED, "GEMM shape N=%d (multiple of 128) K=%d (multiple of 64) unsupported", N, K);
    const int bn = gemm_block_n(N);
    GemmParams p{};
    p.m_tiles = (M + kGemm2TileM - 1) / kGemm2TileM;
    p.n_tiles = N / bn;
    p.k_blocks = K / kGemmBlockK;
    p.m_valid = M;
    p.m_reverse = h->sweep_desc;
    p.ldo = N;
    p.bias = bias;
    p.out = out;
    p.resid = resid;
    p.pos = pos;
    p.np = np;
    p.seq = seq;
    p.colsum = ln.colsum;
    p.stats_in = ln.stats_in;
    p.stats_parts = ln.stats_parts;
    p.stats_ld = ln.stats_ld;
    p.inv_k = 1.0f / static_cast<float>(ln.row_len);
    p.eps = h->cfg.eps;
    p.out16 = ln.out16;
    p.stats_out = ln.stats_out;
    p.out_lo = ln.out_lo;
    p.out16_lo = ln.out16_lo;
    p.resid16 = ln.resid16;
    p.resid16_lo = ln.resid16_lo;
    const bool split = ln.ta_lo != nullptr;
    if (split && !ln.tb_lo) return fail(h, MCM_EINVAL, "split GEMM needs both low-half operands");
    // Epilogue traffic that goes through TMA (fp16 outputs; the residual epilogue of EPI_BIAS_RESID_F32_LN_TMA) needs
    // tensor maps over the caller's buffers with exactly M rows, so that the TMA unit clips the last tile.
    CUtensorMap tout = ta, tout16 = ta;   // placeholders for the kinds that do not use them
    const bool f16_out = epi == EPI_BIAS_F16 || epi == EPI_BIAS_QGELU_F16 || epi == EPI_LN_F16 || epi == EPI_LN_QGELU_F16;
    if (epi == EPI_BIAS_RESID_F32_LN) {
        // the LSU epilogue stays for long-K GEMMs (its time hides behind the main loop and it leaves 6 ring stages);
        // MCM_GEMM_RESID_TMA=0 / 1 forces one or the other (A/B runs)
#ifdef MCM_DEBUG
        static const int force = [] { const char* e = getenv("MCM_GEMM_RESID_TMA"); return e ? atoi(e) : -1; }();
#else
        constexpr int force = -1;
#endif
        const bool use_tma = !split && resid == out && (force >= 0 ? force != 0 : K <= 1024);
        if (use_tma) epi = EPI_BIAS_RESID_F32_LN_TMA;
    }
    if (epi == EPI_BIAS_RESID_H2_LN) {
        if (!ln.resid16 || !ln.resid16_lo || !ln.out16 || !ln.out16_lo || !ln.stats_out)
            return fail(h, MCM_EINVAL, "EPI_BIAS_RESID_H2_LN needs the residual pair, the output pair and the statistics buffer");
        // short-K GEMMs (out_proj) are bound by the residual traffic: all of it through TMA; long-K ones hide the LSU epilogue
        // behind the main loop and keep more ring stages; the split mode's k-loop is 3x longer anyway
        if (!split && ln.resid16 == ln.out16 && ln.resid16_lo == ln.out16_lo && K <= MCM_RESID_H2_TMA_MAX_K) epi = EPI_BIAS_RESID_H2_LN_TMA;
    }
    if (epi == EPI_BIAS_RESID_H2_LN_TMA) {
        int rc = make_tmap_epi(h, &tout, ln.out16, M, N, false, 32);
        if (rc) return rc;
        if ((rc = make_tmap_epi(h, &tout16, ln.out16_lo, M, N, false, 32))) return rc;
    } else if (f16_out && MCM_GEMM_F16_TMA_STORE) {
        int rc = make_tmap_epi(h, &tout, out, M, N, false, (MCM_GEMM_F16_CHUNK_STORE && bn == 256) ? 32 : bn / 4);
        if (rc) return rc;
    } else if (epi == EPI_BIAS_RESID_F32_LN_TMA) {
        int rc = make_tmap_epi(h, &tout, out, M, N, true, 32);
        if (rc) return rc;
        if ((rc = make_tmap_epi(h, &tout16, ln.out16, M, N, false, 32))) return rc;
    }
    p.trace = nullptr;
#ifdef MCM_GEMM_TRACE
    {   // debug build: per-launch cycle counters, printed after the launch (synchronises!)
        static long long* tr = nullptr;
        if (!tr) cudaMalloc(&tr, 8 * sizeof(long long));
        cudaMemsetAsync(tr, 0, 8 * sizeof(long long), st);
        p.trace = tr;
    }
#endif
#ifdef MCM_DEBUG
    static const int dbg_skip = [] { const char* e = getenv("MCM_GEMM_DBG_SKIP"); return e ? atoi(e) : 0; }();
    p.dbg_skip = dbg_skip;
#endif
    ProfScope prof(h, prof_kind, st);
#ifdef MCM_GEMM_TRACE
    struct TraceDump {
        McmHandle* h; long long* tr; cudaStream_t st; int M, N, K, epi;
        ~TraceDump() {
            static const bool on = getenv("MCM_GEMM_TRACE_PRINT") != nullptr;
            if (!on) return;
            long long t[8];
            cudaStreamSynchronize(st);
            cudaMemcpy(t, tr, sizeof t, cudaMemcpyDeviceToHost);
            const double nc = h->num_sms / 2, nt = (double)t[3];
            printf("GEMM_TRACE M=%d N=%d K=%d epi=%d: per tile: mma_total %.0f  wait_acc %.0f  wait_data %.0f | epi_wait %.0f epi_busy %.0f (cycles; %g tiles, %g clusters)\n",
                   M, N, K, epi, t[2] / nt, t[0] / nt, t[1] / nt, t[4] / nt, t[5] / nt, nt, nc);
            fflush(stdout);
        }
    } trace_dump{h, p.trace, st, M, N, K, epi};
#endif
    if (split) {
        // the kinds the split-precision forward uses (the TMA residual epilogue and the plain-bias fp16 kinds have no split form)
#define MCM_GEMM_SPLIT_CASE(E)                                                                                   \
    if (epi == E)                                                                                               \
        return bn == 256 ? launch_gemm2_t<256, E, true>(h, ta, tb, tout, tout16, *ln.ta_lo, *ln.tb_lo, p, st)   \
                         : launch_gemm2_t<128, E, true>(h, ta, tb, tout, tout16, *ln.ta_lo, *ln.tb_lo, p, st);
        MCM_GEMM_SPLIT_CASE(EPI_BIAS_RESID_F32)
        MCM_GEMM_SPLIT_CASE(EPI_POS_F32)
        MCM_GEMM_SPLIT_CASE(EPI_LN_F16)
        MCM_GEMM_SPLIT_CASE(EPI_LN_QGELU_F16)
        MCM_GEMM_SPLIT_CASE(EPI_BIAS_RESID_F32_LN)
        MCM_GEMM_SPLIT_CASE(EPI_BIAS_RESID_H2_LN)
#undef MCM_GEMM_SPLIT_CASE
        return fail(h, MCM_EINVAL, "GEMM epilogue %d has no split-precision form", epi);
    }
#define MCM_GEMM_CASE(E)                                                     \
    if (epi == E) return bn == 256 ? launch_gemm2_t<256, E, false>(h, ta, tb, tout, tout16, ta, tb, p, st) : launch_gemm2_t<128, E, false>(h, ta, tb, tout, tout16, ta, tb, p, st);
    MCM_GEMM_CASE(EPI_BIAS_F16)
    MCM_GEMM_CASE(EPI_BIAS_QGELU_F16)
    MCM_GEMM_CASE(EPI_BIAS_RESID_F32)
    MCM_GEMM_CASE(EPI_POS_F32)
    MCM_GEMM_CASE(EPI_LN_F16)
    MCM_GEMM_CASE(EPI_LN_QGELU_F16)
    MCM_GEMM_CASE(EPI_BIAS_RESID_F32_LN)
    MCM_GEMM_CASE(EPI_BIAS_RESID_F32_LN_TMA)
    MCM_GEMM_CASE(EPI_BIAS_RESID_H2_LN_TMA)
    MCM_GEMM_CASE(EPI_BIAS_RESID_H2_LN)
#undef MCM_GEMM_CASE
    return fail(h, MCM_EINVAL, "unknown GEMM epilogue %d", epi);
}

template <typename F>
int dispatch_vec(McmHandle* h, int D, F&& f) {
    switch (D / 128) {
        case 1: return f(std::integral_constant<int, 1>{});
        case 2: return f(std::integral_constant<int, 2>{});
        case 3: return f(std::integral_constant<int, 3>{});
        case 4: return f(std::integral_constant<int, 4>{});
        case 5: return f(std::integral_constant<int, 5>{});
        case 6: return f(std::integral_constant<int, 6>{});
        case 7: return f(std::integral_constant<int, 7>{});
        case 8: return f(std::integral_constant<int, 8>{});
        default: return fail(h, MCM_EUNSUPPORTED, "width %d unsupported (multiple of 128, <= 1024)", D);
    }
}

int launch_layernorm(McmHandle* h, const float* x, const float* g, const float* b, void* out, int M, int D, float eps,
                     bool out_f16, cudaStream_t st) {
    if (M <= 0) return MCM_OK;
    if (D % 128 != 0) return fail(h, MCM_EUNSUPPORTED, "LayerNorm width %d is not a multiple of 128", D);
    const int grid = (M + (kRowThreads / 32) - 1) / (kRowThreads / 32);
    ProfScope prof(h, MCM_PROF_LAYERNORM, st);
    int rc = dispatch_vec(h, D, [&](auto vec) {
        constexpr int V = decltype(vec)::value;
        if (out_f16)
            return cuda_rc(h, launch_k(layernorm_kernel<V, true>, dim3(grid), dim3(kRowThreads), 0, st, 1, x, g, b, out, M, eps));
        return cuda_rc(h, launch_k(layernorm_kernel<V, false>, dim3(grid), dim3(kRowThreads), 0, st, 1, x, g, b, out, M, eps));
    });
    if (rc) return rc;
    MCM_CUDA(h, cudaGetLastError());
    h->launches++;
    return MCM_OK;
}

// warp-level mma.sync attention: sequences beyond the tcgen05 kernel's 257 tokens, and (qkv_lo != nullptr) the
// three-term attention of the split-precision mode
int launch_attention_mma(McmHandle* h, const op16_t* qkv, const op16_t* qkv_lo, op16_t* out, op16_t* out_lo, int b, int S, int H,
                         cudaStream_t st) {
    const bool split = qkv_lo != nullptr;
    const int keys_pad = (S + 15) / 16 * 16;
    const size_t smem = static_cast<size_t>(split ? 4 : 2) * keys_pad * kAttnLd * sizeof(op16_t);
    if (smem > 200 * 1024) return fail(h, MCM_EUNSUPPORTED, "sequence length %d too long for the attention kernel", S);
    const void* fn = split ? reinterpret_cast<const void*>(attention_mma_kernel<true>) : reinterpret_cast<const void*>(attention_mma_kernel<false>);
    bool first = true;
    for (auto& e : h->smem_attr) first = first && e.first != fn;
    MCM_CUDA(h, ensure_smem(h, fn, (int)smem));
    if (first)   // without this the driver's default L1/shared split leaves room for ONE 60 KB CTA per SM
        MCM_CUDA(h, cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    const int mtiles = (S + 15) / 16;
    int nwarps = (mtiles + 1) / 2;
    if (nwarps > 9) nwarps = 9;
    if (nwarps < 1) nwarps = 1;
    const float scale_log2e = 0.125f * 1.4426950408889634f;  // dh^-0.5 (HF:292) * log2(e)
    ProfScope prof(h, MCM_PROF_ATTENTION, st);
    if (split)
        MCM_CUDA(h, launch_k(attention_mma_kernel<true>, dim3(b * H), dim3(nwarps * 32), smem, st, 1, qkv, qkv_lo, out, out_lo, S, H, keys_pad, scale_log2e));
    else
        MCM_CUDA(h, launch_k(attention_mma_kernel<false>, dim3(b * H), dim3(nwarps * 32), smem, st, 1, qkv, qkv_lo, out, out_lo, S, H, keys_pad, scale_log2e));
    h->launches++;
    return MCM_OK;
}

// tq / tkv: tensor maps over the fused QKV buffer with 128-row and keys_pad-row boxes
// to: 3-D map over the output [images][S][H * 64] with 32-row boxes (the O tiles leave as bulk stores, clipped at S)
int launch_attention(McmHandle* h, const CUtensorMap& tq, const CUtensorMap& tkv, const CUtensorMap& tx, const CUtensorMap& to,
                     const op16_t* qkv, op16_t* out, int b, int S, int H, cudaStream_t st) {
    if (b <= 0) return MCM_OK;
    if (h->attn_mma || S > kAtcMaxS) return launch_attention_mma(h, qkv, nullptr, out, nullptr, b, S, H, st);
    AtcParams p{};
    p.b = b;
    p.S = S;
    p.H = H;
    p.pair_mode = S <= 64 ? 1 : 0;             // two (image, head) items per 128-row unit (ViT-B/32)
    p.keys_pad = p.pair_mode ? 128 : atc_keys_pad(S);
    p.n_extra = S > 256 ? S - 256 : 0;
    p.units_per_item = ((S < 256 ? S : 256) + 127) / 128;   // query rows >= 256 go to the tail-row warp
    p.scale_log2e = 0.125f * 1.4426950408889634f;  // dh^-0.5 (HF:292) * log2(e)
    p.inv_H = 1.0f / static_cast<float>(H);
    p.reverse = h->sweep_desc;
    p.out = out;
    p.trace = nullptr;
#ifdef MCM_ATC_TRACE
    static long long* trace_dev = nullptr;
    if (!trace_dev) { cudaMalloc(&trace_dev, 3 * 16 * 8 * sizeof(long long)); }
    cudaMemsetAsync(trace_dev, 0, 3 * 16 * 8 * sizeof(long long), st);
    p.trace = trace_dev;
#endif
    const int items = p.pair_mode ? (b * H + 1) / 2 : b * H;
    const int grid = items < h->num_sms ? items : h->num_sms;
    ProfScope prof(h, MCM_PROF_ATTENTION, st);
    {
        const int smem = atc_smem_bytes(p.keys_pad);
        // softmax passes with constant trip counts for the two production shapes (attention_tcgen05.cuh, atc_two_pass)
        auto kern = attention_tcgen05_kernel<-1, false>;
        const int s_tc = S - p.n_extra;
        if (!p.pair_mode && p.keys_pad == 208 && s_tc >= 192) kern = attention_tcgen05_kernel<6, true>;
        else if (!p.pair_mode && p.keys_pad == 256 && s_tc >= 256) kern = attention_tcgen05_kernel<8, false>;
        MCM_CUDA(h, ensure_smem(h, reinterpret_cast<const void*>(kern), smem));
        MCM_CUDA(h, launch_k(kern, dim3(grid), dim3(kAtcThreads), smem, st, 1, tq, tkv, tx, to, p));
    }
#ifdef MCM_ATC_TRACE
    {
        static int dumped = 0;
        if (dumped++ == 5) {   // a warm launch
            long long t[3 * 16 * 8];
            cudaStreamSynchronize(st);
            cudaMemcpy(t, trace_dev, sizeof t, cudaMemcpyDeviceToHost);
            long long t0 = t[0];
            const char* names[3] = {"mma", "softmax_g0", "softmax_g1"};
            for (int r = 0; r < 3; ++r)
                for (int u = 0; u < 16; ++u) {
                    printf("ATC_TRACE %s unit %2d:", names[r], u);
                    for (int e = 0; e < 8; ++e) printf(" %8lld", t[(r * 16 + u) * 8 + e] ? t[(r * 16 + u) * 8 + e] - t0 : -1);
                    printf("\n");
                }
            fflush(stdout);
        }
    }
#endif
    h->launches++;
    return MCM_OK;
}

// pooled rows: fp32 `x` (x != nullptr) or the residual pair (xh, xl)
int launch_tail(McmHandle* h, const float* x, const op16_t* xh, const op16_t* xl, size_t row_stride, int b, float T, int kind,
                float* feats, float* scores, cudaStream_t st) {
    if (b <= 0) return MCM_OK;
    ProfScope prof(h, MCM_PROF_TAIL, st);
    MCM_CUDA(h, launch_k(pooled_layernorm_kernel, dim3((b + 7) / 8), dim3(256), 0, st, 1, x, xh, xl, row_stride, h->D, b, h->post_g,
                         h->post_b, h->cfg.eps, h->t_ln));
    float* f = feats ? feats : h->t_feat;
    dim3 g1((h->P + kSgemmTile - 1) / kSgemmTile, (b + kSgemmTile - 1) / kSgemmTile);
    MCM_CUDA(h, launch_k(sgemm_tn_kernel, g1, dim3(256), 0, st, 1, h->t_ln, h->wproj, f, b, h->P, h->D));
    h->launches += 2;
    if (scores) {
        dim3 g2((h->K + kSgemmTile - 1) / kSgemmTile, (b + kSgemmTile - 1) / kSgemmTile);
        MCM_CUDA(h, launch_k(sgemm_tn_kernel, g2, dim3(256), 0, st, 1, f, h->bank, h->t_logit, b, h->K, h->P));
        MCM_CUDA(h, launch_k(score_rows_kernel, dim3((b + 7) / 8), dim3(256), 0, st, 1, f, h->t_logit, h->P, h->K, b, T, kind, scores));
        h->launches += 2;
    }
    MCM_CUDA(h, cudaGetLastError());
    return MCM_OK;
}

// images: fp32 NCHW already normalised (u8 == false) or uint8 NHWC straight from the decoder (u8 == true)
int launch_embed(McmHandle* h, const void* images, bool u8, int b, cudaStream_t st, bool write_x = false) {
    const bool split = h->precision == MCM_PRECISION_SPLIT;
    {
        ProfScope prof(h, MCM_PROF_PATCHIFY, st);
        const size_t smem = patchify_smem_bytes(h->G, h->cfg.patch);
        for (int lo = 0; lo < (split ? 2 : 1); ++lo) {      // split-precision mode: a second pass writes the low halves
            op16_t* dst = lo ? h->patches_lo : h->patches;
            if (u8)
                MCM_CUDA(h, launch_k(patchify_u8_kernel, dim3(b * h->G), dim3(256), smem, st, 1, static_cast<const uint8_t*>(images), dst,
                                     h->G, h->cfg.patch, h->Kp, h->norm, lo));
            else
                MCM_CUDA(h, launch_k(patchify_kernel, dim3(b * h->G), dim3(256), smem, st, 1, static_cast<const float*>(images), dst,
                                     h->G, h->cfg.patch, h->Kp, lo));
            h->launches++;
        }
    }
    GemmLnArgs sp;
    if (split) {
        sp.ta_lo = &h->tm_patches_lo;
        sp.tb_lo = &h->tm_wpatch_lo;
    }
    int rc = launch_gemm(h, MCM_PROF_GEMM_PATCH, h->tm_patches, h->tm_wpatch, b * h->Np, h->D, h->Kp, EPI_POS_F32, nullptr, h->x, nullptr,
                         h->pos, h->Np, h->S, st, sp);
    if (rc) return rc;
    const int M = b * h->S;
    const int grid = (M + (kRowThreads / 32) - 1) / (kRowThreads / 32);
    ProfScope prof(h, MCM_PROF_EMBED_FINISH, st);
    rc = dispatch_vec(h, h->D, [&](auto vec) {
        constexpr int V = decltype(vec)::value;
        return cuda_rc(h, launch_k(embed_finish_kernel<V>, dim3(grid), dim3(kRowThreads), 0, st, 1, h->x, h->xh, h->xh_lo,
                                   h->stats, h->cls, h->pos, h->pre_g, h->pre_b, M, h->S, h->cfg.eps, write_x ? 1 : 0));
    });
    if (rc) return rc;
    MCM_CUDA(h, cudaGetLastError());
    h->launches++;
    return MCM_OK;
}

// embeddings + encoder.  Returns in *pooled / *pooled_stride where the rows the tail pools live:
// the CLS rows of h->x (stride S * D) or, with the last-layer shortcut, the compact h->x_cls (stride D).
// where the tail finds the rows it pools: fp32 rows (x != nullptr) or the residual pair (xh, xl)
struct PooledRows {
    const float* x = nullptr;
    const op16_t *xh = nullptr, *xl = nullptr;
    size_t stride = 0;
};

// embeddings + encoder.  The residual stream lives as an fp16 (hi, lo) pair in xh / xh_lo (gemm_tcgen05.cuh,
// EPI_BIAS_RESID_H2_*): hi is the A operand of the next projection, hi + lo carries ~22 significant bits.
// Returns where the rows the tail pools live: the CLS rows of the pair (stride S * D) or, with the last-layer
// shortcut, the compact fp32 h->x_cls (stride D).
int forward_tower(McmHandle* h, const void* images, bool u8, int b, cudaStream_t st, PooledRows* pooled) {
    int rc = launch_embed(h, images, u8, b, st);
    if (rc) return rc;
    const int M = b * h->S, D = h->D, F = h->F;
    const bool split = h->precision == MCM_PRECISION_SPLIT;
    pooled->x = nullptr;
    pooled->xh = h->xh;
    pooled->xl = h->xh_lo;
    pooled->stride = static_cast<size_t>(h->S) * D;
    const int ld = static_cast<int>(h->m_pad);
    // MCM_SWEEP_ALTERNATE = 1 (default): every streaming kernel of the chain qkv -> attention -> out_proj -> fc1 -> fc2 -> qkv ...
    // walks the token rows in the direction OPPOSITE to its producer's, so it starts on the rows that were written last and are
    // still in L2.  Measured (round 2, ncu --cache-control none over one step at batch 512): DRAM reads 25.89 -> 24.21 GB per
    // step (15 .. 50 MB per hand-off stay resident; every other hand-off scheme tried -- persisting access-policy windows,
    // evict-first loads of dead operands -- made the traffic worse: profiles/r02_l2_handoff_experiment.txt).  At the power cap a
    // DRAM byte is ~72 pJ, i.e. ~0.6 % of the step's energy.  Bench pairs: -1 .. -2 % in two pairs early in the round (before the
    // chunk stores), +1 % over five runs on a second box, +2.7 % in three alternating pairs on a third (27.92 / 28.12 / 27.92 vs
    // 27.52 / 27.12 / 27.13 k images/s with the final kernels).  Results are bit-identical (same digest of the verification stream).
    int dir = 1;
    auto next_dir = [&]() { h->sweep_desc = MCM_SWEEP_ALTERNATE ? dir : 0; dir ^= 1; };
    struct SweepReset { McmHandle* h; ~SweepReset() { h->sweep_desc = 0; } } sweep_reset{h};
    for (int i = 0; i < h->L; ++i) {
        const LayerWeights& w = h->layers[i];
        const bool last = i + 1 == h->L;
        // consumer side of the LayerNorm fold: xh holds the raw fp16 rows of the residual, stats their partial sums
        // (one part after embed_finish, one per 128-column half tile after an out_proj / fc2 epilogue)
        GemmLnArgs ln1, ln2, prod, out_a, fc2_a;
        ln1.colsum = split ? w.cqkv_s : w.cqkv;
        ln1.stats_in = h->stats;
        ln1.stats_parts = i == 0 ? 1 : h->stats_parts;
        ln1.stats_ld = ld;
        ln1.row_len = D;
        ln2 = ln1;
        ln2.colsum = split ? w.c1_s : w.c1;
        ln2.stats_parts = h->stats_parts;
        // producer side: the residual pair is updated in place, the row statistics are a by-product
        prod.resid16 = h->xh;
        prod.resid16_lo = h->xh_lo;
        prod.out16 = h->xh;
        prod.out16_lo = h->xh_lo;
        prod.stats_out = h->stats;
        prod.stats_ld = ld;
        if (split) {     // every operand is an fp16 (hi, lo) pair (gemm_tcgen05.cuh, "Precision modes")
            ln1.ta_lo = &h->tm_xh_lo;   ln1.tb_lo = &w.tm_wqkv_lo;  ln1.out_lo = h->qkv_lo;
            ln2.ta_lo = &h->tm_xh_lo;   ln2.tb_lo = &w.tm_w1_lo;    ln2.out_lo = h->hid_lo;
        }
        out_a = prod;
        fc2_a = prod;
        if (split) {
            out_a.ta_lo = &h->tm_attn_lo;  out_a.tb_lo = &w.tm_wo_lo;
            fc2_a.ta_lo = &h->tm_hid_lo;   fc2_a.tb_lo = &w.tm_w2_lo;
        }
        next_dir();
        if ((rc = launch_gemm(h, MCM_PROF_GEMM_QKV, h->tm_xh, w.tm_wqkv, M, 3 * D, D, EPI_LN_F16, w.dqkv, h->qkv, nullptr, nullptr, 0, 0, st, ln1))) return rc;
        if (last && h->cls_shortcut) {
            h->sweep_desc = 0;
            // only query row 0 of every image is consumed after this point (HF:685); the b CLS rows move to the fp32
            // x_cls, and rows 0 .. b-1 of xh (/ xh_lo) and of stats (free once the QKV projection has run) carry their
            // fp16 copy and statistics
            {
                ProfScope prof(h, MCM_PROF_ATTENTION, st);
                MCM_CUDA(h, launch_k(attention_cls_kernel, dim3((b * h->H + kClsWarps - 1) / kClsWarps), dim3(kClsWarps * 32), 0,
                                     st, 1, h->qkv, split ? h->qkv_lo : nullptr, h->xh, h->xh_lo, h->attn, split ? h->attn_lo : nullptr,
                                     h->x_cls, b, h->S, h->H, 0.125f));
            }
            h->launches++;
            GemmLnArgs out_c, fc2_c;         // b rows: fp32 residual rows, fp16 copy (+ low half in the split mode) for fc1
            out_c.out16 = h->xh;
            out_c.out16_lo = split ? h->xh_lo : nullptr;
            out_c.stats_out = h->stats;
            out_c.stats_ld = ld;
            out_c.ta_lo = out_a.ta_lo;
            out_c.tb_lo = out_a.tb_lo;
            fc2_c.ta_lo = fc2_a.ta_lo;       // plain residual epilogue: no fp16 copy, no statistics
            fc2_c.tb_lo = fc2_a.tb_lo;
            if ((rc = launch_gemm(h, MCM_PROF_GEMM_OUT, h->tm_attn, w.tm_wo, b, D, D, EPI_BIAS_RESID_F32_LN, w.bo, h->x_cls, h->x_cls, nullptr, 0, 0, st, out_c))) return rc;
            if ((rc = launch_gemm(h, MCM_PROF_GEMM_FC1, h->tm_xh, w.tm_w1, b, F, D, EPI_LN_QGELU_F16, w.d1, h->hid, nullptr, nullptr, 0, 0, st, ln2))) return rc;
            if ((rc = launch_gemm(h, MCM_PROF_GEMM_FC2, h->tm_hid, w.tm_w2, b, D, F, EPI_BIAS_RESID_F32, w.b2, h->x_cls, h->x_cls, nullptr, 0, 0, st, fc2_c))) return rc;
            pooled->x = h->x_cls;
            pooled->stride = D;
            break;
        }
        next_dir();
        if (split) {
            if ((rc = launch_attention_mma(h, h->qkv, h->qkv_lo, h->attn, h->attn_lo, b, h->S, h->H, st))) return rc;
        } else {
            if ((rc = launch_attention(h, h->tm_qkv_q, h->tm_qkv_kv, h->tm_qkv_x, h->tm_attn_o, h->qkv, h->attn, b, h->S, h->H, st))) return rc;
        }
        next_dir();
        if ((rc = launch_gemm(h, MCM_PROF_GEMM_OUT, h->tm_attn, w.tm_wo, M, D, D, EPI_BIAS_RESID_H2_LN, w.bo, nullptr, nullptr, nullptr, 0, 0, st, out_a))) return rc;
        next_dir();
        if ((rc = launch_gemm(h, MCM_PROF_GEMM_FC1, h->tm_xh, w.tm_w1, M, F, D, EPI_LN_QGELU_F16, w.d1, h->hid, nullptr, nullptr, 0, 0, st, ln2))) return rc;
        next_dir();
        if ((rc = launch_gemm(h, MCM_PROF_GEMM_FC2, h->tm_hid, w.tm_w2, M, D, F, EPI_BIAS_RESID_H2_LN, w.b2, nullptr, nullptr, nullptr, 0, 0, st, fc2_a))) return rc;
    }
    return MCM_OK;
}

int check_ready(McmHandle* h, int b, bool need_bank) {
    if (!h) return MCM_EINVAL;
    if (!h->finalized) return fail(h, MCM_ESTATE, "weights are not finalized (call mcm_finalize_weights)");
    if (need_bank && h->K <= 0) return fail(h, MCM_ESTATE, "text bank is not set (call mcm_set_text_bank)");
    if (b < 0 || b > h->cfg.max_batch) return fail(h, MCM_EINVAL, "batch %d outside [0, max_batch=%d]", b, h->cfg.max_batch);
    return MCM_OK;
}

template <typename T>
int dev_alloc(McmHandle* h, T** p, size_t n, bool zero) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, n * sizeof(T) + 16);
    if (e != cudaSuccess) return fail(h, MCM_ENOMEM, "cudaMalloc of %zu bytes failed: %s", n * sizeof(T), cudaGetErrorString(e));
    if (zero) {
        e = cudaMemset(q, 0, n * sizeof(T));
        if (e != cudaSuccess) return fail(h, MCM_ECUDA, "cudaMemset failed: %s", cudaGetErrorString(e));
    }
    *p = static_cast<T*>(q);
    return MCM_OK;
}

// where one HF tensor goes
struct Dest {
    float* f32 = nullptr;            // fp32 destination (also the fp32 master of a LayerNorm-folded weight), or
    op16_t* fp16 = nullptr;   // fp16 destination (converted)
    op16_t* fp16_lo = nullptr;   // its low half (split-precision mode)
    int64_t rows = 0;
    int cols = 0, dst_ld = 0;
    int slot = -1;
};

bool parse_layer_key(const char* key, int* layer, const char** rest) {
    static const char* pre = "vision_model.encoder.layers.";
    const size_t n = strlen(pre);
    if (strncmp(key, pre, n) != 0) return false;
    char* end = nullptr;
    long v = strtol(key + n, &end, 10);
    if (end == key + n || *end != '.') return false;
    *layer = static_cast<int>(v);
    *rest = end + 1;
    return true;
}

// slot numbering: 0..7 globals, then 16 per layer
enum { SLOT_CLS = 0, SLOT_PATCH, SLOT_POS, SLOT_PRE_G, SLOT_PRE_B, SLOT_POST_G, SLOT_POST_B, SLOT_PROJ, SLOT_GLOBALS };
const char* kLayerNames[16] = {"layer_norm1.weight", "layer_norm1.bias", "layer_norm2.weight", "layer_norm2.bias",
                               "self_attn.q_proj.weight", "self_attn.q_proj.bias", "self_attn.k_proj.weight",
                               "self_attn.k_proj.bias", "self_attn.v_proj.weight", "self_attn.v_proj.bias",
                               "self_attn.out_proj.weight", "self_attn.out_proj.bias", "mlp.fc1.weight", "mlp.fc1.bias",
                               "mlp.fc2.weight", "mlp.fc2.bias"};

bool resolve_key(McmHandle* h, const char* key, Dest* d) {
    const int D = h->D, F = h->F, P = h->P;
    auto f32 = [&](float* p, int64_t n, int slot) { d->f32 = p; d->rows = 1; d->cols = (int)n; d->slot = slot; return true; };
    auto b16 = [&](op16_t* p, op16_t* p_lo, int64_t rows, int cols, int ld, int slot) {
        d->fp16 = p; d->fp16_lo = p_lo; d->rows = rows; d->cols = cols; d->dst_ld = ld; d->slot = slot; return true;
    };
    if (!strcmp(key, "vision_model.embeddings.class_embedding")) return f32(h->cls, D, SLOT_CLS);
    if (!strcmp(key, "vision_model.embeddings.patch_embedding.weight")) return b16(h->wpatch, h->wpatch_lo, D, h->Kpatch, h->Kp, SLOT_PATCH);
    if (!strcmp(key, "vision_model.embeddings.position_embedding.weight")) return f32(h->pos, (int64_t)h->S * D, SLOT_POS);
    if (!strcmp(key, "vision_model.pre_layrnorm.weight")) return f32(h->pre_g, D, SLOT_PRE_G);
    if (!strcmp(key, "vision_model.pre_layrnorm.bias")) return f32(h->pre_b, D, SLOT_PRE_B);
    if (!strcmp(key, "vision_model.post_layernorm.weight")) return f32(h->post_g, D, SLOT_POST_G);
    if (!strcmp(key, "vision_model.post_layernorm.bias")) return f32(h->post_b, D, SLOT_POST_B);
    if (!strcmp(key, "visual_projection.weight")) return f32(h->wproj, (int64_t)P * D, SLOT_PROJ);
    int li = 0;
    const char* rest = nullptr;
    if (!parse_layer_key(key, &li, &rest) || li < 0 || li >= h->L) return false;
    LayerWeights& w = h->layers[li];
    int which = -1;
    for (int i = 0; i < 16; ++i)
        if (!strcmp(rest, kLayerNames[i])) which = i;
    if (which < 0) return false;
    const int slot = SLOT_GLOBALS + li * 16 + which;
    switch (which) {
        case 0: return f32(w.ln1g, D, slot);
        case 1: return f32(w.ln1b, D, slot);
        case 2: return f32(w.ln2g, D, slot);
        case 3: return f32(w.ln2b, D, slot);
        case 4: return f32(w.wqkv32, (int64_t)D * D, slot);   // q/k/v and fc1 weights are folded with their LayerNorm at finalize
        case 5: return f32(w.bqkv, D, slot);
        case 6: return f32(w.wqkv32 + (size_t)D * D, (int64_t)D * D, slot);
        case 7: return f32(w.bqkv + D, D, slot);
        case 8: return f32(w.wqkv32 + (size_t)2 * D * D, (int64_t)D * D, slot);
        case 9: return f32(w.bqkv + 2 * D, D, slot);
        case 10: return b16(w.wo, w.wo_lo, D, D, D, slot);
        case 11: return f32(w.bo, D, slot);
        case 12: return f32(w.w132, (int64_t)F * D, slot);
        case 13: return f32(w.b1, F, slot);
        case 14: return b16(w.w2, w.w2_lo, D, F, F, slot);
        case 15: return f32(w.b2, D, slot);
    }
    return false;
}

}  // namespace

extern "C" {

int32_t mcm_abi_version(void) { return MCM_ABI_VERSION; }

const char* mcm_last_error(const McmHandle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

double mcm_flops_per_image(const McmConfig* c, int32_t K) {
    if (!c || c->patch <= 0) return 0.0;
    const double G = c->image_size / c->patch, S = G * G + 1, D = c->width, F = c->mlp, P = c->proj, L = c->layers;
    const double p2 = 3.0 * c->patch * c->patch;
    return 2.0 * (S - 1) * p2 * D + L * (8.0 * S * D * D + 4.0 * S * D * F + 4.0 * S * S * D) + 2.0 * D * P + 2.0 * P * K;
}

int mcm_create(const McmConfig* cfg, McmHandle** out) {
    if (!cfg || !out) return fail(nullptr, MCM_EINVAL, "mcm_create: NULL argument");
    *out = nullptr;
    if (cfg->patch <= 0 || cfg->image_size <= 0 || cfg->image_size % cfg->patch != 0 || (cfg->patch & 1))
        return fail(nullptr, MCM_EINVAL, "image_size %d must be a multiple of an even patch size (%d)", cfg->image_size, cfg->patch);
    if (cfg->image_size % 4 != 0 || patchify_smem_bytes(cfg->image_size / cfg->patch, cfg->patch) > 48 * 1024)
        return fail(nullptr, MCM_EUNSUPPORTED, "image_size %d (multiple of 4) x patch %d: a row of patches must fit 48 KB of shared memory",
                    cfg->image_size, cfg->patch);
    if (cfg->width <= 0 || cfg->width % 128 != 0 || cfg->width > 1024)
        return fail(nullptr, MCM_EUNSUPPORTED, "width %d must be a multiple of 128, at most 1024", cfg->width);
    if (2 * (cfg->width / gemm_block_n(cfg->width)) > kMaxStatsParts)
        return fail(nullptr, MCM_EUNSUPPORTED, "width %d: the LayerNorm fold keeps at most %d partial row statistics (width must be a "
                    "multiple of 256, or at most 512)", cfg->width, kMaxStatsParts);
    if (cfg->heads <= 0 || cfg->width != cfg->heads * 64)
        return fail(nullptr, MCM_EUNSUPPORTED, "width / heads must be 64 (got %d / %d)", cfg->width, cfg->heads);
    if (cfg->mlp <= 0 || cfg->mlp % 128 != 0) return fail(nullptr, MCM_EUNSUPPORTED, "mlp %d must be a multiple of 128", cfg->mlp);
    if (cfg->proj <= 0 || cfg->proj % 16 != 0) return fail(nullptr, MCM_EUNSUPPORTED, "proj %d must be a multiple of 16", cfg->proj);
    if (cfg->image_size / cfg->patch * (cfg->image_size / cfg->patch) + 1 > kClsMaxS)
        return fail(nullptr, MCM_EUNSUPPORTED, "more than %d tokens per image are not supported", kClsMaxS);
    if (cfg->layers <= 0 || cfg->max_batch <= 0) return fail(nullptr, MCM_EINVAL, "layers and max_batch must be positive");

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, MCM_ECUDA, "no CUDA device available (%s); mcm_b200 has no CPU path", cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, MCM_EINVAL, "device %d out of range (%d devices)", cfg->device, ndev);
    cudaDeviceProp prop{};
    if ((e = cudaGetDeviceProperties(&prop, cfg->device)) != cudaSuccess)
        return fail(nullptr, MCM_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, MCM_EUNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", cfg->device,
                    prop.major, prop.minor);
    DeviceGuard guard(cfg->device);
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cur != cfg->device) return fail(nullptr, MCM_ECUDA, "cudaSetDevice(%d) failed", cfg->device);

    McmHandle* h = new McmHandle();
    h->cfg = *cfg;
    h->num_sms = prop.multiProcessorCount;
    h->G = cfg->image_size / cfg->patch;
    h->Np = h->G * h->G;
    h->S = h->Np + 1;
    h->D = cfg->width;
    h->H = cfg->heads;
    h->F = cfg->mlp;
    h->P = cfg->proj;
    h->L = cfg->layers;
    h->Kpatch = 3 * cfg->patch * cfg->patch;
    h->Kp = (h->Kpatch + kGemmBlockK - 1) / kGemmBlockK * kGemmBlockK;
    h->m_pad = (static_cast<int64_t>(cfg->max_batch) * h->S + 255) / 256 * 256;
    h->mp_pad = (static_cast<int64_t>(cfg->max_batch) * h->Np + 255) / 256 * 256;
    const int D = h->D, F = h->F;

#define MCM_TRY(expr)               \
    do {                            \
        int rc__ = (expr);          \
        if (rc__) {                 \
            g_create_error = h->err; \
            mcm_destroy(h);         \
            return rc__;            \
        }                           \
    } while (0)

    MCM_TRY(dev_alloc(h, &h->wpatch, (size_t)D * h->Kp, true));
    MCM_TRY(dev_alloc(h, &h->wpatch_lo, (size_t)D * h->Kp, true));
    {
        cudaError_t ee = cudaEventCreateWithFlags(&h->ev_busy, cudaEventDisableTiming);
        if (ee != cudaSuccess) MCM_TRY(fail(h, MCM_ECUDA, "cudaEventCreate: %s", cudaGetErrorString(ee)));
    }
    MCM_TRY(dev_alloc(h, &h->t_scores, (size_t)cfg->max_batch, false));
    MCM_TRY(dev_alloc(h, &h->cls, D, true));
    MCM_TRY(dev_alloc(h, &h->pos, (size_t)h->S * D, true));
    MCM_TRY(dev_alloc(h, &h->pre_g, D, true));
    MCM_TRY(dev_alloc(h, &h->pre_b, D, true));
    MCM_TRY(dev_alloc(h, &h->post_g, D, true));
    MCM_TRY(dev_alloc(h, &h->post_b, D, true));
    MCM_TRY(dev_alloc(h, &h->wproj, (size_t)h->P * D, true));
    h->layers.resize(h->L);
    for (auto& w : h->layers) {
        MCM_TRY(dev_alloc(h, &w.wqkv, (size_t)3 * D * D, false));
        MCM_TRY(dev_alloc(h, &w.wo, (size_t)D * D, false));
        MCM_TRY(dev_alloc(h, &w.w1, (size_t)F * D, false));
        MCM_TRY(dev_alloc(h, &w.w2, (size_t)D * F, false));
        MCM_TRY(dev_alloc(h, &w.wqkv_lo, (size_t)3 * D * D, false));
        MCM_TRY(dev_alloc(h, &w.wo_lo, (size_t)D * D, false));
        MCM_TRY(dev_alloc(h, &w.w1_lo, (size_t)F * D, false));
        MCM_TRY(dev_alloc(h, &w.w2_lo, (size_t)D * F, false));
        MCM_TRY(dev_alloc(h, &w.cqkv_s, (size_t)3 * D, false));
        MCM_TRY(dev_alloc(h, &w.c1_s, F, false));
        MCM_TRY(dev_alloc(h, &w.wqkv32, (size_t)3 * D * D, false));
        MCM_TRY(dev_alloc(h, &w.w132, (size_t)F * D, false));
        MCM_TRY(dev_alloc(h, &w.cqkv, (size_t)3 * D, false));
        MCM_TRY(dev_alloc(h, &w.dqkv, (size_t)3 * D, false));
        MCM_TRY(dev_alloc(h, &w.c1, F, false));
        MCM_TRY(dev_alloc(h, &w.d1, F, false));
        MCM_TRY(dev_alloc(h, &w.bqkv, (size_t)3 * D, false));
        MCM_TRY(dev_alloc(h, &w.bo, D, false));
        MCM_TRY(dev_alloc(h, &w.b1, F, false));
        MCM_TRY(dev_alloc(h, &w.b2, D, false));
        MCM_TRY(dev_alloc(h, &w.ln1g, D, false));
        MCM_TRY(dev_alloc(h, &w.ln1b, D, false));
        MCM_TRY(dev_alloc(h, &w.ln2g, D, false));
        MCM_TRY(dev_alloc(h, &w.ln2b, D, false));
        const uint32_t wdiv = 2;   // each CTA of a pair loads half of the W tile
        const uint32_t bnD = gemm_block_n(D) / wdiv, bn3D = gemm_block_n(3 * D) / wdiv, bnF = gemm_block_n(F) / wdiv;
        MCM_TRY(make_tmap(h, &w.tm_wqkv, w.wqkv, 3 * D, D, bn3D));
        MCM_TRY(make_tmap(h, &w.tm_wo, w.wo, D, D, bnD));
        MCM_TRY(make_tmap(h, &w.tm_w1, w.w1, F, D, bnF));
        MCM_TRY(make_tmap(h, &w.tm_w2, w.w2, D, F, bnD));
        MCM_TRY(make_tmap(h, &w.tm_wqkv_lo, w.wqkv_lo, 3 * D, D, bn3D));
        MCM_TRY(make_tmap(h, &w.tm_wo_lo, w.wo_lo, D, D, bnD));
        MCM_TRY(make_tmap(h, &w.tm_w1_lo, w.w1_lo, F, D, bnF));
        MCM_TRY(make_tmap(h, &w.tm_w2_lo, w.w2_lo, D, F, bnD));
    }
    MCM_TRY(make_tmap(h, &h->tm_wpatch, h->wpatch, D, h->Kp, gemm_block_n(D) / 2));
    MCM_TRY(make_tmap(h, &h->tm_wpatch_lo, h->wpatch_lo, D, h->Kp, gemm_block_n(D) / 2));
    h->loaded.assign(SLOT_GLOBALS + 16 * h->L, 0);
    h->stage_elems = (size_t)std::max(std::max((size_t)F * D, (size_t)D * h->Kpatch), std::max((size_t)h->S * D, (size_t)h->P * D));
    MCM_TRY(dev_alloc(h, &h->stage, h->stage_elems, false));

    // activation workspace (padding rows are zeroed once and only ever read as GEMM A rows whose
    // outputs are masked by m_valid)
    MCM_TRY(dev_alloc(h, &h->patches, (size_t)h->mp_pad * h->Kp, true));
    MCM_TRY(dev_alloc(h, &h->x, (size_t)h->m_pad * D, true));
    MCM_TRY(dev_alloc(h, &h->xh, (size_t)h->m_pad * D, true));
    MCM_TRY(dev_alloc(h, &h->xh_lo, (size_t)h->m_pad * D, true));      // the residual stream is the pair (xh, xh_lo)
    h->stats_parts = 2 * (D / gemm_block_n(D));   // one partial per half tile of a GEMM with N = D
    MCM_TRY(dev_alloc(h, &h->stats, (size_t)h->stats_parts * h->m_pad, true));
    MCM_TRY(dev_alloc(h, &h->qkv, (size_t)h->m_pad * 3 * D, true));
    MCM_TRY(dev_alloc(h, &h->attn, (size_t)h->m_pad * D, true));
    MCM_TRY(dev_alloc(h, &h->hid, (size_t)h->m_pad * F, true));
    MCM_TRY(dev_alloc(h, &h->x_cls, (size_t)((cfg->max_batch + 255) / 256 * 256) * D, true));
    MCM_TRY(dev_alloc(h, &h->t_ln, (size_t)cfg->max_batch * D, false));
    MCM_TRY(dev_alloc(h, &h->t_feat, (size_t)cfg->max_batch * h->P, false));
    MCM_TRY(make_tmap(h, &h->tm_patches, h->patches, h->mp_pad, h->Kp, kGemmBlockM));
    MCM_TRY(make_tmap(h, &h->tm_xh, h->xh, h->m_pad, D, kGemmBlockM));
    MCM_TRY(make_tmap(h, &h->tm_xh_lo, h->xh_lo, h->m_pad, D, kGemmBlockM));
    MCM_TRY(make_tmap(h, &h->tm_attn, h->attn, h->m_pad, D, kGemmBlockM));
    MCM_TRY(make_tmap(h, &h->tm_hid, h->hid, h->m_pad, F, kGemmBlockM));
    {
#ifdef MCM_DEBUG
        const char* e = getenv("MCM_ATTN_MMA");      // A/B switch: warp-level mma.sync attention for every shape
        h->attn_mma = e && e[0] == '1';
#endif
        if (h->S <= kAtcMaxS) {
            MCM_TRY(make_tmap(h, &h->tm_qkv_q, h->qkv, h->m_pad, 3 * D, 128));
            MCM_TRY(make_tmap_kv(h, &h->tm_qkv_kv, h->qkv, cfg->max_batch, h->S, 3 * D, atc_kv_box_rows(h->S)));
            MCM_TRY(make_tmap(h, &h->tm_qkv_x, h->qkv, h->m_pad, 3 * D, 8));
            MCM_TRY(make_tmap_kv(h, &h->tm_attn_o, h->attn, cfg->max_batch, h->S, D, 32));
        }
    }
#undef MCM_TRY
    *out = h;
    return MCM_OK;
}

void mcm_destroy(McmHandle* h) {
    if (!h) return;
    DeviceGuard guard(h->cfg.device);
    cudaDeviceSynchronize();
    auto fr = [](void* p) { if (p) cudaFree(p); };
    for (auto& g : h->graphs) cudaGraphExecDestroy(g.second.first);
    h->graphs.clear();
    if (h->ev_busy) cudaEventDestroy(h->ev_busy);
    if (h->s_capture) cudaStreamDestroy(h->s_capture);
    fr(h->t_scores); fr(h->wpatch_lo);
    fr(h->patches_lo); fr(h->xh_lo); fr(h->qkv_lo); fr(h->attn_lo); fr(h->hid_lo);
    fr(h->wpatch); fr(h->cls); fr(h->pos); fr(h->pre_g); fr(h->pre_b); fr(h->post_g); fr(h->post_b); fr(h->wproj);
    for (auto& w : h->layers) {
        fr(w.wqkv); fr(w.wo); fr(w.w1); fr(w.w2); fr(w.bqkv); fr(w.bo); fr(w.b1); fr(w.b2);
        fr(w.ln1g); fr(w.ln1b); fr(w.ln2g); fr(w.ln2b);
        fr(w.wqkv32); fr(w.w132); fr(w.cqkv); fr(w.dqkv); fr(w.c1); fr(w.d1);
        fr(w.wqkv_lo); fr(w.wo_lo); fr(w.w1_lo); fr(w.w2_lo); fr(w.cqkv_s); fr(w.c1_s);
    }
    fr(h->stats);
    fr(h->stage); fr(h->bank); fr(h->patches); fr(h->x); fr(h->xh); fr(h->qkv); fr(h->attn); fr(h->hid);
    fr(h->img_buf[0]); fr(h->img_buf[1]); fr(h->scores_buf);
    fr(h->x_cls); fr(h->t_ln); fr(h->t_feat); fr(h->t_logit);
    fr(h->maha_lt); fr(h->maha_c); fr(h->maha_g);
    fr(h->raw_buf[0]); fr(h->raw_buf[1]); fr(h->rz_buf);
    for (int i = 0; i < 2; ++i) {
        fr(h->rc_dev[i]);
        if (h->rc_host[i]) cudaFreeHost(h->rc_host[i]);
        if (h->rc_ev[i]) cudaEventDestroy(h->rc_ev[i]);
    }
    for (int i = 0; i < 2; ++i) {
        if (h->ev_h2d[i]) cudaEventDestroy(h->ev_h2d[i]);
        if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
    }
    for (auto& r : h->prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : h->prof_pool) cudaEventDestroy(e);
    if (h->s_copy) cudaStreamDestroy(h->s_copy);
    if (h->s_comp) cudaStreamDestroy(h->s_comp);
    delete h;
}

int mcm_load_weight(McmHandle* h, const char* key, const float* data, int64_t numel, int32_t* used) {
    if (used) *used = 0;
    if (!h || !key || !data) return fail(h, MCM_EINVAL, "mcm_load_weight: NULL argument");
    Dest d;
    if (!resolve_key(h, key, &d)) return MCM_OK;  // not part of the vision path
    const int64_t want = d.rows * d.cols;
    if (numel != want) return fail(h, MCM_EINVAL, "%s: expected %lld elements, got %lld", key, (long long)want, (long long)numel);
    DeviceGuard guard(h->cfg.device);
    if (d.f32) {
        MCM_CUDA(h, cudaMemcpy(d.f32, data, want * sizeof(float), cudaMemcpyDefault));
    } else {
        MCM_CUDA(h, cudaMemcpy(h->stage, data, want * sizeof(float), cudaMemcpyDefault));
        convert_rows_f16_kernel<<<1024, 256>>>(h->stage, d.fp16, d.fp16_lo, d.rows, d.cols, d.dst_ld);
        MCM_CUDA(h, cudaGetLastError());
        MCM_CUDA(h, cudaDeviceSynchronize());
    }
    h->loaded[d.slot] = 1;
    h->finalized = false;
    if (used) *used = 1;
    return MCM_OK;
}

int mcm_finalize_weights(McmHandle* h) {
    if (!h) return MCM_EINVAL;
    static const char* globals[SLOT_GLOBALS] = {"vision_model.embeddings.class_embedding",
                                                "vision_model.embeddings.patch_embedding.weight",
                                                "vision_model.embeddings.position_embedding.weight",
                                                "vision_model.pre_layrnorm.weight", "vision_model.pre_layrnorm.bias",
                                                "vision_model.post_layernorm.weight", "vision_model.post_layernorm.bias",
                                                "visual_projection.weight"};
    for (size_t i = 0; i < h->loaded.size(); ++i) {
        if (h->loaded[i]) continue;
        if (i < SLOT_GLOBALS) return fail(h, MCM_ESTATE, "weight %s was never loaded", globals[i]);
        const int li = (int)(i - SLOT_GLOBALS) / 16, which = (int)(i - SLOT_GLOBALS) % 16;
        return fail(h, MCM_ESTATE, "weight vision_model.encoder.layers.%d.%s was never loaded", li, kLayerNames[which]);
    }
    DeviceGuard guard(h->cfg.device);
    // LayerNorm fold (gemm_tcgen05.cuh): W' = fp16(gamma o W), c = row sums of W', d = beta @ W^T + b
    // (+ the low halves of W' and the row sums of hi + lo for the split-precision mode)
    for (auto& w : h->layers) {
        fold_ln_weight_kernel<<<(3 * h->D + 7) / 8, 256>>>(w.wqkv32, w.ln1g, w.ln1b, w.bqkv, w.wqkv, w.cqkv, w.dqkv, 3 * h->D, h->D,
                                                          w.wqkv_lo, w.cqkv_s);
        fold_ln_weight_kernel<<<(h->F + 7) / 8, 256>>>(w.w132, w.ln2g, w.ln2b, w.b1, w.w1, w.c1, w.d1, h->F, h->D, w.w1_lo, w.c1_s);
    }
    MCM_CUDA(h, cudaGetLastError());
    MCM_CUDA(h, cudaDeviceSynchronize());
    h->finalized = true;
    return MCM_OK;
}

int mcm_set_text_bank(McmHandle* h, const float* bank, int32_t K, int32_t already_unit) {
    if (!h || !bank) return fail(h, MCM_EINVAL, "mcm_set_text_bank: NULL argument");
    if (K <= 0) return fail(h, MCM_EINVAL, "K must be positive (got %d)", K);
    DeviceGuard guard(h->cfg.device);
    MCM_CUDA(h, cudaDeviceSynchronize());
    for (auto& g : h->graphs) cudaGraphExecDestroy(g.second.first);     // captured graphs point at the old bank / logits buffers
    h->graphs.clear();
    if (h->bank) { cudaFree(h->bank); h->bank = nullptr; h->K = 0; }
    if (h->t_logit) { cudaFree(h->t_logit); h->t_logit = nullptr; }
    int rc = dev_alloc(h, &h->bank, (size_t)K * h->P, false);
    if (rc) return rc;
    if ((rc = dev_alloc(h, &h->t_logit, (size_t)h->cfg.max_batch * K, false))) return rc;
    MCM_CUDA(h, cudaMemcpy(h->bank, bank, (size_t)K * h->P * sizeof(float), cudaMemcpyDefault));
    if (!already_unit) {
        normalize_rows_kernel<<<(K + 7) / 8, 256>>>(h->bank, K, h->P);
        MCM_CUDA(h, cudaGetLastError());
    }
    MCM_CUDA(h, cudaDeviceSynchronize());
    h->K = K;
    return MCM_OK;
}

}  // extern "C"

namespace {

// feats [b,P] (un-normalised projected features) -> Mahalanobis scores
int launch_maha(McmHandle* h, float* feats, int b, float* scores, cudaStream_t st) {
    ProfScope prof(h, MCM_PROF_TAIL, st);
    if (h->maha_normalize) MCM_CUDA(h, launch_k(normalize_rows_kernel, dim3((b + 7) / 8), dim3(256), 0, st, 1, feats, b, h->P));
    dim3 g1((h->P + kSgemmTile - 1) / kSgemmTile, (b + kSgemmTile - 1) / kSgemmTile);
    MCM_CUDA(h, launch_k(sgemm_tn_kernel, g1, dim3(256), 0, st, 1, static_cast<const float*>(feats), static_cast<const float*>(h->maha_lt),
                         h->maha_g, b, h->P, h->P));
    MCM_CUDA(h, launch_k(maha_min_dist_kernel, dim3((b + 7) / 8), dim3(256), 0, st, 1, static_cast<const float*>(h->maha_g),
                         static_cast<const float*>(h->maha_c), h->P, h->maha_K, b, scores));
    h->launches += h->maha_normalize ? 3 : 2;
    return MCM_OK;
}

// ---- one forward = tower + tail, ordered against the handle's previous forward, optionally replayed from a CUDA graph ----
enum FwdMode { FWD_FEATURES = 0, FWD_SCORE = 1, FWD_MAHA = 2 };

// the launches of one forward; `feats` / `scores` are where the tail writes
int enqueue_forward(McmHandle* h, const void* images, bool u8, int b, int mode, float T, int kind, float* feats, float* scores,
                    cudaStream_t st) {
    PooledRows pr;
    int rc = forward_tower(h, images, u8, b, st, &pr);
    if (rc) return rc;
    if (mode == FWD_FEATURES) return launch_tail(h, pr.x, pr.xh, pr.xl, pr.stride, b, 1.0f, SCORE_MCM, feats, nullptr, st);
    if (mode == FWD_SCORE) return launch_tail(h, pr.x, pr.xh, pr.xl, pr.stride, b, T, kind, nullptr, scores, st);
    if ((rc = launch_tail(h, pr.x, pr.xh, pr.xl, pr.stride, b, 1.0f, SCORE_MCM, h->t_feat, nullptr, st))) return rc;
    return launch_maha(h, h->t_feat, b, scores, st);
}

// Every forward of a handle uses the same activation workspace: wait for the previous one (which may have been
// enqueued on another stream -- the caller's, or the handle's own copy / compute streams) before touching it.
void forward_begin(McmHandle* h, cudaStream_t st) {
    if (h->busy_recorded && h->busy_stream != st) cudaStreamWaitEvent(st, h->ev_busy, 0);
}
void forward_end(McmHandle* h, cudaStream_t st) {
    cudaEventRecord(h->ev_busy, st);
    h->busy_recorded = true;
    h->busy_stream = st;
}

int run_forward(McmHandle* h, const void* images, bool u8, int b, int mode, float T, int kind, float* out, cudaStream_t st) {
    forward_begin(h, st);
    int rc = MCM_OK;
    const size_t out_elems = mode == FWD_FEATURES ? (size_t)b * h->P : (size_t)b;
    if (h->use_graph && !h->prof_on) {
        // the graph writes into handle-owned buffers (t_feat / t_scores); a device-to-device copy behind it delivers
        // the result, so one graph serves every destination pointer
        float* g_out = mode == FWD_FEATURES ? h->t_feat : h->t_scores;
        uint32_t t_bits;
        memcpy(&t_bits, &T, sizeof t_bits);
        const McmHandle::GraphKey key{images, u8 ? 1 : 0, b, mode, kind, h->precision, h->cls_shortcut ? 1 : 0, h->K, t_bits};
        auto it = h->graphs.find(key);
        if (it == h->graphs.end()) {
            // warm pass outside the capture: first-use attribute calls / tensor-map encodes happen here
            rc = enqueue_forward(h, images, u8, b, mode, T, kind, mode == FWD_FEATURES ? g_out : nullptr,
                                 mode == FWD_FEATURES ? nullptr : g_out, st);
            if (rc) return rc;
            cudaGraph_t graph = nullptr;
            cudaGraphExec_t exec = nullptr;
            int64_t n_kernels = 0;
            // capture on a stream of the handle's own: the caller's stream may be the legacy default stream (PyTorch's
            // default), which cannot be captured; the nodes carry no stream, the replay goes to the caller's
            cudaError_t e = cudaSuccess;
            if (!h->s_capture) e = cudaStreamCreateWithFlags(&h->s_capture, cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaStreamBeginCapture(h->s_capture, cudaStreamCaptureModeThreadLocal);
            if (e == cudaSuccess) {
                const int64_t launches0 = h->launches;
                rc = enqueue_forward(h, images, u8, b, mode, T, kind, mode == FWD_FEATURES ? g_out : nullptr,
                                     mode == FWD_FEATURES ? nullptr : g_out, h->s_capture);
                n_kernels = h->launches - launches0;
                h->launches = launches0;      // recorded, not launched
                e = cudaStreamEndCapture(h->s_capture, &graph);
                if (rc == MCM_OK && e == cudaSuccess) e = cudaGraphInstantiate(&exec, graph, 0);
                if (graph) cudaGraphDestroy(graph);
            }
            if (rc) return rc;
            if (e != cudaSuccess) return fail(h, MCM_ECUDA, "CUDA graph capture of the forward failed: %s", cudaGetErrorString(e));
            if (h->graphs.size() >= 64) {     // callers that never reuse an input buffer: bounded
                for (auto& g : h->graphs) cudaGraphExecDestroy(g.second.first);
                h->graphs.clear();
            }
            it = h->graphs.emplace(key, std::make_pair(exec, n_kernels)).first;
        }
        MCM_CUDA(h, cudaGraphLaunch(it->second.first, st));
        h->launches += it->second.second;
        MCM_CUDA(h, cudaMemcpyAsync(out, g_out, out_elems * sizeof(float), cudaMemcpyDeviceToDevice, st));
    } else {
        rc = enqueue_forward(h, images, u8, b, mode, T, kind, mode == FWD_FEATURES ? out : nullptr, mode == FWD_FEATURES ? nullptr : out, st);
    }
    forward_end(h, st);
    return rc;
}

int image_features_any(McmHandle* h, const void* images, bool u8, int32_t b, float* feats, void* stream) {
    int rc = check_ready(h, b, false);
    if (rc) return rc;
    if (b == 0) return MCM_OK;
    if (!images || !feats) return fail(h, MCM_EINVAL, "mcm_image_features: NULL buffer");
    DeviceGuard guard(h->cfg.device);
    return run_forward(h, images, u8, b, FWD_FEATURES, 1.0f, SCORE_MCM, feats, static_cast<cudaStream_t>(stream));
}

int score_any(McmHandle* h, const void* images, bool u8, int32_t b, float T, int32_t kind, float* scores, void* stream) {
    int rc = check_ready(h, b, true);
    if (rc) return rc;
    if (b == 0) return MCM_OK;
    if (!images || !scores) return fail(h, MCM_EINVAL, "mcm_score: NULL buffer");
    if (!(T > 0.f)) return fail(h, MCM_EINVAL, "temperature must be positive (got %g)", (double)T);
    if (kind < MCM_SCORE_MCM || kind > MCM_SCORE_VAR) return fail(h, MCM_EINVAL, "unknown score kind %d", kind);
    DeviceGuard guard(h->cfg.device);
    return run_forward(h, images, u8, b, FWD_SCORE, T, kind, scores, static_cast<cudaStream_t>(stream));
}

int score_stream_host_any(McmHandle* h, const void* images_host_v, bool u8, int64_t n, int32_t batch, float T, int32_t kind,
                          float* scores_host);

}  // namespace

extern "C" {

int mcm_image_features(McmHandle* h, const float* images, int32_t b, float* feats, void* stream) {
    return image_features_any(h, images, false, b, feats, stream);
}
int mcm_image_features_u8(McmHandle* h, const uint8_t* images, int32_t b, float* feats, void* stream) {
    return image_features_any(h, images, true, b, feats, stream);
}
int mcm_score(McmHandle* h, const float* images, int32_t b, float T, int32_t kind, float* scores, void* stream) {
    return score_any(h, images, false, b, T, kind, scores, stream);
}
int mcm_score_u8(McmHandle* h, const uint8_t* images, int32_t b, float T, int32_t kind, float* scores, void* stream) {
    return score_any(h, images, true, b, T, kind, scores, stream);
}
int mcm_score_stream_host(McmHandle* h, const float* images_host, int64_t n, int32_t batch, float T, int32_t kind,
                          float* scores_host) {
    return score_stream_host_any(h, images_host, false, n, batch, T, kind, scores_host);
}
int mcm_score_stream_host_u8(McmHandle* h, const uint8_t* images_host, int64_t n, int32_t batch, float T, int32_t kind,
                             float* scores_host) {
    return score_stream_host_any(h, images_host, true, n, batch, T, kind, scores_host);
}

int mcm_resize_crop_u8(McmHandle* h, const uint8_t* src, const int64_t* offsets, const int32_t* hs, const int32_t* ws, int32_t n,
                       uint8_t* dst, void* stream) {
    if (!h) return MCM_EINVAL;
    if (n == 0) return MCM_OK;
    if (n < 0 || !src || !offsets || !hs || !ws || !dst) return fail(h, MCM_EINVAL, "mcm_resize_crop_u8: bad argument");
    DeviceGuard guard(h->cfg.device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int size = h->cfg.image_size;
    RcPlan plan;
    if (const char* msg = rc_plan(offsets, hs, ws, n, size, plan)) return fail(h, MCM_EUNSUPPORTED, "mcm_resize_crop_u8: %s", msg);
    auto up16 = [](size_t v) { return (v + 15) / 16 * 16; };
    const size_t b_img = up16(plan.images.size() * sizeof(RcImage)), b_tile = up16(plan.tiles.size() * sizeof(RcTile));
    const size_t b_pool = up16(plan.pool.size() * sizeof(int32_t)), total = b_img + b_tile + b_pool;
    const int slot = h->rc_slot;
    h->rc_slot ^= 1;
    if (!h->rc_ev[slot]) MCM_CUDA(h, cudaEventCreateWithFlags(&h->rc_ev[slot], cudaEventDisableTiming));
    MCM_CUDA(h, cudaEventSynchronize(h->rc_ev[slot]));      // the copy that last used this slot's host buffer has finished
    if (total > h->rc_cap[slot]) {
        if (h->rc_host[slot]) cudaFreeHost(h->rc_host[slot]);
        if (h->rc_dev[slot]) cudaFree(h->rc_dev[slot]);
        h->rc_host[slot] = nullptr;
        h->rc_dev[slot] = nullptr;
        h->rc_cap[slot] = 0;
        const size_t cap = total * 2;
        MCM_CUDA(h, cudaMallocHost(reinterpret_cast<void**>(&h->rc_host[slot]), cap));
        int rc = dev_alloc(h, &h->rc_dev[slot], cap, false);
        if (rc) return rc;
        h->rc_cap[slot] = cap;
    }
    uint8_t* hb = h->rc_host[slot];
    memcpy(hb, plan.images.data(), plan.images.size() * sizeof(RcImage));
    memcpy(hb + b_img, plan.tiles.data(), plan.tiles.size() * sizeof(RcTile));
    memcpy(hb + b_img + b_tile, plan.pool.data(), plan.pool.size() * sizeof(int32_t));
    uint8_t* db = h->rc_dev[slot];
    MCM_CUDA(h, cudaMemcpyAsync(db, hb, total, cudaMemcpyHostToDevice, st));
    MCM_CUDA(h, cudaEventRecord(h->rc_ev[slot], st));
    if (plan.max_smem > h->rc_smem_attr) {
        MCM_CUDA(h, cudaFuncSetAttribute(resize_crop_u8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRcMaxSmem));
        h->rc_smem_attr = kRcMaxSmem;
    }
    resize_crop_u8_kernel<<<static_cast<unsigned>(plan.tiles.size()), 256, plan.max_smem, st>>>(
        src, reinterpret_cast<const RcImage*>(db), reinterpret_cast<const RcTile*>(db + b_img),
        reinterpret_cast<const int32_t*>(db + b_img + b_tile), dst, size);
    MCM_CUDA(h, cudaGetLastError());
    h->launches++;
    return MCM_OK;
}

int mcm_dbg_resize_tables(int32_t h, int32_t w, int32_t size, int32_t* ksize2, int32_t* table_h, int32_t* table_v, int32_t cap) {
    if (h <= 0 || w <= 0 || size <= 0 || !ksize2 || !table_h || !table_v) return MCM_EINVAL;
    RcPlan plan;
    const int64_t off = 0;
    if (rc_plan(&off, &h, &w, 1, size, plan)) return MCM_EUNSUPPORTED;
    const RcImage& im = plan.images[0];
    ksize2[0] = im.ksize_h;
    ksize2[1] = im.ksize_v;
    const int nh = size * (2 + im.ksize_h), nv = size * (2 + im.ksize_v);
    if (nh > cap || nv > cap) return MCM_ENOMEM;
    memcpy(table_h, plan.pool.data() + im.coef_h, nh * sizeof(int32_t));
    memcpy(table_v, plan.pool.data() + im.coef_v, nv * sizeof(int32_t));
    return MCM_OK;
}


int mcm_set_maha(McmHandle* h, const float* lt, const float* centres, int32_t K, int32_t normalize) {
    if (!h || !lt || !centres) return fail(h, MCM_EINVAL, "mcm_set_maha: NULL argument");
    if (K <= 0) return fail(h, MCM_EINVAL, "K must be positive (got %d)", K);
    DeviceGuard guard(h->cfg.device);
    MCM_CUDA(h, cudaDeviceSynchronize());
    auto fr = [](float*& p) { if (p) cudaFree(p); p = nullptr; };
    fr(h->maha_lt); fr(h->maha_c); fr(h->maha_g);
    h->maha_K = 0;
    int rc;
    if ((rc = dev_alloc(h, &h->maha_lt, (size_t)h->P * h->P, false))) return rc;
    if ((rc = dev_alloc(h, &h->maha_c, (size_t)K * h->P, false))) return rc;
    if ((rc = dev_alloc(h, &h->maha_g, (size_t)h->cfg.max_batch * h->P, false))) return rc;
    MCM_CUDA(h, cudaMemcpy(h->maha_lt, lt, (size_t)h->P * h->P * sizeof(float), cudaMemcpyDefault));
    MCM_CUDA(h, cudaMemcpy(h->maha_c, centres, (size_t)K * h->P * sizeof(float), cudaMemcpyDefault));
    h->maha_K = K;
    h->maha_normalize = normalize != 0;
    return MCM_OK;
}

int mcm_maha_score(McmHandle* h, const float* images, int32_t b, float* scores, void* stream) {
    int rc = check_ready(h, b, false);
    if (rc) return rc;
    if (h->maha_K <= 0) return fail(h, MCM_ESTATE, "Mahalanobis statistics are not set (call mcm_set_maha)");
    if (b == 0) return MCM_OK;
    if (!images || !scores) return fail(h, MCM_EINVAL, "mcm_maha_score: NULL buffer");
    DeviceGuard guard(h->cfg.device);
    return run_forward(h, images, false, b, FWD_MAHA, 1.0f, SCORE_MCM, scores, static_cast<cudaStream_t>(stream));
}

int mcm_dbg_maha_from_features(McmHandle* h, const float* feats, int32_t b, float* scores, void* stream) {
    if (!h || !feats || !scores) return fail(h, MCM_EINVAL, "mcm_dbg_maha_from_features: NULL argument");
    if (h->maha_K <= 0) return fail(h, MCM_ESTATE, "Mahalanobis statistics are not set (call mcm_set_maha)");
    if (b <= 0 || b > h->cfg.max_batch) return fail(h, MCM_EINVAL, "batch %d outside (0, max_batch=%d]", b, h->cfg.max_batch);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MCM_CUDA(h, cudaMemcpyAsync(h->t_feat, feats, (size_t)b * h->P * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return launch_maha(h, h->t_feat, b, scores, st);
}

int mcm_score_stream_host_images(McmHandle* h, const uint8_t* packed_host, const int64_t* offsets, const int32_t* hs, const int32_t* ws,
                                 int64_t n, int32_t batch, float T, int32_t kind, float* scores_host) {
    int rc = check_ready(h, batch, true);
    if (rc) return rc;
    if (n == 0) return MCM_OK;
    if (n < 0 || batch <= 0) return fail(h, MCM_EINVAL, "n must be >= 0 and batch positive");
    if (!packed_host || !offsets || !hs || !ws || !scores_host) return fail(h, MCM_EINVAL, "mcm_score_stream_host_images: NULL argument");
    DeviceGuard guard(h->cfg.device);
    const size_t out_elems = (size_t)3 * h->cfg.image_size * h->cfg.image_size;
    if (!h->s_copy) {
        MCM_CUDA(h, cudaStreamCreateWithFlags(&h->s_copy, cudaStreamNonBlocking));
        MCM_CUDA(h, cudaStreamCreateWithFlags(&h->s_comp, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            MCM_CUDA(h, cudaEventCreateWithFlags(&h->ev_h2d[i], cudaEventDisableTiming));
            MCM_CUDA(h, cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
            if ((rc = dev_alloc(h, &h->img_buf[i], (size_t)h->cfg.max_batch * out_elems, false))) return rc;
        }
    }
    if (!h->rz_buf && (rc = dev_alloc(h, &h->rz_buf, (size_t)h->cfg.max_batch * out_elems, false))) return rc;
    if (n > h->scores_cap) {
        if (h->scores_buf) cudaFree(h->scores_buf);
        h->scores_buf = nullptr;
        h->scores_cap = 0;
        if ((rc = dev_alloc(h, &h->scores_buf, (size_t)n, false))) return rc;
        h->scores_cap = n;
    }
    std::vector<int64_t> rel(static_cast<size_t>(batch));
    int64_t done = 0;
    int it = 0;
    while (done < n) {
        const int cur = static_cast<int>(std::min<int64_t>(batch, n - done));
        const int slot = it & 1;
        // byte span of this batch in the packed host buffer (images need not be in offset order)
        int64_t lo = offsets[done], hi = 0;
        for (int i = 0; i < cur; ++i) {
            const int64_t o = offsets[done + i], e = o + (int64_t)hs[done + i] * ws[done + i] * 3;
            if (o < 0 || hs[done + i] <= 0 || ws[done + i] <= 0) return fail(h, MCM_EINVAL, "image %lld: bad offset / size", (long long)(done + i));
            lo = std::min(lo, o);
            hi = std::max(hi, e);
        }
        for (int i = 0; i < cur; ++i) rel[i] = offsets[done + i] - lo;
        const size_t span = static_cast<size_t>(hi - lo);
        if (it >= 2) MCM_CUDA(h, cudaStreamWaitEvent(h->s_copy, h->ev_done[slot], 0));
        if (span > h->raw_cap[slot]) {      // grow: the slot's previous user has to be done with the old buffer
            if (it >= 2) MCM_CUDA(h, cudaEventSynchronize(h->ev_done[slot]));
            if (h->raw_buf[slot]) cudaFree(h->raw_buf[slot]);
            h->raw_buf[slot] = nullptr;
            h->raw_cap[slot] = 0;
            if ((rc = dev_alloc(h, &h->raw_buf[slot], span + span / 4, false))) return rc;
            h->raw_cap[slot] = span + span / 4;
        }
        MCM_CUDA(h, cudaMemcpyAsync(h->raw_buf[slot], packed_host + lo, span, cudaMemcpyHostToDevice, h->s_copy));
        MCM_CUDA(h, cudaEventRecord(h->ev_h2d[slot], h->s_copy));
        MCM_CUDA(h, cudaStreamWaitEvent(h->s_comp, h->ev_h2d[slot], 0));
        if ((rc = mcm_resize_crop_u8(h, h->raw_buf[slot], rel.data(), hs + done, ws + done, cur, h->rz_buf, h->s_comp))) return rc;
        if ((rc = score_any(h, h->rz_buf, true, cur, T, kind, h->scores_buf + done, h->s_comp))) return rc;
        MCM_CUDA(h, cudaEventRecord(h->ev_done[slot], h->s_comp));
        done += cur;
        ++it;
    }
    MCM_CUDA(h, cudaMemcpyAsync(scores_host, h->scores_buf, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, h->s_comp));
    MCM_CUDA(h, cudaStreamSynchronize(h->s_comp));
    return MCM_OK;
}

int mcm_set_normalization(McmHandle* h, const float* mean3, const float* std3) {
    if (!h || !mean3 || !std3) return fail(h, MCM_EINVAL, "mcm_set_normalization: NULL argument");
    for (int c = 0; c < 3; ++c)      // validate everything before touching the state
        if (!(std3[c] > 0.f)) return fail(h, MCM_EINVAL, "std[%d] must be positive (got %g)", c, (double)std3[c]);
    for (int c = 0; c < 3; ++c) {
        h->norm.mean[c] = mean3[c];
        h->norm.std[c] = std3[c];
    }
    return MCM_OK;
}

}  // extern "C"

namespace {

int score_stream_host_any(McmHandle* h, const void* images_host_v, bool u8, int64_t n, int32_t batch, float T, int32_t kind,
                          float* scores_host) {
    int rc = check_ready(h, batch, true);
    if (rc) return rc;
    if (n == 0) return MCM_OK;
    if (n < 0 || batch <= 0) return fail(h, MCM_EINVAL, "n must be >= 0 and batch positive");
    if (!images_host_v || !scores_host) return fail(h, MCM_EINVAL, "mcm_score_stream_host: NULL buffer");
    DeviceGuard guard(h->cfg.device);
    const size_t img_elems = (size_t)3 * h->cfg.image_size * h->cfg.image_size;
    const size_t esz = u8 ? 1 : sizeof(float);
    const uint8_t* images_host = static_cast<const uint8_t*>(images_host_v);
    if (!h->s_copy) {
        MCM_CUDA(h, cudaStreamCreateWithFlags(&h->s_copy, cudaStreamNonBlocking));
        MCM_CUDA(h, cudaStreamCreateWithFlags(&h->s_comp, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            MCM_CUDA(h, cudaEventCreateWithFlags(&h->ev_h2d[i], cudaEventDisableTiming));
            MCM_CUDA(h, cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
            if ((rc = dev_alloc(h, &h->img_buf[i], (size_t)h->cfg.max_batch * img_elems, false))) return rc;
        }
    }
    if (n > h->scores_cap) {
        if (h->scores_buf) cudaFree(h->scores_buf);
        h->scores_buf = nullptr;
        h->scores_cap = 0;
        if ((rc = dev_alloc(h, &h->scores_buf, (size_t)n, false))) return rc;
        h->scores_cap = n;
    }
    int64_t done = 0;
    int it = 0;
    while (done < n) {
        const int cur = static_cast<int>(std::min<int64_t>(batch, n - done));
        const int slot = it & 1;
        if (it >= 2) MCM_CUDA(h, cudaStreamWaitEvent(h->s_copy, h->ev_done[slot], 0));
        MCM_CUDA(h, cudaMemcpyAsync(h->img_buf[slot], images_host + (size_t)done * img_elems * esz, (size_t)cur * img_elems * esz,
                                    cudaMemcpyHostToDevice, h->s_copy));
        MCM_CUDA(h, cudaEventRecord(h->ev_h2d[slot], h->s_copy));
        MCM_CUDA(h, cudaStreamWaitEvent(h->s_comp, h->ev_h2d[slot], 0));
        if ((rc = score_any(h, h->img_buf[slot], u8, cur, T, kind, h->scores_buf + done, h->s_comp))) return rc;
        MCM_CUDA(h, cudaEventRecord(h->ev_done[slot], h->s_comp));
        done += cur;
        ++it;
    }
    MCM_CUDA(h, cudaMemcpyAsync(scores_host, h->scores_buf, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, h->s_comp));
    MCM_CUDA(h, cudaStreamSynchronize(h->s_comp));
    return MCM_OK;
}

}  // namespace

extern "C" {

int64_t mcm_launch_count(const McmHandle* h) { return h ? h->launches : 0; }
void mcm_reset_launch_count(McmHandle* h) { if (h) h->launches = 0; }

int mcm_set_option(McmHandle* h, int32_t option, int32_t value) {
    if (!h) return MCM_EINVAL;
    switch (option) {
        case MCM_OPT_CLS_SHORTCUT: h->cls_shortcut = value != 0; return MCM_OK;
        case MCM_OPT_CUDA_GRAPH: h->use_graph = value != 0; return MCM_OK;
        case MCM_OPT_PRECISION: {
            if (value != MCM_PRECISION_FP16 && value != MCM_PRECISION_SPLIT) return fail(h, MCM_EINVAL, "unknown precision mode %d", value);
            if (value == MCM_PRECISION_SPLIT && !h->hid_lo) {
                // low halves of the fp16 activation buffers + their tensor maps, on first use
                DeviceGuard guard(h->cfg.device);
                MCM_CUDA(h, cudaDeviceSynchronize());
                const int D = h->D;
                int rc;
                if ((rc = dev_alloc(h, &h->patches_lo, (size_t)h->mp_pad * h->Kp, true))) return rc;
                if ((rc = dev_alloc(h, &h->qkv_lo, (size_t)h->m_pad * 3 * D, true))) return rc;
                if ((rc = dev_alloc(h, &h->attn_lo, (size_t)h->m_pad * D, true))) return rc;
                if ((rc = make_tmap(h, &h->tm_patches_lo, h->patches_lo, h->mp_pad, h->Kp, kGemmBlockM))) return rc;
                if ((rc = make_tmap(h, &h->tm_attn_lo, h->attn_lo, h->m_pad, D, kGemmBlockM))) return rc;
                op16_t* hl = nullptr;
                if ((rc = dev_alloc(h, &hl, (size_t)h->m_pad * h->F, true))) return rc;
                if ((rc = make_tmap(h, &h->tm_hid_lo, hl, h->m_pad, h->F, kGemmBlockM))) { cudaFree(hl); return rc; }
                h->hid_lo = hl;     // set last: marks the whole set as complete
            }
            h->precision = value;
            return MCM_OK;
        }
        default: return fail(h, MCM_EINVAL, "unknown option %d", option);
    }
}

int mcm_profile_enable(McmHandle* h, int32_t on) {
    if (!h) return MCM_EINVAL;
    h->prof_on = on != 0;
    return MCM_OK;
}

int mcm_profile_read(McmHandle* h, double* ms, int64_t* counts, int32_t reset) {
    if (!h) return MCM_EINVAL;
    DeviceGuard guard(h->cfg.device);
    MCM_CUDA(h, cudaDeviceSynchronize());
    for (auto& r : h->prof_recs) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess && r.kind >= 0 && r.kind < MCM_PROF_KINDS) {
            h->prof_ms[r.kind] += t;
            h->prof_n[r.kind] += 1;
        }
        h->prof_pool.push_back(r.a);
        h->prof_pool.push_back(r.b);
    }
    h->prof_recs.clear();
    for (int i = 0; i < MCM_PROF_KINDS; ++i) {
        if (ms) ms[i] = h->prof_ms[i];
        if (counts) counts[i] = h->prof_n[i];
        if (reset) { h->prof_ms[i] = 0; h->prof_n[i] = 0; }
    }
    return MCM_OK;
}

// ------------------------------------------------------------------ per-kernel entry points ----
int mcm_dbg_gemm(McmHandle* h, const void* a, const void* w, const float* bias, const float* resid, void* out, int32_t M,
                 int32_t N, int32_t K, int32_t epi, void* stream) {
    if (!h || !a || !w || !out) return fail(h, MCM_EINVAL, "mcm_dbg_gemm: NULL argument");
    if (epi < 0 || epi > 2) return fail(h, MCM_EINVAL, "mcm_dbg_gemm: epi must be 0, 1 or 2");
    if (M <= 0) return fail(h, MCM_EINVAL, "mcm_dbg_gemm: M must be positive");
    if (epi == 2 && !resid) return fail(h, MCM_EINVAL, "mcm_dbg_gemm: epi 2 needs resid");
    if (N % 128 != 0 || K % 64 != 0) return fail(h, MCM_EUNSUPPORTED, "mcm_dbg_gemm: N %% 128 and K %% 64 must be 0");
    CUtensorMap ta, tb;
    int rc;
    if ((rc = make_tmap(h, &ta, a, M, K, kGemmBlockM))) return rc;
    if ((rc = make_tmap(h, &tb, w, N, K, gemm_block_n(N) / 2))) return rc;
    return launch_gemm(h, MCM_PROF_GEMM_OTHER, ta, tb, M, N, K, epi, bias, out, resid, nullptr, 0, 0, static_cast<cudaStream_t>(stream));
}

int mcm_dbg_gemm_split(McmHandle* h, const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo, const float* bias,
                       const float* resid, float* out, int32_t M, int32_t N, int32_t K, void* stream) {
    if (!h || !a_hi || !a_lo || !w_hi || !w_lo || !bias || !resid || !out) return fail(h, MCM_EINVAL, "mcm_dbg_gemm_split: NULL argument");
    if (M <= 0) return fail(h, MCM_EINVAL, "mcm_dbg_gemm_split: M must be positive");
    if (N % 128 != 0 || K % 64 != 0) return fail(h, MCM_EUNSUPPORTED, "mcm_dbg_gemm_split: N %% 128 and K %% 64 must be 0");
    DeviceGuard guard(h->cfg.device);
    CUtensorMap ta, tb, tal, tbl;
    int rc;
    if ((rc = make_tmap(h, &ta, a_hi, M, K, kGemmBlockM))) return rc;
    if ((rc = make_tmap(h, &tal, a_lo, M, K, kGemmBlockM))) return rc;
    if ((rc = make_tmap(h, &tb, w_hi, N, K, gemm_block_n(N) / 2))) return rc;
    if ((rc = make_tmap(h, &tbl, w_lo, N, K, gemm_block_n(N) / 2))) return rc;
    GemmLnArgs sp;
    sp.ta_lo = &tal;
    sp.tb_lo = &tbl;
    return launch_gemm(h, MCM_PROF_GEMM_OTHER, ta, tb, M, N, K, EPI_BIAS_RESID_F32, bias, out, resid, nullptr, 0, 0,
                       static_cast<cudaStream_t>(stream), sp);
}

int mcm_dbg_attention_split(McmHandle* h, const void* qkv_hi, const void* qkv_lo, void* o_hi, void* o_lo, int32_t b, int32_t S,
                            int32_t H, void* stream) {
    if (!h || !qkv_hi || !qkv_lo || !o_hi || !o_lo) return fail(h, MCM_EINVAL, "mcm_dbg_attention_split: NULL argument");
    if (b <= 0 || S <= 0 || H <= 0) return fail(h, MCM_EINVAL, "mcm_dbg_attention_split: b, S, H must be positive");
    DeviceGuard guard(h->cfg.device);
    return launch_attention_mma(h, static_cast<const op16_t*>(qkv_hi), static_cast<const op16_t*>(qkv_lo), static_cast<op16_t*>(o_hi),
                                static_cast<op16_t*>(o_lo), b, S, H, static_cast<cudaStream_t>(stream));
}

// ncclAllGather, resolved at run time from the NCCL the process already uses (no link-time dependency: the library
// must load on boxes and in processes that never touch NCCL)
int mcm_allgather_scores(McmHandle* h, void* nccl_comm, const float* local_dev, int32_t n_local_padded, float* all_dev, void* stream) {
    if (!h || !nccl_comm || !local_dev || !all_dev) return fail(h, MCM_EINVAL, "mcm_allgather_scores: NULL argument");
    if (n_local_padded <= 0) return fail(h, MCM_EINVAL, "mcm_allgather_scores: n_local_padded must be positive");
    typedef int (*AllGatherFn)(const void*, void*, size_t, int, void*, cudaStream_t);
    typedef const char* (*ErrStrFn)(int);
    static AllGatherFn fn = nullptr;
    static ErrStrFn errstr = nullptr;
    if (!fn) {
        void* sym = dlsym(RTLD_DEFAULT, "ncclAllGather");
        void* lib = nullptr;
        if (!sym) {
            lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // the copy already mapped into the process (torch's)
            if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW);
            if (lib) sym = dlsym(lib, "ncclAllGather");
        }
        if (!sym) return fail(h, MCM_EUNSUPPORTED, "mcm_allgather_scores: no NCCL in this process and libnccl.so.2 cannot be loaded");
        fn = reinterpret_cast<AllGatherFn>(sym);
        errstr = reinterpret_cast<ErrStrFn>(lib ? dlsym(lib, "ncclGetErrorString") : dlsym(RTLD_DEFAULT, "ncclGetErrorString"));
    }
    DeviceGuard guard(h->cfg.device);
    const int r = fn(local_dev, all_dev, static_cast<size_t>(n_local_padded), /*ncclFloat32*/ 7, nccl_comm, static_cast<cudaStream_t>(stream));
    if (r != 0) return fail(h, MCM_ECUDA, "ncclAllGather failed: %s", errstr ? errstr(r) : "unknown NCCL error");
    return MCM_OK;
}

int mcm_dbg_fold_ln(McmHandle* h, const float* w, const float* gamma, const float* beta, const float* bias, void* w16, float* c,
                    float* d, int32_t N, int32_t K, void* stream) {
    if (!h || !w || !gamma || !beta || !bias || !w16 || !c || !d) return fail(h, MCM_EINVAL, "mcm_dbg_fold_ln: NULL argument");
    if (N <= 0 || K <= 0) return fail(h, MCM_EINVAL, "mcm_dbg_fold_ln: N and K must be positive");
    fold_ln_weight_kernel<<<(N + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(w, gamma, beta, bias, static_cast<op16_t*>(w16), c,
                                                                                     d, N, K);
    MCM_CUDA(h, cudaGetLastError());
    return MCM_OK;
}

int mcm_dbg_gemm_ln(McmHandle* h, const void* a, const void* w, const float* d, const float* c, const float* stats, int32_t parts,
                    int32_t row_len, void* out, int32_t M, int32_t N, int32_t K, int32_t gelu, void* stream) {
    if (!h || !a || !w || !d || !c || !stats || !out) return fail(h, MCM_EINVAL, "mcm_dbg_gemm_ln: NULL argument");
    if (M <= 0 || parts <= 0 || row_len <= 0) return fail(h, MCM_EINVAL, "mcm_dbg_gemm_ln: M, parts and row_len must be positive");
    if (N % 128 != 0 || K % 64 != 0) return fail(h, MCM_EUNSUPPORTED, "mcm_dbg_gemm_ln: N %% 128 and K %% 64 must be 0");
    CUtensorMap ta, tb;
    int rc;
    if ((rc = make_tmap(h, &ta, a, M, K, kGemmBlockM))) return rc;
    if ((rc = make_tmap(h, &tb, w, N, K, gemm_block_n(N) / 2))) return rc;
    GemmLnArgs ln;
    ln.colsum = c;
    ln.stats_in = reinterpret_cast<const float2*>(stats);
    ln.stats_parts = parts;
    ln.stats_ld = M;
    ln.row_len = row_len;
    return launch_gemm(h, MCM_PROF_GEMM_OTHER, ta, tb, M, N, K, gelu ? EPI_LN_QGELU_F16 : EPI_LN_F16, d, out, nullptr, nullptr, 0, 0,
                       static_cast<cudaStream_t>(stream), ln);
}

int mcm_dbg_gemm_resid_ln(McmHandle* h, const void* a, const void* w, const float* bias, const float* resid, float* out, void* out16,
                          float* stats, int32_t M, int32_t N, int32_t K, int32_t* parts, void* stream) {
    if (!h || !a || !w || !bias || !resid || !out || !out16 || !stats || !parts)
        return fail(h, MCM_EINVAL, "mcm_dbg_gemm_resid_ln: NULL argument");
    if (M <= 0) return fail(h, MCM_EINVAL, "mcm_dbg_gemm_resid_ln: M must be positive");
    if (N % 128 != 0 || K % 64 != 0) return fail(h, MCM_EUNSUPPORTED, "mcm_dbg_gemm_resid_ln: N %% 128 and K %% 64 must be 0");
    CUtensorMap ta, tb;
    int rc;
    if ((rc = make_tmap(h, &ta, a, M, K, kGemmBlockM))) return rc;
    if ((rc = make_tmap(h, &tb, w, N, K, gemm_block_n(N) / 2))) return rc;
    *parts = 2 * (N / gemm_block_n(N));
    GemmLnArgs ln;
    ln.out16 = static_cast<op16_t*>(out16);
    ln.stats_out = reinterpret_cast<float2*>(stats);
    ln.stats_ld = M;
    return launch_gemm(h, MCM_PROF_GEMM_OTHER, ta, tb, M, N, K, EPI_BIAS_RESID_F32_LN, bias, out, resid, nullptr, 0, 0,
                       static_cast<cudaStream_t>(stream), ln);
}

int mcm_dbg_gemm_resid_h2(McmHandle* h, const void* a, const void* w, const float* bias, void* x_hi, void* x_lo, float* stats,
                          int32_t M, int32_t N, int32_t K, int32_t* parts, void* stream) {
    if (!h || !a || !w || !bias || !x_hi || !x_lo || !stats || !parts) return fail(h, MCM_EINVAL, "mcm_dbg_gemm_resid_h2: NULL argument");
    if (M <= 0) return fail(h, MCM_EINVAL, "mcm_dbg_gemm_resid_h2: M must be positive");
    if (N % 128 != 0 || K % 64 != 0) return fail(h, MCM_EUNSUPPORTED, "mcm_dbg_gemm_resid_h2: N %% 128 and K %% 64 must be 0");
    DeviceGuard guard(h->cfg.device);
    CUtensorMap ta, tb;
    int rc;
    if ((rc = make_tmap(h, &ta, a, M, K, kGemmBlockM))) return rc;
    if ((rc = make_tmap(h, &tb, w, N, K, gemm_block_n(N) / 2))) return rc;
    *parts = 2 * (N / gemm_block_n(N));
    GemmLnArgs ln;
    ln.resid16 = ln.out16 = static_cast<op16_t*>(x_hi);
    ln.resid16_lo = ln.out16_lo = static_cast<op16_t*>(x_lo);
    ln.stats_out = reinterpret_cast<float2*>(stats);
    ln.stats_ld = M;
    return launch_gemm(h, MCM_PROF_GEMM_OTHER, ta, tb, M, N, K, EPI_BIAS_RESID_H2_LN, bias, nullptr, nullptr, nullptr, 0, 0,
                       static_cast<cudaStream_t>(stream), ln);
}

int mcm_dbg_layernorm(McmHandle* h, const float* x, const float* g, const float* b, void* out, int32_t M, int32_t D,
                      float eps, int32_t out_f16, void* stream) {
    if (!h || !x || !g || !b || !out) return fail(h, MCM_EINVAL, "mcm_dbg_layernorm: NULL argument");
    return launch_layernorm(h, x, g, b, out, M, D, eps, out_f16 != 0, static_cast<cudaStream_t>(stream));
}

int mcm_dbg_attention(McmHandle* h, const void* qkv, void* o, int32_t b, int32_t S, int32_t H, void* stream) {
    if (!h || !qkv || !o) return fail(h, MCM_EINVAL, "mcm_dbg_attention: NULL argument");
    if (b <= 0 || S <= 0 || H <= 0) return fail(h, MCM_EINVAL, "mcm_dbg_attention: b, S, H must be positive");
    CUtensorMap tq, tkv, tx, to;
    if (S <= kAtcMaxS && !h->attn_mma) {
        int rc;
        if ((rc = make_tmap(h, &tq, qkv, static_cast<uint64_t>(b) * S, 3ull * H * 64, 128))) return rc;
        if ((rc = make_tmap_kv(h, &tkv, qkv, static_cast<uint64_t>(b), static_cast<uint64_t>(S), 3ull * H * 64, atc_kv_box_rows(S)))) return rc;
        if ((rc = make_tmap(h, &tx, qkv, static_cast<uint64_t>(b) * S, 3ull * H * 64, 8))) return rc;
        if ((rc = make_tmap_kv(h, &to, o, static_cast<uint64_t>(b), static_cast<uint64_t>(S), 64ull * H, 32))) return rc;
    }
    return launch_attention(h, tq, tkv, tx, to, static_cast<const op16_t*>(qkv), static_cast<op16_t*>(o), b, S, H,
                            static_cast<cudaStream_t>(stream));
}

int mcm_dbg_tail(McmHandle* h, const float* x, int32_t b, float T, int32_t kind, float* feats, float* scores, void* stream) {
    int rc = check_ready(h, b, scores != nullptr);
    if (rc) return rc;
    if (!x) return fail(h, MCM_EINVAL, "mcm_dbg_tail: NULL argument");
    return launch_tail(h, x, nullptr, nullptr, static_cast<size_t>(h->S) * h->D, b, T, kind, feats, scores, static_cast<cudaStream_t>(stream));
}

int mcm_dbg_embed(McmHandle* h, const float* images, int32_t b, float* x, void* stream) {
    int rc = check_ready(h, b, false);
    if (rc) return rc;
    if (!images || !x) return fail(h, MCM_EINVAL, "mcm_dbg_embed: NULL argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // run only pre_layrnorm here: the fused LN1 output goes to the workspace, x is copied out
    if ((rc = launch_embed(h, images, false, b, st, /*write_x=*/true))) return rc;
    MCM_CUDA(h, cudaMemcpyAsync(x, h->x, (size_t)b * h->S * h->D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return MCM_OK;
}

}  // extern "C"
