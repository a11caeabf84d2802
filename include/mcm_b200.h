/* mcm_b200 -- C ABI of the B200-native MCM scoring path.
 *
 * Drop-in boundary for ONE hot path of deeplearning-wisc/MCM:
 *
 *     CLIP ViT image-encoder forward -> L2-normalise -> cosine vs pre-encoded prompt bank
 *     -> softmax(./T) -> max            (reference: utils/detection_util.py:209-249,
 *                                        get_ood_scores_clip; the arithmetic it delegates to is
 *                                        transformers.models.clip.modeling_clip, "HF:" below)
 *
 * The reference is pure Python and has no FFI; its seam for this path is the Python function
 * `get_ood_scores_clip(args, net, loader, test_labels)` and the duck-typed `net`
 * (`.get_image_features(pixel_values=)`).  `mcm_b200/detection_util.py` and `mcm_b200/engine.py`
 * re-create those two seams on top of the entry points below through ctypes.
 *
 * Conventions
 *   - plain C types only: pointers, sizes, a raw `cudaStream_t` passed as `void*`.
 *   - every function returns 0 on success, a non-zero MCM_E* code on failure; the text of the last
 *     failure is available from mcm_last_error(handle) (or mcm_last_error(NULL) for mcm_create).
 *   - a handle is bound to one CUDA device and is NOT thread-safe (the reference drives one device
 *     from one Python thread, eval_ood_detection.py:57-58).  Every entry point selects the handle's
 *     device for its own duration and restores the caller's current device before it returns.
 *   - device work is enqueued on the caller's stream and is asynchronous; the caller synchronises
 *     when it reads the scores.  Nothing is allocated inside mcm_score / mcm_image_features.
 *   - one handle owns ONE activation workspace: consecutive forwards on the same handle are ordered
 *     against each other on the device (event wait), whatever streams they are enqueued on -- also
 *     the mcm_score_stream_host* calls, which run on the handle's own streams.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef MCM_B200_H_
#define MCM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCM_ABI_VERSION 2

enum {
    MCM_OK = 0,
    MCM_EINVAL = 1,     /* bad argument (shape, range, NULL)                    */
    MCM_ESTATE = 2,     /* call order: weights missing, bank not set, ...       */
    MCM_ECUDA = 3,      /* a CUDA runtime / driver call failed                  */
    MCM_ENOMEM = 4,
    MCM_EUNSUPPORTED = 5 /* shape outside what the sm_100a kernels are built for */
};

/* Reduction applied to the [b,K] cosine logits; mirrors `args.score`
 * (eval_ood_detection.py:36-37, utils/detection_util.py:233-248).  'maha' is a different method
 * and not part of this path. */
enum {
    MCM_SCORE_MCM = 0,       /* -max_k softmax(z/T)          :236,248 */
    MCM_SCORE_MAX_LOGIT = 1, /* -max_k z                     :234,248 */
    MCM_SCORE_ENERGY = 2,    /* -T * logsumexp(z/T)          :239     */
    MCM_SCORE_ENTROPY = 3,   /* entropy(softmax(z/T))        :243     */
    MCM_SCORE_VAR = 4        /* -var_k softmax(z/T)          :246     */
};

/* Shape of the CLIP vision tower (HF CLIPVisionConfig + projection_dim,
 * HF:configuration_clip.py; the reference picks it with --CLIP_ckpt, utils/train_eval_util.py:19-21). */
typedef struct McmConfig {
    int32_t image_size;  /* 224                                                     */
    int32_t patch;       /* 16 (B/16), 14 (L/14), 32 (B/32)                         */
    int32_t width;       /* hidden_size D; multiple of 128, <= 1024                 */
    int32_t layers;      /* num_hidden_layers L                                     */
    int32_t heads;       /* num_attention_heads H; D / H must be 64                 */
    int32_t mlp;         /* intermediate_size F; multiple of 128                    */
    int32_t proj;        /* projection_dim P; multiple of 16                        */
    float eps;           /* layer_norm_eps (1e-5)                                   */
    int32_t max_batch;   /* largest `b` a single mcm_score call will be given       */
    int32_t device;      /* CUDA ordinal                                            */
} McmConfig;

typedef struct McmHandle McmHandle;

/* Allocate packed-weight storage, the activation workspace for `max_batch` images and the TMA
 * descriptors on `cfg->device`.  Replaces the model half of set_model_clip
 * (utils/train_eval_util.py:15-26: CLIPModel.from_pretrained(...).cuda()). */
int mcm_create(const McmConfig* cfg, McmHandle** out);
void mcm_destroy(McmHandle* h);
const char* mcm_last_error(const McmHandle* h);

/* Hand one fp32 tensor of the HuggingFace CLIPModel state_dict to the engine, by its HF key
 * (SURVEY.md 8b "Weights in"; e.g. "vision_model.encoder.layers.3.mlp.fc1.weight").  `data` may be
 * a host or a device pointer (cudaMemcpyDefault); `numel` must match the configured shape.  Keys
 * that are not part of the vision path (text tower, logit_scale, position_ids) are ignored and
 * reported through *used = 0.  Replaces CLIPModel.load_state_dict for this path. */
int mcm_load_weight(McmHandle* h, const char* hf_key, const float* data, int64_t numel, int32_t* used);

/* Verify that every tensor of the vision path arrived and convert/pack them for the kernels
 * (fp16 GEMM operands, fused QKV weight, padded patch filter, layer_norm1 / layer_norm2 folded into
 * the q/k/v and fc1 weights).  Must precede any compute call; call it again after re-loading tensors. */
int mcm_finalize_weights(McmHandle* h);

/* Install the pre-encoded prompt bank: `bank` is [K,P] fp32 (host or device), one row per entry
 * of `test_labels`.  Rows are L2-normalised on the device exactly like
 * utils/detection_util.py:231 unless `already_unit` is non-zero.  The bank is copied. */
int mcm_set_text_bank(McmHandle* h, const float* bank, int32_t K, int32_t already_unit);

/* net.get_image_features(pixel_values=images)  (HF:829-863; called at utils/detection_util.py:225).
 * images_dev: [b,3,image_size,image_size] fp32 NCHW on the device, already CLIP-normalised.
 * feats_dev : [b,P] fp32, the un-normalised projected features. */
int mcm_image_features(McmHandle* h, const float* images_dev, int32_t b, float* feats_dev, void* stream);

/* One batch of utils/detection_util.py:225-248 with the bank pre-encoded:
 * scores_dev[i] = reduce_k(normalise(features_i) . bank_k), `score_kind` one of MCM_SCORE_*.
 * T follows args.T (eval_ood_detection.py:31; an int there, any positive float here). */
int mcm_score(McmHandle* h, const float* images_dev, int32_t b, float T, int32_t score_kind, float* scores_dev,
              void* stream);

/* Whole stream of the loop at utils/detection_util.py:220-249 from HOST memory: n images are cut
 * into batches of `batch` (<= max_batch), copied host->device on a copy stream while the previous
 * batch is scored, and the n scores are copied back.  Synchronous.  images_host should be pinned
 * for full copy bandwidth but need not be. */
int mcm_score_stream_host(McmHandle* h, const float* images_host, int64_t n, int32_t batch, float T,
                          int32_t score_kind, float* scores_host);

/* uint8 ingest.  Same three calls, but the images are what the reference's preprocess holds
 * BEFORE its last two steps (utils/train_eval_util.py:29-34: Resize(224) -> CenterCrop(224) -> ToTensor -> Normalize):
 * uint8 [b, image_size, image_size, 3] HWC, the layout of a decoded PIL image / JPEG decoder output.  ToTensor
 * (x / 255) and Normalize ((x - mean) / std, fp32, torchvision's operation order) are fused into the patch gather,
 * so the fp16 patch rows -- and therefore features and scores -- are bit-identical to mcm_score on the fp32 tensor
 * the reference's DataLoader would have produced, at a quarter of the PCIe and HBM bytes.
 * mcm_set_normalization replaces the default CLIP constants (utils/train_eval_util.py:27-28). */
int mcm_set_normalization(McmHandle* h, const float* mean3, const float* std3);
int mcm_image_features_u8(McmHandle* h, const uint8_t* images_dev, int32_t b, float* feats_dev, void* stream);
int mcm_score_u8(McmHandle* h, const uint8_t* images_dev, int32_t b, float T, int32_t score_kind, float* scores_dev,
                 void* stream);
int mcm_score_stream_host_u8(McmHandle* h, const uint8_t* images_host, int64_t n, int32_t batch, float T,
                             int32_t score_kind, float* scores_host);

/* Resize(image_size) + CenterCrop(image_size), the FIRST two steps of the reference preprocess (utils/train_eval_util.py:
 * 29-31; run there by torchvision on PIL images in the DataLoader workers), on the device for a batch of decoded RGB
 * images of different sizes: image i is uint8 [hs[i], ws[i], 3] (HWC) at byte offsets[i] of the packed device buffer
 * `src_dev`; offsets / hs / ws are HOST arrays of n entries.  dst_dev is uint8 [n, image_size, image_size, 3], ready
 * for mcm_score_u8 / mcm_image_features_u8, and bit-identical to torchvision.transforms.Resize + CenterCrop on PIL
 * images (Pillow's antialiased two-pass fixed-point bilinear resampler, torchvision's size / crop-offset rules).
 * Asynchronous on `stream`; the per-call resampling tables are built on the host and copied with the launch. */
int mcm_resize_crop_u8(McmHandle* h, const uint8_t* src_dev, const int64_t* offsets_host, const int32_t* hs_host,
                       const int32_t* ws_host, int32_t n, uint8_t* dst_dev, void* stream);

/* The whole loop of utils/detection_util.py:220-249 INCLUDING the preprocess, from decoded images in HOST memory: image i is
 * uint8 [hs[i], ws[i], 3] at byte offsets[i] of `packed_host` (pinned for full copy bandwidth).  Batches of `batch` images
 * are copied host->device on a copy stream while the previous batch is resized, cropped, normalised and scored.  Synchronous. */
int mcm_score_stream_host_images(McmHandle* h, const uint8_t* packed_host, const int64_t* offsets, const int32_t* hs,
                                 const int32_t* ws, int64_t n, int32_t batch, float T, int32_t score_kind, float* scores_host);

/* The host half of mcm_resize_crop_u8 for ONE h x w image (no device needed): the fixed-point resampling tables of the
 * `size` output columns / rows that survive the crop, `[first source index, count, k[ksize]]` per output, ksize2 =
 * {ksize_h, ksize_v}; table_h / table_v hold `cap` int32 each.  Lets the CPU tests check the planner against Pillow. */
int mcm_dbg_resize_tables(int32_t h, int32_t w, int32_t size, int32_t* ksize2, int32_t* table_h, int32_t* table_v,
                          int32_t cap);

/* Mahalanobis baseline, the reference's `--score maha` (utils/detection_util.py:182-207; eval_ood_detection.py:72-79):
 *   score_i = -max_k( -0.5 (f_i - mu_k)^T P (f_i - mu_k) )   over the K class means mu_k and the shared precision P.
 * mcm_set_maha takes the statistics pre-whitened by the host (mcm_b200/engine.py does it in fp64): with P = L L^T
 * (Cholesky), `lt` = L^T [P,P] row-major and `centres` = mu_k L [K,P]; `normalize` mirrors args.normalize (features
 * L2-normalised first, :197-198).  mcm_maha_score replaces the whole batch loop body (:193-204): the reference's
 * K-iteration Python loop of two GEMMs per class becomes one whitening GEMM + K squared distances per image.
 * mcm_dbg_maha_from_features runs the same tail on given [b,P] features (tests). */
int mcm_set_maha(McmHandle* h, const float* lt, const float* centres, int32_t K, int32_t normalize);
int mcm_maha_score(McmHandle* h, const float* images_dev, int32_t b, float* scores_dev, void* stream);
int mcm_dbg_maha_from_features(McmHandle* h, const float* feats_dev, int32_t b, float* scores_dev, void* stream);

/* Number of kernels of this library launched on the handle's device since the last reset
 * (bench.py reports it as `gpu_launches`). */
int64_t mcm_launch_count(const McmHandle* h);
void mcm_reset_launch_count(McmHandle* h);

/* Engine options.  MCM_OPT_CLS_SHORTCUT (default 1): in the LAST encoder layer run attention,
 * out_proj, layer_norm2 and the MLP only for the CLS query row of every image -- the only row the
 * reference consumes afterwards (pooled = last_hidden_state[:, 0], HF:685).  Results are
 * identical; 0 runs the full last layer (the number bench.py reports beside the default). */
enum {
    MCM_OPT_CLS_SHORTCUT = 1,
    /* MCM_OPT_PRECISION: arithmetic of the tensor-core operands.
     *   MCM_PRECISION_FP16 (0, default): one fp16 value per operand element (fp32 accumulation, residual stream,
     *     LayerNorm statistics, softmax and tail).  Scores within ~1e-7 of the fp32 reference, AUROC within 0.01 pt;
     *     FPR95 -- a COUNT of images above a threshold -- can move by a few images per 10 000.
     *   MCM_PRECISION_SPLIT (1): every operand element is an fp16 (hi, lo) pair (~22 significant bits) and every
     *     product the three-term sum A_hi W_hi + A_lo W_hi + A_hi W_lo: fp32-class results at 3x the tensor work.
     *     This is the mode in which AUROC / FPR95 agree with the reference (fp32 end to end,
     *     utils/detection_util.py:225-236) to the 0.05 pt of the parity bar on every stream.
     *   The first switch to MCM_PRECISION_SPLIT allocates the low-half activation buffers (synchronises). */
    MCM_OPT_PRECISION = 2,
    /* MCM_OPT_CUDA_GRAPH (default 0): mcm_score / mcm_image_features (and their _u8 / stream_host forms) replay a
     * CUDA graph captured per (entry point, batch size, option set) instead of ~70 individual launches:
     * for small batches, where the forward is launch-bound. */
    MCM_OPT_CUDA_GRAPH = 3
};
enum { MCM_PRECISION_FP16 = 0, MCM_PRECISION_SPLIT = 1 };
int mcm_set_option(McmHandle* h, int32_t option, int32_t value);

/* Collation of the per-rank scores of a sharded stream: every rank contributes `n_local_padded`
 * fp32 scores (the tail rank pads, mirroring the reference's own `[:len(loader.dataset)]` trim,
 * utils/detection_util.py:249), all ranks receive the concatenation in rank order.  `nccl_comm` is the caller's
 * `ncclComm_t` (passed as void*; the library resolves ncclAllGather from the NCCL already loaded in the process --
 * torch's -- or from libnccl.so.2, and fails with MCM_EUNSUPPORTED if there is none).  Asynchronous on `stream`. */
int mcm_allgather_scores(McmHandle* h, void* nccl_comm, const float* local_dev, int32_t n_local_padded, float* all_dev,
                         void* stream);

/* Optional per-launch timing for bench.py's roofline: while enabled, every kernel launch of a
 * forward is bracketed by CUDA events on the launching stream.  mcm_profile_read synchronises the
 * device and returns accumulated milliseconds and launch counts per kernel kind (arrays of
 * MCM_PROF_KINDS entries); `reset` clears the accumulators. */
enum {
    MCM_PROF_PATCHIFY = 0, MCM_PROF_GEMM_PATCH = 1, MCM_PROF_EMBED_FINISH = 2, MCM_PROF_GEMM_QKV = 3,
    MCM_PROF_ATTENTION = 4, MCM_PROF_GEMM_OUT = 5, MCM_PROF_LAYERNORM = 6, MCM_PROF_GEMM_FC1 = 7,
    MCM_PROF_GEMM_FC2 = 8, MCM_PROF_TAIL = 9, MCM_PROF_GEMM_OTHER = 10, MCM_PROF_KINDS = 11
};
int mcm_profile_enable(McmHandle* h, int32_t on);
int mcm_profile_read(McmHandle* h, double* ms, int64_t* counts, int32_t reset);

/* Algorithmic FLOPs per image of the configured tower with a K-row bank. */
double mcm_flops_per_image(const McmConfig* cfg, int32_t K);

int32_t mcm_abi_version(void);

/* ---- per-kernel entry points (used by tests/ to check every kernel against a torch fp32
 *      reference of the same op, and by bench.py for the per-kernel roofline).  All pointers are
 *      device pointers; all calls are asynchronous on `stream`. ---- */

/* out = epilogue(A[M,K] @ W[N,K]^T): A, W fp16 row-major.
 * epi 0: out fp16 = acc + bias            1: out fp16 = quick_gelu(acc + bias)
 * epi 2: out f32  = resid + acc + bias    (resid may alias out) */
int mcm_dbg_gemm(McmHandle* h, const void* a_f16, const void* w_f16, const float* bias, const float* resid,
                 void* out, int32_t M, int32_t N, int32_t K, int32_t epi, void* stream);
/* The split-precision GEMM on given operand pairs (a = a_hi + a_lo, w = w_hi + w_lo, all fp16 row-major):
 * out f32 = resid + (a_hi w_hi^T + a_lo w_hi^T + a_hi w_lo^T) + bias  (resid may alias out). */
int mcm_dbg_gemm_split(McmHandle* h, const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo,
                       const float* bias, const float* resid, float* out, int32_t M, int32_t N, int32_t K, void* stream);
/* The LayerNorm-folded projections the forward actually runs (csrc/gemm_tcgen05.cuh): layer_norm1 / layer_norm2
 * (HF:371,380) never materialise; the projection reads the RAW fp16 residual rows and its epilogue applies the row
 * statistics.
 *   mcm_dbg_fold_ln      : W f32 [N,K], gamma/beta [K], bias [N] -> w16 = fp16(gamma o W), c[N] = row sums of w16,
 *                          d[N] = bias + beta @ W^T
 *   mcm_dbg_gemm_resid_ln: out f32 = resid + A W^T + bias (resid may alias out), out16 = fp16(out), and
 *                          stats = float2 [*parts][M] partial (sum, sum of squares) over column slices of out
 *                          (allocate N / 64 parts; *parts returns how many were written)
 *   mcm_dbg_gemm_ln      : out fp16 = [quick_gelu](rstd * (A w16^T) - rstd * mu * c + d) with mu / rstd of every
 *                          row taken from `stats` (float2 [parts][M]) over `row_len` elements, eps from the config */
int mcm_dbg_fold_ln(McmHandle* h, const float* w, const float* gamma, const float* beta, const float* bias, void* w16,
                    float* c, float* d, int32_t N, int32_t K, void* stream);
int mcm_dbg_gemm_ln(McmHandle* h, const void* a_f16, const void* w16, const float* d, const float* c, const float* stats,
                    int32_t parts, int32_t row_len, void* out_f16, int32_t M, int32_t N, int32_t K, int32_t gelu,
                    void* stream);
int mcm_dbg_gemm_resid_ln(McmHandle* h, const void* a_f16, const void* w_f16, const float* bias, const float* resid,
                          float* out, void* out16, float* stats, int32_t M, int32_t N, int32_t K, int32_t* parts,
                          void* stream);
/* The residual epilogue the forward runs: the residual stream is an fp16 (hi, lo) pair, updated IN PLACE:
 * (x_hi, x_lo) <- split(x_hi + x_lo + A W^T + bias), stats = float2 [*parts][M] partial (sum, sum of squares).
 * K <= 1024 takes the TMA form of the epilogue, longer K the LSU form. */
int mcm_dbg_gemm_resid_h2(McmHandle* h, const void* a_f16, const void* w_f16, const float* bias, void* x_hi, void* x_lo,
                          float* stats, int32_t M, int32_t N, int32_t K, int32_t* parts, void* stream);
/* nn.LayerNorm over the last dim (HF:359-361): x f32 [M,D] -> out fp16 (out_f16 != 0) or f32. */
int mcm_dbg_layernorm(McmHandle* h, const float* x, const float* gamma, const float* beta, void* out, int32_t M,
                      int32_t D, float eps, int32_t out_f16, void* stream);
/* softmax(q k^T / 8) v per (image, head) (HF:261-279,318-331): qkv fp16 [b*S, 3*H*64] -> o fp16 [b*S, H*64]. */
int mcm_dbg_attention(McmHandle* h, const void* qkv_f16, void* o_f16, int32_t b, int32_t S, int32_t H,
                      void* stream);
/* The same in the split-precision mode: qkv = qkv_hi + qkv_lo, o = o_hi + o_lo. */
int mcm_dbg_attention_split(McmHandle* h, const void* qkv_hi, const void* qkv_lo, void* o_hi, void* o_lo, int32_t b,
                            int32_t S, int32_t H, void* stream);
/* CLS pool + post_layernorm + visual_projection (HF:685-686,860-861) + the scoring tail.
 * x f32 [b*S, D] (row b*S is the CLS token); feats (may be NULL) [b,P]; scores (may be NULL) [b]. */
int mcm_dbg_tail(McmHandle* h, const float* x, int32_t b, float T, int32_t score_kind, float* feats, float* scores,
                 void* stream);
/* embeddings + pre_layrnorm (HF:202-218,677): images f32 [b,3,H,W] -> x f32 [b*S, D]. */
int mcm_dbg_embed(McmHandle* h, const float* images, int32_t b, float* x, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MCM_B200_H_ */
