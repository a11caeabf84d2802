"""CPU oracle for the MCM hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  The product package
(``mcm_b200``) never imports it and has no CPU fallback.

What it is
----------
A plain torch-CPU restatement (fp32 by default, fp64 on request) of the
arithmetic of the reference hot path:

  * ``utils/detection_util.py:209-249``  get_ood_scores_clip  (the scoring loop)
  * the HuggingFace ``transformers`` CLIP vision tower it delegates to
    (``transformers.models.clip.modeling_clip``, installed version 5.5.0 --
    an un-vendored, un-pinned dependency of the reference, ``README.md:23``;
    call sites ``utils/train_eval_util.py:9,23`` and
    ``utils/detection_util.py:225,229``).  Line numbers prefixed ``HF:`` below
    are into that file.
  * the metric layer ``utils/detection_util.py:47-119``.

Parity pinning
--------------
The reference ships no tests / golden vectors, so the pin
is the reference function itself, executed in the authoring container through
``oracle/reference_shims.py`` (API-drift shims only; no arithmetic changed) on
seeded inputs.  ``oracle/make_golden.py`` does that, checks this restatement
against it, and writes the fixtures under ``tests/golden/``.  Status:
"pinned against the reference run here; unpinned by reference-owned tests"
(there are none).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch


@dataclass(frozen=True)
class VisionCfg:
    """Shape of one CLIP vision tower (HF: configuration_clip.py CLIPVisionConfig)."""
    image_size: int = 224
    patch: int = 16
    width: int = 768        # hidden_size D
    layers: int = 12
    heads: int = 12
    mlp: int = 3072         # intermediate_size F
    proj: int = 512         # projection_dim P
    eps: float = 1e-5

    @property
    def grid(self) -> int:
        return self.image_size // self.patch

    @property
    def seq(self) -> int:
        return self.grid * self.grid + 1


CFGS = {
    "ViT-B/16": VisionCfg(224, 16, 768, 12, 12, 3072, 512),
    "ViT-B/32": VisionCfg(224, 32, 768, 12, 12, 3072, 512),
    "ViT-L/14": VisionCfg(224, 14, 1024, 24, 16, 4096, 768),
    # small towers for fast CPU cases (dh stays 64 like every CLIP tower)
    "tiny": VisionCfg(224, 32, 128, 2, 2, 256, 64),
    "small": VisionCfg(224, 16, 256, 3, 4, 512, 128),
}


# the Normalize constants of the reference preprocess, utils/train_eval_util.py:27-28
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def preprocess_u8(images_u8, mean=CLIP_MEAN, std=CLIP_STD):
    """Last two steps of the reference ``val_preprocess`` (``utils/train_eval_util.py:29-34``) on decoded pixels:
    ``ToTensor`` (torchvision ``functional.to_tensor``: HWC uint8 -> CHW, ``.to(float32).div(255)``) then
    ``Normalize`` (``functional.normalize``: ``tensor.sub_(mean).div_(std)`` with fp32 ``mean`` / ``std`` tensors).
    ``images_u8``: uint8 ``[n, H, W, 3]`` -> fp32 ``[n, 3, H, W]``.  Pinned against torchvision itself on PIL images
    in ``tests/test_oracle_golden.py``."""
    x = torch.as_tensor(images_u8)
    assert x.dtype == torch.uint8 and x.dim() == 4 and x.shape[3] == 3
    x = x.permute(0, 3, 1, 2).contiguous().to(torch.float32).div(255)
    m = torch.as_tensor(mean, dtype=torch.float32).view(1, 3, 1, 1)
    sd = torch.as_tensor(std, dtype=torch.float32).view(1, 3, 1, 1)
    return x.sub_(m).div_(sd)


def _ln(x, w, b, eps):
    """nn.LayerNorm: biased variance, eps inside the sqrt, affine (HF:359-361,659-661)."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def _lin(x, w, b=None):
    y = x @ w.t()
    return y if b is None else y + b


def vision_embeddings(pixels, sd, cfg: VisionCfg):
    """CLIPVisionEmbeddings.forward, HF:202-218: patch conv (no bias, stride =
    kernel = patch) -> flatten/transpose -> cat CLS -> + position embedding."""
    B = pixels.shape[0]
    if pixels.shape[2] != cfg.image_size or pixels.shape[3] != cfg.image_size:
        raise ValueError(  # HF:204-207
            f"Input image size ({pixels.shape[2]}*{pixels.shape[3]}) doesn't match model "
            f"({cfg.image_size}*{cfg.image_size}).")
    w = sd["vision_model.embeddings.patch_embedding.weight"]          # [D,3,p,p]
    g, p = cfg.grid, cfg.patch
    # non-overlapping conv == per-patch dot product with the flattened filter
    patches = pixels.reshape(B, 3, g, p, g, p).permute(0, 2, 4, 1, 3, 5).reshape(B, g * g, 3 * p * p)
    pe = patches @ w.reshape(cfg.width, -1).t()                        # [B,Np,D]
    cls = sd["vision_model.embeddings.class_embedding"].expand(B, 1, -1)
    x = torch.cat([cls, pe], dim=1)
    return x + sd["vision_model.embeddings.position_embedding.weight"].unsqueeze(0)


def encoder_layer(x, sd, i, cfg: VisionCfg):
    """CLIPEncoderLayer.forward HF:363-384 with CLIPAttention.forward HF:300-336
    (eager attention HF:261-279 == the math SDPA computes; no mask, not causal,
    dropout 0) and CLIPMLP.forward HF:347-351 (quick_gelu, activations.py:117-123)."""
    pre = f"vision_model.encoder.layers.{i}."
    B, S, D = x.shape
    H = cfg.heads
    dh = D // H
    h = _ln(x, sd[pre + "layer_norm1.weight"], sd[pre + "layer_norm1.bias"], cfg.eps)
    q = _lin(h, sd[pre + "self_attn.q_proj.weight"], sd[pre + "self_attn.q_proj.bias"])
    k = _lin(h, sd[pre + "self_attn.k_proj.weight"], sd[pre + "self_attn.k_proj.bias"])
    v = _lin(h, sd[pre + "self_attn.v_proj.weight"], sd[pre + "self_attn.v_proj.bias"])
    q = q.view(B, S, H, dh).transpose(1, 2)
    k = k.view(B, S, H, dh).transpose(1, 2)
    v = v.view(B, S, H, dh).transpose(1, 2)
    att = torch.softmax((q @ k.transpose(-1, -2)) * (dh ** -0.5), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, S, D)
    x = x + _lin(o, sd[pre + "self_attn.out_proj.weight"], sd[pre + "self_attn.out_proj.bias"])
    h = _ln(x, sd[pre + "layer_norm2.weight"], sd[pre + "layer_norm2.bias"], cfg.eps)
    h = _lin(h, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"])
    h = h * torch.sigmoid(1.702 * h)
    return x + _lin(h, sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])


def image_features(pixels, sd, cfg: VisionCfg, return_hidden=False):
    """CLIPModel.get_image_features HF:829-863 -> CLIPVisionTransformer.forward
    HF:667-691: embeddings, pre_layrnorm (sic), L layers, CLS pool, post_layernorm,
    visual_projection (no bias).  Returns the un-normalised [B,P] features."""
    x = vision_embeddings(pixels, sd, cfg)
    x = _ln(x, sd["vision_model.pre_layrnorm.weight"], sd["vision_model.pre_layrnorm.bias"], cfg.eps)
    for i in range(cfg.layers):
        x = encoder_layer(x, sd, i, cfg)
    pooled = _ln(x[:, 0, :], sd["vision_model.post_layernorm.weight"],
                 sd["vision_model.post_layernorm.bias"], cfg.eps)
    f = _lin(pooled, sd["visual_projection.weight"])
    return (f, x) if return_hidden else f


def scores_from_features(feats, bank, T=1, score="MCM"):
    """utils/detection_util.py:226,232-248 given un-normalised image features
    [B,P] and the (already unit-row) text bank [K,P].  Returns float32 numpy [B]."""
    f = feats.float() if feats.dtype != torch.float64 else feats
    f = f / f.norm(dim=-1, keepdim=True)
    out = f @ bank.t()
    if score == "max-logit":
        smax = out.cpu().numpy()
    else:
        smax = torch.softmax(out / T, dim=1).cpu().numpy()
    if score == "energy":
        return -(T * torch.logsumexp(out / T, dim=1)).cpu().numpy()
    if score == "entropy":
        from scipy.stats import entropy
        return entropy(smax, axis=1)
    if score == "var":
        return -np.var(smax, axis=1)
    if score in ("MCM", "max-logit"):
        return -np.max(smax, axis=1)
    raise ValueError(score)


@torch.no_grad()
def ood_scores(images, sd, cfg: VisionCfg, bank, T=1, score="MCM", batch=64, dtype=torch.float32):
    """The whole loop of utils/detection_util.py:209-249 with the bank pre-encoded
    (it is batch-invariant under eval()+no_grad).
    ``images`` is a float32 tensor/ndarray [N,3,H,W]; returns float32 numpy [N]."""
    images = torch.as_tensor(images)
    sdc = {k: v.to(dtype) for k, v in sd.items() if k.startswith("vision_model.") or k == "visual_projection.weight"}
    bank = torch.as_tensor(bank).to(dtype)
    out = []
    for s in range(0, images.shape[0], batch):
        f = image_features(images[s:s + batch].to(dtype), sdc, cfg)
        out.append(np.asarray(scores_from_features(f, bank, T, score), dtype=np.float32))
    if not out:
        return np.zeros((0,), dtype=np.float32)
    return np.concatenate(out, axis=0)[: images.shape[0]].copy()


@torch.no_grad()
def ood_scores_as_shipped(images, sd, cfg: VisionCfg, text_model, tokenizer, test_labels, T=1, score="MCM", batch=256):
    """The loop of utils/detection_util.py:209-249 AS SHIPPED: the prompts are tokenised and the whole text tower is run
    again for EVERY image batch (:228-231), the [b, K] softmax is reduced on the host (:236,248).  The vision side is this
    module's restatement; the text side is the HuggingFace text tower the reference itself calls (``text_model`` needs
    ``get_text_features(input_ids=, attention_mask=)``).  Same values as :func:`ood_scores` with the bank pre-encoded --
    eval() + no_grad make the text tower deterministic -- at the reference's own cost; used by bench.py's cpu_baseline."""
    images = torch.as_tensor(images)
    sdc = {k: v.float() for k, v in sd.items() if k.startswith("vision_model.") or k == "visual_projection.weight"}
    out = []
    for s in range(0, images.shape[0], batch):
        f = image_features(images[s:s + batch].float(), sdc, cfg)
        tok = tokenizer([f"a photo of a {c}" for c in test_labels], padding=True, return_tensors="pt")     # :228
        tf = text_model.get_text_features(input_ids=tok["input_ids"], attention_mask=tok["attention_mask"])
        tf = (tf.pooler_output if hasattr(tf, "pooler_output") else tf).float()                             # :229-230
        tf = tf / tf.norm(dim=-1, keepdim=True)                                                             # :231
        out.append(np.asarray(scores_from_features(f, tf, T, score), dtype=np.float32))
    if not out:
        return np.zeros((0,), dtype=np.float32)
    return np.concatenate(out, axis=0)[: images.shape[0]].copy()


# ----------------------------------------------------------------------------
# metric layer, utils/detection_util.py:47-119
# ----------------------------------------------------------------------------
def stable_cumsum(arr, rtol=1e-05, atol=1e-08):
    """utils/detection_util.py:47-63."""
    out = np.cumsum(arr, dtype=np.float64)
    expected = np.sum(arr, dtype=np.float64)
    if not np.allclose(out[-1], expected, rtol=rtol, atol=atol):
        raise RuntimeError("cumsum was found to be unstable")
    return out


def fpr_at_recall(y_true, y_score, recall_level=0.95):
    """utils/detection_util.py:66-106 (FPR at the threshold whose recall is
    closest to ``recall_level``; positives = label 1)."""
    y_true = np.asarray(y_true) == 1
    y_score = np.asarray(y_score)
    order = np.argsort(y_score, kind="mergesort")[::-1]
    y_score = y_score[order]
    y_true = y_true[order]
    distinct = np.where(np.diff(y_score))[0]
    thr = np.r_[distinct, y_true.size - 1]
    tps = stable_cumsum(y_true)[thr]
    fps = 1 + thr - tps
    recall = tps / tps[-1]
    last = tps.searchsorted(tps[-1])
    sl = slice(last, None, -1)
    recall, fps = np.r_[recall[sl], 1], np.r_[fps[sl], 0]
    cutoff = np.argmin(np.abs(recall - recall_level))
    return fps[cutoff] / np.sum(np.logical_not(y_true))


def auroc(labels, scores):
    """Area under ROC by the rank statistic with mid-ranks for ties
    (what sklearn.metrics.roc_auc_score, utils/detection_util.py:115, evaluates)."""
    labels = np.asarray(labels).astype(bool)
    scores = np.asarray(scores, dtype=np.float64)
    order = np.argsort(scores, kind="mergesort")
    s = scores[order]
    ranks = np.empty(len(s), dtype=np.float64)
    i = 0
    n = len(s)
    # mid-ranks
    bounds = np.r_[0, np.where(np.diff(s))[0] + 1, n]
    for a, b in zip(bounds[:-1], bounds[1:]):
        ranks[a:b] = 0.5 * (a + b - 1) + 1.0
    r = np.empty(n, dtype=np.float64)
    r[order] = ranks
    n_pos = labels.sum()
    n_neg = n - n_pos
    return (r[labels].sum() - n_pos * (n_pos + 1) / 2.0) / (n_pos * n_neg)


def average_precision(labels, scores):
    """sklearn.metrics.average_precision_score (utils/detection_util.py:116):
    sum over distinct thresholds of (R_n - R_{n-1}) * P_n."""
    labels = np.asarray(labels).astype(bool)
    scores = np.asarray(scores)
    order = np.argsort(scores, kind="mergesort")[::-1]
    s = scores[order]
    y = labels[order]
    distinct = np.where(np.diff(s))[0]
    thr = np.r_[distinct, y.size - 1]
    tps = np.cumsum(y, dtype=np.float64)[thr]
    fps = 1 + thr - tps
    precision = tps / (tps + fps)
    recall = tps / tps[-1]
    return float(np.sum(np.diff(np.r_[0.0, recall]) * precision))


def get_measures(_pos, _neg, recall_level=0.95):
    """utils/detection_util.py:108-119; returns (auroc, aupr, fpr)."""
    pos = np.array(_pos[:]).reshape((-1, 1))
    neg = np.array(_neg[:]).reshape((-1, 1))
    examples = np.squeeze(np.vstack((pos, neg)))
    labels = np.zeros(len(examples), dtype=np.int32)
    labels[: len(pos)] += 1
    return auroc(labels, examples), average_precision(labels, examples), fpr_at_recall(labels, examples, recall_level)


def flops_per_image(cfg: VisionCfg, K: int) -> float:
    """Algorithmic FLOPs per image."""
    S, D, L, P, p = cfg.seq, cfg.width, cfg.layers, cfg.proj, cfg.patch
    F = cfg.mlp
    return (2.0 * (S - 1) * (3 * p * p) * D + L * (8.0 * S * D * D + 4.0 * S * D * F + 4.0 * S * S * D)
            + 2.0 * D * P + 2.0 * P * K)


# ----------------------------------------------------------------------------- Mahalanobis baseline ---
# The reference's `--score maha` path (utils/detection_util.py:148-207; eval_ood_detection.py:72-79,87-88):
# class means + one shared precision matrix estimated from ID training features, then per image
#   score = -max_k ( -0.5 * (f - mu_k)^T P (f - mu_k) )  =  min_k 0.5 * (f - mu_k)^T P (f - mu_k).

def maha_mean_prec(features, batch_labels, n_cls, normalize=False):
    """``get_mean_prec`` (``utils/detection_util.py:148-180``) given the per-batch feature tensors and label tensors
    the reference loop sees.  Reproduces it literally, including that ``classwise_idx`` collects the BATCH index
    ``idx`` once per sample of the batch (``:166-167``) and that those batch indices then index the ROWS of the
    concatenated feature matrix (``:171-172``)."""
    from collections import defaultdict
    classwise_idx = defaultdict(list)
    all_features = []
    for idx, (f, labels) in enumerate(zip(features, batch_labels)):
        f = f.float().clone()
        if normalize:
            f /= f.norm(dim=-1, keepdim=True)
        for label in labels:
            classwise_idx[int(label)].append(idx)
        all_features.append(f)
    all_features = torch.cat(all_features)
    classwise_mean = torch.empty(n_cls, all_features.shape[1])
    for cls in range(n_cls):
        classwise_mean[cls] = torch.mean(all_features[classwise_idx[cls]].float(), dim=0)
        if normalize:
            classwise_mean[cls] /= classwise_mean[cls].norm(dim=-1, keepdim=True)
    cov = torch.cov(all_features.T.double())
    precision = torch.linalg.inv(cov).float()
    return classwise_mean, precision


def maha_scores_from_features(features, classwise_mean, precision, normalize=False):
    """One batch of ``get_Mahalanobis_score`` (``:195-205``): features ``[b, P]`` -> ``[b]`` float32."""
    f = torch.as_tensor(features).float().clone()
    if normalize:
        f /= f.norm(dim=-1, keepdim=True)
    cols = []
    for i in range(classwise_mean.shape[0]):
        zero_f = f - classwise_mean[i]
        cols.append((-0.5 * torch.mm(torch.mm(zero_f, precision), zero_f.t()).diag()).view(-1, 1))
    score, _ = torch.max(torch.cat(cols, 1), dim=1)
    return (-score).numpy().astype(np.float32)


def maha_scores(images, sd, cfg: VisionCfg, classwise_mean, precision, batch=64, normalize=False, in_dist=True):
    """``get_Mahalanobis_score`` (``:182-207``) over a stream, including its batch rule: for ``in_dist=False`` the
    loop stops at batch ``len(dataset) // batch_size`` (``:191-192``), i.e. a trailing partial batch of an OOD set is
    dropped."""
    images = torch.as_tensor(images)
    n = images.shape[0]
    out = []
    with torch.no_grad():
        for bi, s in enumerate(range(0, n, batch)):
            if bi >= n // batch and in_dist is False:
                break
            out.append(maha_scores_from_features(image_features(images[s:s + batch], sd, cfg), classwise_mean, precision, normalize))
    return np.concatenate(out) if out else np.zeros((0,), np.float32)
