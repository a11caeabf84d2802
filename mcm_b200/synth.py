"""Deterministic synthetic weights, prompt banks and evaluation streams.

There are no pretrained CLIP weights, tokenizer vocabularies or datasets in the
build / GPU containers, so everything the hot path consumes is
generated from seeds.  Generation is numpy ``Philox`` based, hence bit-identical
on every machine that has the same numpy -- that is what lets golden fixtures
made in the authoring container be replayed on the GPU box.

State-dict keys follow HuggingFace ``CLIPModel`` exactly (SURVEY.md section 8b,
"Weights in"), so a real checkpoint's ``state_dict()`` is a drop-in replacement.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch


@dataclass(frozen=True)
class VisionCfg:
    """Shape of a CLIP vision tower (mirrors HF ``CLIPVisionConfig`` + projection_dim)."""
    image_size: int = 224
    patch: int = 16
    width: int = 768
    layers: int = 12
    heads: int = 12
    mlp: int = 3072
    proj: int = 512
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
    "tiny": VisionCfg(224, 32, 128, 2, 2, 256, 64),
    "small": VisionCfg(224, 16, 256, 3, 4, 512, 128),
}

# reference: utils/train_eval_util.py:19-21
CKPT_TO_CFG = {
    "openai/clip-vit-base-patch16": "ViT-B/16",
    "openai/clip-vit-base-patch32": "ViT-B/32",
    "openai/clip-vit-large-patch14": "ViT-L/14",
}


def _rng(seed: int, stream: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=[int(seed), int(stream)]))


def synth_vision_state_dict(cfg: VisionCfg, seed: int = 5) -> dict[str, torch.Tensor]:
    """Seeded fp32 weights for the vision tower + visual projection.

    Unlike HF's ``_init_weights`` (zero biases, unit LayerNorm) every affine term
    is non-trivial so each fused epilogue (bias, LayerNorm gamma/beta, residual)
    is exercised by the parity tests.
    """
    D, F, P, L, p, S = cfg.width, cfg.mlp, cfg.proj, cfg.layers, cfg.patch, cfg.seq
    g = _rng(seed, 0)

    def nrm(shape, std, mean=0.0):
        return torch.from_numpy((g.standard_normal(shape, dtype=np.float32) * np.float32(std) + np.float32(mean)))

    sd = {}
    v = "vision_model."
    sd[v + "embeddings.class_embedding"] = nrm((D,), 0.5)
    sd[v + "embeddings.patch_embedding.weight"] = nrm((D, 3, p, p), (3 * p * p) ** -0.5)
    sd[v + "embeddings.position_embedding.weight"] = nrm((S, D), 0.3)
    for name in ("pre_layrnorm", "post_layernorm"):
        sd[v + name + ".weight"] = nrm((D,), 0.1, 1.0)
        sd[v + name + ".bias"] = nrm((D,), 0.05)
    for i in range(L):
        pre = f"{v}encoder.layers.{i}."
        for ln in ("layer_norm1", "layer_norm2"):
            sd[pre + ln + ".weight"] = nrm((D,), 0.1, 1.0)
            sd[pre + ln + ".bias"] = nrm((D,), 0.05)
        for nm in ("q_proj", "k_proj"):
            sd[pre + f"self_attn.{nm}.weight"] = nrm((D, D), 1.5 * D ** -0.5)
            sd[pre + f"self_attn.{nm}.bias"] = nrm((D,), 0.1)
        sd[pre + "self_attn.v_proj.weight"] = nrm((D, D), D ** -0.5)
        sd[pre + "self_attn.v_proj.bias"] = nrm((D,), 0.05)
        sd[pre + "self_attn.out_proj.weight"] = nrm((D, D), 0.7 * D ** -0.5)
        sd[pre + "self_attn.out_proj.bias"] = nrm((D,), 0.05)
        sd[pre + "mlp.fc1.weight"] = nrm((F, D), D ** -0.5)
        sd[pre + "mlp.fc1.bias"] = nrm((F,), 0.1)
        sd[pre + "mlp.fc2.weight"] = nrm((D, F), 0.7 * F ** -0.5)
        sd[pre + "mlp.fc2.bias"] = nrm((D,), 0.05)
    sd["visual_projection.weight"] = nrm((P, D), D ** -0.5)
    return sd


def synth_images(n: int, seed: int, image_size: int = 224, mean: float = 0.0, std: float = 1.0) -> np.ndarray:
    """[n,3,H,W] float32 'already CLIP-normalised' pixels, N(mean, std)."""
    g = _rng(seed, 1)
    x = g.standard_normal((n, 3, image_size, image_size), dtype=np.float32)
    if std != 1.0:
        x *= np.float32(std)
    if mean != 0.0:
        x += np.float32(mean)
    return x


def synth_images_u8(n: int, seed: int, image_size: int = 224) -> np.ndarray:
    """``[n, H, W, 3]`` uint8 "decoded" pixels (HWC, uniform over 0..255) for the uint8 ingest path."""
    return _rng(seed, 7).integers(0, 256, size=(n, image_size, image_size, 3), dtype=np.uint8)


def synth_prototype_stream(n: int, protos: np.ndarray, seed: int, noise: float) -> np.ndarray:
    """ID stream of the prototype harness: image i = prototype[i % K] + noise * N(0,1)."""
    K = protos.shape[0]
    g = _rng(seed, 2)
    x = g.standard_normal((n,) + protos.shape[1:], dtype=np.float32)
    x *= np.float32(noise)
    x += protos[np.arange(n) % K]
    return x


def synth_unit_bank(K: int, P: int, seed: int) -> np.ndarray:
    """[K,P] float32 unit rows (a stand-in for an encoded prompt bank)."""
    g = _rng(seed, 3)
    b = g.standard_normal((K, P), dtype=np.float32)
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    return b.astype(np.float32)


def centred_prototype_bank(proto_feats: np.ndarray) -> np.ndarray:
    """Prompt bank of the prototype harness: unit rows along (feature_k - mean feature).

    Random-init towers map every image to nearly the same direction; removing the
    common component makes cosines, and so MCM scores, depend on the image.
    """
    f = np.asarray(proto_feats, dtype=np.float64)
    f = f / np.linalg.norm(f, axis=1, keepdims=True)
    c = f - f.mean(axis=0, keepdims=True)
    c /= np.linalg.norm(c, axis=1, keepdims=True)
    return c.astype(np.float32)
