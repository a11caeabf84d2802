"""Run the UNMODIFIED reference ``get_ood_scores_clip`` on CPU -- TEST INFRASTRUCTURE.

Works only where ``/root/reference`` is mounted (the authoring container); the
GPU box never imports this (``REFERENCE_AVAILABLE`` is False there).

The shims change no arithmetic:
  1. ``ShimCLIP(CLIPModel)``: transformers 5.5 returns a ``BaseModelOutputWithPooling``
     from ``get_image_features`` / ``get_text_features`` while the reference
     (``utils/detection_util.py:225,229``) was written for the 4.x API that
     returned the projected tensor -> return ``.pooler_output``.
  2. ``FakeTokenizer``: no vocab offline; deterministic ids with BOS/EOS so the
     EOS pooling of the text tower (HF:577-584) works.
  3. ``Tensor.cuda`` is a no-op on CPU (the function hard-codes ``.cuda()``,
     ``utils/detection_util.py:222-223,229-230``).
"""
from __future__ import annotations

import contextlib
import os
import sys
import types
import zlib

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("MCM_REFERENCE_ROOT", "/root/reference")
REFERENCE_AVAILABLE = os.path.isfile(os.path.join(REFERENCE_ROOT, "utils", "detection_util.py"))


def load_reference_detection_util():
    """Import ``utils.detection_util`` from the reference tree under a private name."""
    import importlib.util
    name = "_mcm_reference_detection_util"
    if name in sys.modules:
        return sys.modules[name]
    path = os.path.join(REFERENCE_ROOT, "utils", "detection_util.py")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


class FakeTokenizer:
    """Stand-in for ``CLIPTokenizer`` (``utils/detection_util.py:216,228``)."""
    BOS, EOS = 49406, 49407

    @classmethod
    def from_pretrained(cls, name, *a, **k):
        return cls()

    def __call__(self, texts, padding=True, return_tensors="pt"):
        rows = []
        for t in texts:
            words = str(t).split()
            ids = [self.BOS] + [1 + (zlib.crc32(w.encode()) % 49000) for w in words][:75] + [self.EOS]
            rows.append(ids)
        L = max(len(r) for r in rows)
        input_ids = torch.full((len(rows), L), self.EOS, dtype=torch.long)
        mask = torch.zeros((len(rows), L), dtype=torch.long)
        for i, r in enumerate(rows):
            input_ids[i, : len(r)] = torch.tensor(r)
            mask[i, : len(r)] = 1
        return {"input_ids": input_ids, "attention_mask": mask}


def make_shim_clip(cfg, vision_sd=None, seed=5, text_layers=2, text_width=128, text_heads=2):
    """HF ``CLIPModel`` with the given vision shape (random text tower, small by
    default to keep the per-batch text re-encode of the reference cheap), wrapped
    so the 4.x-style calls of the reference work.  ``vision_sd`` (HF keys)
    overrides the vision tower + projection weights."""
    from transformers import CLIPConfig, CLIPModel

    class ShimCLIP(CLIPModel):
        bank_override = None   # optional [K,P] tensor returned instead of running the text tower
        captured_bank = None

        def get_image_features(self, *a, **k):
            out = super().get_image_features(*a, **k)
            return out.pooler_output if hasattr(out, "pooler_output") else out

        def get_text_features(self, *a, **k):
            if self.bank_override is not None:
                return self.bank_override.clone()
            out = super().get_text_features(*a, **k)
            out = out.pooler_output if hasattr(out, "pooler_output") else out
            type(self).captured_bank = out.detach().clone()
            return out

    torch.manual_seed(seed)
    config = CLIPConfig(
        vision_config=dict(hidden_size=cfg.width, intermediate_size=cfg.mlp, num_hidden_layers=cfg.layers,
                           num_attention_heads=cfg.heads, patch_size=cfg.patch, image_size=cfg.image_size),
        text_config=dict(hidden_size=text_width, intermediate_size=4 * text_width, num_hidden_layers=text_layers,
                         num_attention_heads=text_heads),
        projection_dim=cfg.proj,
    )
    model = ShimCLIP(config)
    if vision_sd is not None:
        missing, unexpected = model.load_state_dict(vision_sd, strict=False)
        assert not unexpected, unexpected
        assert not [m for m in missing if m.startswith("vision_model.") and "position_ids" not in m
                    or m == "visual_projection.weight"], missing
    return model.eval()


@contextlib.contextmanager
def cpu_cuda_noop():
    """Make ``Tensor.cuda()`` the identity while the reference loop runs on CPU."""
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


class _ListLoader:
    """Minimal DataLoader stand-in: iterable of (images, labels) with ``.dataset``."""

    def __init__(self, images, batch_size):
        self.images = torch.as_tensor(images)
        self.batch_size = batch_size
        self.dataset = range(self.images.shape[0])

    def __len__(self):
        return (self.images.shape[0] + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = self.images.shape[0]
        for s in range(0, n, self.batch_size):
            x = self.images[s:s + self.batch_size]
            yield x, torch.zeros(x.shape[0], dtype=torch.long)


def run_reference_scores(model, images, test_labels, T=1, score="MCM", batch_size=64):
    """``get_ood_scores_clip`` of the reference, unmodified, on CPU."""
    du = load_reference_detection_util()
    du.CLIPTokenizer = FakeTokenizer
    args = types.SimpleNamespace(ckpt="synthetic", model="CLIP", score=score, T=T, batch_size=batch_size)
    loader = _ListLoader(images, batch_size)
    with cpu_cuda_noop():
        return du.get_ood_scores_clip(args, model, loader, list(test_labels))


def run_reference_measures(in_score, out_score):
    """``get_measures(-in, -out)`` exactly as ``get_and_print_results`` calls it
    (``utils/detection_util.py:259``)."""
    du = load_reference_detection_util()
    return du.get_measures(-np.asarray(in_score), -np.asarray(out_score))
