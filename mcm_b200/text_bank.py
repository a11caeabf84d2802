"""Prompt-bank builder: encode the K class prompts ONCE.

Two modes:

* single template -- exactly what the reference does inside its batch loop
  (``utils/detection_util.py:228-231``): ``"a photo of a {c}"`` (no trailing period), text
  features L2-normalised per class;
* template ensemble (BASELINE config 3, "80 templates averaged") -- the reference ships the 80
  OpenAI prompt templates as ``utils/imagenet_templates.py:openai_imagenet_template`` (a list of
  ``lambda c: str``) but never uses them; this builder accepts that list (or any list of callables /
  ``str.format`` patterns), and forms the standard CLIP zero-shot ensemble: per-template features
  L2-normalised, averaged over templates, re-normalised.

The text tower is not a kernel target (it runs once per label set); it stays the HF module.
Banks are cached per (checkpoint, labels, templates) so repeated ``get_ood_scores_clip`` calls on
the ID and OOD streams share one encoding.
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional, Sequence, Union

import torch

Template = Union[str, Callable[[str], str]]

REFERENCE_TEMPLATE = "a photo of a {}"          # utils/detection_util.py:228

_CACHE = {}


def render(template: Template, name: str) -> str:
    return template(name) if callable(template) else template.format(name)


def _features(net, tokenizer, texts, chunk):
    out = []
    for s in range(0, len(texts), chunk):
        tok = tokenizer(texts[s:s + chunk], padding=True, return_tensors="pt")
        f = net.get_text_features(input_ids=tok["input_ids"], attention_mask=tok["attention_mask"])
        f = f.pooler_output if hasattr(f, "pooler_output") else f      # transformers >= 5
        out.append(f.float().cpu())
    return torch.cat(out)


@torch.no_grad()
def build_text_bank(net, tokenizer, class_names: Iterable[str], templates: Optional[Sequence[Template]] = None,
                    chunk: int = 512, cache_key=None) -> torch.Tensor:
    """``[K, P]`` fp32 unit-norm prompt bank for ``class_names`` (one row per class, class order kept).

    ``net`` needs ``get_text_features(input_ids=, attention_mask=)`` (HF ``CLIPModel`` or
    :class:`mcm_b200.engine.B200ClipNet`); ``tokenizer`` is called like ``CLIPTokenizer``.
    """
    names = [str(c) for c in class_names]
    tpls = list(templates) if templates else [REFERENCE_TEMPLATE]
    key = None
    if cache_key is not None:
        key = (cache_key, tuple(names), tuple(render(t, "{}") for t in tpls))
        if key in _CACHE:
            return _CACHE[key].clone()
    K = len(names)
    acc = None
    for t in tpls:
        f = _features(net, tokenizer, [render(t, n) for n in names], chunk)
        f = f / f.norm(dim=-1, keepdim=True)                         # per-template normalisation
        acc = f if acc is None else acc + f
    bank = acc / len(tpls)
    bank = bank / bank.norm(dim=-1, keepdim=True)
    assert bank.shape[0] == K
    if key is not None:
        _CACHE[key] = bank.clone()
    return bank
