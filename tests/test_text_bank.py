"""Prompt-bank builder (host logic, CPU): single-template mode reproduces the reference's per-batch
text encoding (utils/detection_util.py:228-231); ensemble mode = normalise, mean, re-normalise."""
import numpy as np
import torch

from mcm_b200 import text_bank
from oracle import reference_shims as R


class _Net:
    """Deterministic stand-in text tower: features depend on the token ids only."""

    def __init__(self, P=16):
        g = torch.Generator().manual_seed(0)
        self.table = torch.randn(50000, P, generator=g)

    def get_text_features(self, input_ids=None, attention_mask=None):
        m = attention_mask.unsqueeze(-1).float()
        return (self.table[input_ids] * m).sum(1)


def test_single_template_matches_reference_prompting():
    net, tok = _Net(), R.FakeTokenizer()
    names = ["goldfish", "great white shark", "hen"]
    bank = text_bank.build_text_bank(net, tok, names)
    t = tok([f"a photo of a {c}" for c in names], padding=True, return_tensors="pt")     # detection_util.py:228
    ref = net.get_text_features(input_ids=t["input_ids"], attention_mask=t["attention_mask"]).float()
    ref = ref / ref.norm(dim=-1, keepdim=True)                                            # :231
    assert bank.shape == (3, 16)
    torch.testing.assert_close(bank, ref, rtol=0, atol=1e-6)


def test_template_ensemble_and_cache():
    net, tok = _Net(), R.FakeTokenizer()
    names = ["tench", "goldfish"]
    tpls = [lambda c: f"a bad photo of a {c}.", "a sculpture of a {}.", lambda c: f"itap of my {c}."]
    bank = text_bank.build_text_bank(net, tok, names, tpls, chunk=1, cache_key="ckpt")
    per = []
    for t in tpls:
        tk = tok([text_bank.render(t, n) for n in names], padding=True, return_tensors="pt")
        f = net.get_text_features(input_ids=tk["input_ids"], attention_mask=tk["attention_mask"])
        per.append(f / f.norm(dim=-1, keepdim=True))
    ref = torch.stack(per).mean(0)
    ref = ref / ref.norm(dim=-1, keepdim=True)
    torch.testing.assert_close(bank, ref, rtol=0, atol=1e-6)
    np.testing.assert_allclose(bank.norm(dim=-1).numpy(), 1.0, atol=1e-6)
    again = text_bank.build_text_bank(None, None, names, tpls, cache_key="ckpt")      # served from the cache
    assert torch.equal(again, bank)
