"""Fused tail: CLS pool + post_layernorm + visual_projection + normalise + cosine + score reductions
(HF modeling_clip.py:685-686,860-861; utils/detection_util.py:226,232-248) vs the oracle, all fp32."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SCORES = ["MCM", "max-logit", "energy", "entropy", "var"]


@pytest.mark.parametrize("cfg_name,K,b,T", [("tiny", 10, 7, 1.0), ("small", 100, 9, 2.0), ("ViT-B/16", 1000, 5, 1.0),
                                            ("ViT-B/16", 1, 4, 1.0), ("tiny", 1001, 3, 0.01), ("tiny", 33, 2, 100.0)])
def test_tail_matches_oracle(engine_factory, cfg_name, K, b, T):
    from mcm_b200 import synth
    from oracle import clip_mcm_oracle as O
    eng, sd, cfg = engine_factory(cfg_name, 5, 16)
    bank = synth.synth_unit_bank(K, cfg.proj, 3)
    eng.set_text_bank(bank * 2.5)          # un-normalised on purpose: the engine normalises rows (:231)
    g = torch.Generator().manual_seed(K + b)
    x = torch.randn(b, cfg.seq, cfg.width, generator=g) * 2 + 0.3
    pooled = O._ln(x[:, 0], sd["vision_model.post_layernorm.weight"], sd["vision_model.post_layernorm.bias"], cfg.eps)
    feats_ref = pooled @ sd["visual_projection.weight"].t()
    xd = x.reshape(b * cfg.seq, cfg.width).cuda()
    for sc in SCORES:
        feats, scores = eng.dbg_tail(xd, b, T, sc)
        torch.cuda.synchronize()
        assert (feats.cpu() - feats_ref).abs().max().item() <= 2e-4
        ref = O.scores_from_features(feats_ref, torch.from_numpy(bank), T, sc)
        got = scores.cpu().numpy()
        assert got.dtype == np.float32 and got.shape == (b,)
        np.testing.assert_allclose(got, ref, rtol=2e-4, atol=2e-6, err_msg=sc)
