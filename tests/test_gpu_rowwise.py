"""LayerNorm / embedding kernels vs torch fp32 (HF modeling_clip.py:202-218, 359-361, 677)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("D", [128, 256, 768, 1024])
@pytest.mark.parametrize("M", [1, 197, 1000])
@pytest.mark.parametrize("out_f16", [True, False])
def test_layernorm(engine_factory, M, D, out_f16):
    eng, _, _ = engine_factory("tiny", 5, 8)
    g = torch.Generator(device="cuda").manual_seed(D + M)
    x = torch.randn(M, D, device="cuda", generator=g) * 3 + 0.5
    gamma = torch.randn(D, device="cuda", generator=g) * 0.1 + 1
    beta = torch.randn(D, device="cuda", generator=g) * 0.1
    out = eng.dbg_layernorm(x, gamma, beta, 1e-5, out_f16)
    ref = torch.nn.functional.layer_norm(x, (D,), gamma, beta, 1e-5)
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    assert err <= (4e-2 if out_f16 else 2e-5), err


@pytest.mark.parametrize("cfg_name,b", [("tiny", 5), ("small", 3), ("ViT-B/16", 2)])
def test_embeddings_and_pre_layernorm(engine_factory, cfg_name, b):
    from mcm_b200 import synth
    from oracle import clip_mcm_oracle as O
    eng, sd, cfg = engine_factory(cfg_name, 5, 8)
    imgs = torch.from_numpy(synth.synth_images(b, 11))
    x = eng.dbg_embed(imgs.cuda()).cpu().reshape(b, cfg.seq, cfg.width)
    ref = O.vision_embeddings(imgs, sd, cfg)
    ref = O._ln(ref, sd["vision_model.pre_layrnorm.weight"], sd["vision_model.pre_layrnorm.bias"], cfg.eps)
    # fp16 patch pixels x fp16 filter, fp32 accumulate: relative 2^-8 per product, averaged over 3p^2 terms
    err = (x - ref).abs().max().item()
    assert err <= 3e-2, err
    # the CLS row involves no GEMM: fp32-exact
    assert (x[:, 0] - ref[:, 0]).abs().max().item() <= 1e-5
