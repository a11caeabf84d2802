"""softmax(q k^T / 8) v per (image, head) vs torch fp32 on the same fp16 q, k, v
(HF modeling_clip.py:261-279 / sdpa, no mask, not causal)."""
import pytest
import torch

from helpers import report

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("b,S,H", [(1, 50, 2), (3, 197, 4), (2, 197, 12), (2, 257, 16), (1, 16, 1), (1, 17, 1),
                                   (1, 255, 2), (2, 256, 3), (3, 257, 2), (37, 257, 16), (1, 258, 2), (1, 290, 1),
                                   (3, 50, 5), (2, 64, 3), (5, 33, 1), (7, 50, 12), (1, 65, 2)])
def test_attention(engine_factory, b, S, H):
    """S <= 64 (ViT-B/32): two (image, head) items per 128-row unit, odd item counts included; S <= 256: tensor-core keys only; S = 257 (ViT-L/14): 256 tensor-core keys + the extra key on the CUDA
    cores; S > 257 falls back to the warp-level mma.sync kernel."""
    eng, _, _ = engine_factory("tiny", 5, 8)
    g = torch.Generator(device="cuda").manual_seed(S * 13 + H)
    D = H * 64
    qkv = torch.randn(b * S, 3 * D, device="cuda", generator=g)
    qkv[:, :2 * D] *= 1.5   # realistic logit spread
    qkv = qkv.to(torch.float16)
    out = eng.dbg_attention(qkv, b, S, H).float().reshape(b, S, H, 64)
    q, k, v = [t.float().reshape(b, S, H, 64).transpose(1, 2) for t in qkv.split(D, dim=1)]
    ref = torch.softmax(q @ k.transpose(-1, -2) * 0.125, dim=-1) @ v
    ref = ref.transpose(1, 2)
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item()
    report("attention", dict(case="random", b=b, S=S, H=H, max_abs_err=err, ref_absmax=ref.abs().max().item()))
    assert err <= 4e-3, err     # fp16 probabilities / fp16 output rounding: measured <= 1.1e-3 on outputs up to 3.8


def test_attention_extra_key_dominates(engine_factory):
    """ViT-L/14 geometry with the 257th key carrying the row maximum for most rows and a distinctive value row:
    exercises the extra-key branch of the row max, the row sum and the O update."""
    eng, _, _ = engine_factory("tiny", 5, 8)
    b, S, H = 3, 257, 4
    D = H * 64
    g = torch.Generator(device="cuda").manual_seed(99)
    qkv = torch.randn(b, S, 3 * D, device="cuda", generator=g)
    u = torch.randn(D, device="cuda", generator=g)
    qkv[:, :, :D] += u                                        # every query shares a component with ...
    qkv[:, 256, D:2 * D] = 1.5 * u                            # ... the key of token 256: logit ~ 0.125 * 1.5 * 64 = 12
    qkv[:, 256, 2 * D:] += 5.0
    qkv = qkv.reshape(b * S, 3 * D).to(torch.float16)
    out = eng.dbg_attention(qkv, b, S, H).float().reshape(b, S, H, 64)
    q, k, v = [t.float().reshape(b, S, H, 64).transpose(1, 2) for t in qkv.split(D, dim=1)]
    prob = torch.softmax(q @ k.transpose(-1, -2) * 0.125, dim=-1)
    ref = (prob @ v).transpose(1, 2)
    torch.cuda.synchronize()
    assert prob[..., 256].max().item() > 0.5                 # the extra key really matters in this case
    err = (out - ref).abs().max().item()
    report("attention", dict(case="extra_key", b=b, S=S, H=H, max_abs_err=err, ref_absmax=ref.abs().max().item()))
    assert err <= 8e-3, err     # measured 2.0e-3 on outputs up to 7.7


@pytest.mark.parametrize("S", [197, 257, 50])
def test_attention_late_maximum(engine_factory, S):
    """Rows whose maximum sits in a late key chunk and beats everything before it by far more than 2^8
    (twice, in two different chunks): the case a lazily updated softmax reference maximum must rescale for."""
    eng, _, _ = engine_factory("tiny", 5, 8)
    b, H = 2, 3
    D = H * 64
    g = torch.Generator(device="cuda").manual_seed(7 + S)
    qkv = torch.randn(b, S, 3 * D, device="cuda", generator=g)
    u = torch.randn(D, device="cuda", generator=g)
    qkv[:, :, :D] = 0.3 * qkv[:, :, :D] + u
    qkv[:, S // 2, D:2 * D] = 3.0 * u        # logit ~ 0.125 * 64 * 3 = 24 above the rest
    qkv[:, S - 7, D:2 * D] = 6.0 * u         # and another 24 above that one
    qkv[:, S // 2, 2 * D:] -= 3.0
    qkv[:, S - 7, 2 * D:] += 2.0
    qkv = qkv.reshape(b * S, 3 * D).to(torch.float16)
    out = eng.dbg_attention(qkv, b, S, H).float().reshape(b, S, H, 64)
    q, k, v = [t.float().reshape(b, S, H, 64).transpose(1, 2) for t in qkv.split(D, dim=1)]
    ref = (torch.softmax(q @ k.transpose(-1, -2) * 0.125, dim=-1) @ v).transpose(1, 2)
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item()
    report("attention", dict(case="late_max", b=b, S=S, H=H, max_abs_err=err, ref_absmax=ref.abs().max().item()))
    assert err <= 4e-3, err     # measured 1e-6: the row is one key
