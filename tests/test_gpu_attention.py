"""softmax(q k^T / 8) v per (image, head) vs torch fp32 on the same fp16 q, k, v
(HF modeling_clip.py:261-279 / sdpa, no mask, not causal)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("b,S,H", [(1, 50, 2), (3, 197, 4), (2, 197, 12), (2, 257, 16), (1, 16, 1), (1, 17, 1)])
def test_attention(engine_factory, b, S, H):
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
    assert err <= 3e-2, err     # fp16 probabilities / fp16 output rounding
