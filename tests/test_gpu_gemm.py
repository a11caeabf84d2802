"""tcgen05 GEMM (TMA -> smem ring -> tcgen05.mma -> TMEM -> fused epilogue) vs torch fp32 of the
same fp16 operands.  Replaces the nn.Linear SGEMMs of HF CLIP (modeling_clip.py:310-312,334,348-350)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, w, bias, resid, epi):
    acc = a.float() @ w.float().t() + bias
    if epi == 1:
        acc = acc * torch.sigmoid(1.702 * acc)      # quick_gelu, HF activations.py:117-123
    if epi == 2:
        acc = acc + resid
    return acc


SHAPES = [
    (128, 128, 64),      # one tile, one k-block
    (128, 128, 256),     # k loop
    (128, 256, 128),     # BLOCK_N = 256
    (256, 384, 128),     # BLOCK_N = 128, 3 n-tiles, 2 m-tiles
    (200, 128, 128),     # ragged M (rows masked by m_valid, TMA zero-fills)
    (1576, 2304, 768),   # 8 images x 197 tokens, QKV shape of ViT-B/16
    (1576, 768, 3072),   # fc2 shape, long k loop (48 k-blocks wrap the smem ring many times)
    (197 * 160, 768, 768),  # more tiles than SMs: persistent loop + both TMEM accumulator stages
]


@pytest.mark.parametrize("epi", [0, 1, 2])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_matches_torch(engine_factory, M, N, K, epi):
    eng, _, _ = engine_factory("tiny", 5, 8)
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K + epi)
    a = (torch.randn(M, K, device="cuda", generator=g)).to(torch.float16)
    w = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).to(torch.float16)
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g) if epi == 2 else None
    out = eng.dbg_gemm(a, w, bias, resid, epi)
    torch.cuda.synchronize()
    ref = _ref(a, w, bias, resid, epi)
    err = (out.float() - ref).abs().max().item()
    # fp32 accumulation of exact fp16 products: only the summation order and (epi 0/1) the final
    # fp16 rounding of |values| <~ 8 differ
    tol = 2e-3 if epi == 2 else 6e-2
    assert err <= tol, f"M={M} N={N} K={K} epi={epi}: max|d|={err}"
    if epi != 2:
        rel = ((out.float() - ref).abs() / (ref.abs() + 1.0)).max().item()
        assert rel <= 1e-2


def test_gemm_in_place_residual(engine_factory):
    eng, _, _ = engine_factory("tiny", 5, 8)
    g = torch.Generator(device="cuda").manual_seed(1)
    M, N, K = 640, 256, 512
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.float16)
    w = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).to(torch.float16)
    bias = torch.randn(N, device="cuda", generator=g)
    x = torch.randn(M, N, device="cuda", generator=g)
    ref = _ref(a, w, bias, x.clone(), 2)
    import ctypes as C
    from mcm_b200 import _lib
    rc = eng._lib.mcm_dbg_gemm(eng._h, C.c_void_p(a.data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(bias.data_ptr()),
                               C.c_void_p(x.data_ptr()), C.c_void_p(x.data_ptr()), M, N, K, 2, eng._stream())
    _lib.check(rc, eng._h)
    torch.cuda.synchronize()
    assert (x - ref).abs().max().item() <= 2e-3
