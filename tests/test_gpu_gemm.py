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


# ---- LayerNorm folded into the projection (csrc/gemm_tcgen05.cuh): producer epilogue writes the fp16
#      copy + partial row statistics, consumer epilogue applies mu / rstd; HF:371-372,380-381 ----
LN_SHAPES = [
    (256, 256, 128, 256),     # M, D (= N of the producer, K of the consumer), K of the producer, N of the consumer
    (200, 128, 64, 384),      # ragged M, BLOCK_N = 128 on both sides
    (1576, 768, 768, 2304),   # out_proj -> q/k/v of ViT-B/16
    (197 * 160, 768, 128, 256),   # more tiles than clusters: residual prefetch across tiles
    (197 * 160, 768, 64, 2304),   # q/k/v shape of ViT-B/16 at 160 images: 15 tiles per cluster, chunk stores, every ring wrap
    (197 * 40, 1024, 64, 3072),   # the same with K = 1024 (ViT-L/14)
]


@pytest.mark.parametrize("gelu,in_place", [(False, False), (True, False), (False, True)])
@pytest.mark.parametrize("M,D,K1,N2", LN_SHAPES)
def test_gemm_layernorm_fold_matches_torch(engine_factory, M, D, K1, N2, gelu, in_place):
    eng, _, _ = engine_factory("tiny", 5, 8)
    g = torch.Generator(device="cuda").manual_seed(M + D + K1 + N2)
    a = torch.randn(M, K1, device="cuda", generator=g).to(torch.float16)
    w_o = (torch.randn(D, K1, device="cuda", generator=g) * K1 ** -0.5).to(torch.float16)
    b_o = torch.randn(D, device="cuda", generator=g)
    # a residual stream with a per-row offset and spread, so that mu and rstd matter
    resid = torch.randn(M, D, device="cuda", generator=g) * (0.5 + torch.rand(M, 1, device="cuda", generator=g) * 3) \
        + torch.randn(M, 1, device="cuda", generator=g)
    gamma = 1.0 + 0.2 * torch.randn(D, device="cuda", generator=g)
    beta = 0.3 * torch.randn(D, device="cuda", generator=g)
    w_p = torch.randn(N2, D, device="cuda", generator=g) * D ** -0.5
    b_p = torch.randn(N2, device="cuda", generator=g)

    # in_place aliases the residual and the output like the forward does: the epilogue then runs through TMA
    x, x16, stats = eng.dbg_gemm_resid_ln(a, w_o, b_o, resid, in_place=in_place)
    w16, c, d = eng.dbg_fold_ln(w_p, gamma, beta, b_p)
    y = eng.dbg_gemm_ln(x16, w16, d, c, stats, D, gelu=gelu)
    torch.cuda.synchronize()

    x_ref = _ref(a, w_o, b_o, resid, 2)
    assert (x - x_ref).abs().max().item() <= 2e-3
    assert torch.equal(x16, x.to(torch.float16))                      # the fp16 copy is the rounded fp32 result
    s = stats.sum(0)
    assert torch.allclose(s[:, 0], x.sum(1), rtol=1e-4, atol=1e-2)    # partial sums add up to the row statistics
    assert torch.allclose(s[:, 1], (x * x).sum(1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(c, w16.float().sum(1), rtol=1e-4, atol=1e-4)
    assert torch.allclose(d, b_p + w_p @ beta, rtol=1e-4, atol=1e-4)

    xn = torch.nn.functional.layer_norm(x_ref, (D,), gamma, beta, eng.cfg.eps)
    y_ref = xn @ w_p.t() + b_p
    if gelu:
        y_ref = y_ref * torch.sigmoid(1.702 * y_ref)
    # same error budget as a projection of the fp16-rounded LayerNorm output (operand rounding 2^-11)
    err = (y.float() - y_ref).abs().max().item()
    rel = ((y.float() - y_ref).abs() / (y_ref.abs() + 1.0)).max().item()
    assert err <= 6e-2 and rel <= 1e-2, f"M={M} D={D} N={N2}: max|d|={err} rel={rel}"


# ---- the residual epilogue the forward runs: the residual stream is an fp16 (hi, lo) pair updated in place
#      (EPI_BIAS_RESID_H2_LN: through TMA for K <= 1024, through the LSU for longer K) ----
@pytest.mark.parametrize("M,N,K", [(256, 256, 128), (200, 128, 64), (1576, 768, 768), (1576, 768, 3072), (197 * 160, 768, 768),
                                   (197 * 40, 1024, 4096), (300, 384, 1088)])
def test_gemm_residual_pair_in_place(engine_factory, M, N, K):
    eng, _, _ = engine_factory("tiny", 5, 8)
    g = torch.Generator(device="cuda").manual_seed(M + 5 * N + K)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.float16)
    w = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).to(torch.float16)
    bias = torch.randn(N, device="cuda", generator=g)
    x = torch.randn(M, N, device="cuda", generator=g) * (0.5 + 3 * torch.rand(M, 1, device="cuda", generator=g)) \
        + torch.randn(M, 1, device="cuda", generator=g)
    x_hi = x.to(torch.float16)
    x_lo = (x - x_hi.float()).to(torch.float16)
    x_in = x_hi.double() + x_lo.double()                      # the value the pair carries (~22 bits of x)
    ref = x_in + a.double() @ w.double().t() + bias.double()
    stats = eng.dbg_gemm_resid_h2(a, w, bias, x_hi, x_lo)
    torch.cuda.synchronize()
    got = x_hi.double() + x_lo.double()
    err = (got - ref).abs().max().item()
    assert err <= 2e-3, err                                     # fp32 accumulation order (same bound as the fp32 epilogue)
    # the pair is a proper split of the fp32 result: hi = fp16(v), so |lo| = |v - hi| <= half an ulp of hi (2^-11 relative)
    assert ((x_hi.double() - got).abs() <= 2.0 ** -11 * got.abs() + 1e-7).all()
    assert ((got - ref).abs() / (ref.abs() + 1e-3)).median().item() <= 2e-6
    s = stats.sum(0).double()
    assert torch.allclose(s[:, 0], got.sum(1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(s[:, 1], (got * got).sum(1), rtol=1e-4, atol=1e-2)
