"""Hand-written tcgen05 3xTF32 GEMM (csrc/ua2_umma.cu: TMA -> on-chip hi/lo split into tensor memory -> tcgen05.mma kind::tf32 with the
A operand from TMEM -> tcgen05.ld epilogue; stream-K over (tile, k-block) units) through the C ABI (ua2_tc_linear_f32), against an fp64
product of the same fp32 operands.  It replaces F.linear of lit_model.py:424, 511, 592-595 for many-row calls (forward_prefix,
batched frames), the codec transformer's and the flow decoder's linears and the wide convolutions as implicit GEMMs.

Bar: fp32-class.  3xTF32 drops the lo*lo term (~2^-22 relative); the TMEM accumulator rounds toward zero once per k-step of 8, a bias
of about steps / 2 ulp of the running sum (measured, tests/test_zz_options_gpu.py): tol = max(4e-6, 1.5 * (3K / 8) * 2^-24)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = [  # (M, N, K, swiglu, rmsnorm, residual): every tile width (32 / 64 / 128 / 256 activation rows), ragged M / N / K (TMA zero fill),
    # one-CTA problems and stream-K splits inside tiles, two weight matrices with N % 128 != 0
    (32, 256, 64, 0, 0, 0), (32, 128, 32, 0, 0, 0), (1, 128, 128, 0, 0, 0), (7, 260, 100, 0, 0, 0), (32, 5120, 3072, 0, 1, 0),
    (32, 3072, 8192, 0, 0, 1), (32, 8192, 3072, 1, 1, 0), (64, 1024, 512, 0, 0, 0), (50, 12300, 2048, 0, 0, 0), (128, 768, 768, 0, 1, 1),
    (147, 2304, 768, 0, 0, 0), (300, 512, 2048, 1, 0, 0), (1000, 1536, 1040, 0, 0, 0), (1024, 5120, 3072, 0, 1, 0), (513, 1344, 512, 1, 0, 0),
]


@pytest.mark.parametrize("M,N,K,sw,rn,rs", CASES)
def test_tc_linear_matches_fp64_product(M, N, K, sw, rn, rs):
    from uniaudio2_b200 import _lib

    L, P = _lib.lib(), _lib.ptr
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    x = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    W2 = torch.randn(N, K, generator=g) / K ** 0.5 if sw else None
    nw = torch.rand(K, generator=g) + 0.5 if rn else None
    r = torch.randn(M, N, generator=g) if rs else None
    xn = (x * torch.rsqrt((x ** 2).mean(-1, keepdim=True) + 1e-5) * nw).double() if rn else x.double()  # lit_model.py:883-890
    ref = xn @ W.double().t()
    if sw:
        ref = torch.nn.functional.silu(ref) * (xn @ W2.double().t())  # lit_model.py:591-594
    if rs:
        ref = ref + r.double()
    xd, Wd = x.cuda(), W.cuda()
    W2d = W2.cuda() if sw else None
    nwd = nw.cuda() if rn else None
    rd = r.cuda() if rs else None
    y = torch.full((M, N), float("nan"), device="cuda")
    _lib.check(L.ua2_tc_linear_f32(P(xd), P(Wd), P(W2d), P(nwd), 1e-5, P(rd), P(y), M, N, K, None))
    torch.cuda.synchronize()
    assert bool(torch.isfinite(y).all())
    tol = max(4e-6, 1.5 * (3 * K / 8) * 2.0 ** -24) * (2.0 if sw else 1.0)
    assert float((y.cpu().double() - ref).abs().max()) <= tol * max(1.0, float(ref.abs().max()))


def test_tc_linear_is_deterministic_and_rejects_bad_shapes():
    """Stream-K side slots are summed in CTA order: two runs are bit-equal.  Argument checks of the C ABI."""
    from uniaudio2_b200 import _lib

    L, P = _lib.lib(), _lib.ptr
    x = torch.randn(40, 1024, device="cuda")
    W = torch.randn(1280, 1024, device="cuda") / 32
    ys = []
    for _ in range(2):
        y = torch.empty(40, 1280, device="cuda")
        _lib.check(L.ua2_tc_linear_f32(P(x), P(W), None, None, 1e-5, None, P(y), 40, 1280, 1024, None))
        torch.cuda.synchronize()
        ys.append(y)
    assert torch.equal(ys[0], ys[1])
    with pytest.raises(ValueError):
        _lib.check(L.ua2_tc_linear_f32(P(x), P(W), None, None, 1e-5, None, P(ys[0]), 40, 1280, 1022, None))
    with pytest.raises(ValueError):
        _lib.check(L.ua2_tc_linear_f32(None, P(W), None, None, 1e-5, None, P(ys[0]), 40, 1280, 1024, None))
