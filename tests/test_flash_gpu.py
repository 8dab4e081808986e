"""Tensor-core attention of the flow-matching decoder (csrc/ua2_flash.cu: tcgen05 kind::f16 for Q K^T and P V, softmax statistics in
fp32 registers, P handed to the second MMA through tensor memory) through the C ABI (ua2_flash_attn_bf16), against fp64 softmax
attention of the SAME bf16-rounded operands.  It replaces F.scaled_dot_product_attention inside diffusers' Attention as called from
ReasoningCodec_film/models/attention.py:338-357 under the reference's bf16 autocast (reason_tokenizer.py:265).

Bar (floating point, bf16 class): the only roundings beyond the operands are P -> bf16 before the second product (2^-9 relative per
term, averaged over the row) and ex2.approx; tolerance 1e-2 of the output's max magnitude, and the row sums must be consistent
(a constant V gives that constant back to 2e-3)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(q, k, v):
    from uniaudio2_b200 import _lib

    L, P = _lib.lib(), _lib.ptr
    B, H, T, hs = q.shape
    out = torch.full((B, T, H * hs), float("nan"), device="cuda")
    _lib.check(L.ua2_flash_attn_bf16(P(q), P(k), P(v), P(out), B, T, H, hs, None))
    torch.cuda.synchronize()
    return out


def _ref(q, k, v):
    qd, kd, vd = q.double().cpu(), k.double().cpu(), v.double().cpu()
    s = qd @ kd.transpose(-1, -2) / q.shape[-1] ** 0.5
    o = torch.softmax(s, dim=-1) @ vd  # (B, H, T, hs)
    return o.permute(0, 2, 1, 3).reshape(q.shape[0], q.shape[2], -1)


@pytest.mark.parametrize("B,H,T", [(1, 1, 128), (1, 2, 1), (2, 3, 100), (1, 2, 129), (2, 24, 500), (1, 4, 1000), (3, 2, 257), (1, 1, 2048)])
@pytest.mark.parametrize("scale,sbuf", [(1.0, 0), (4.0, 1), (4.0, 2)])
def test_flash_attention_matches_fp64_softmax(B, H, T, scale, sbuf):
    """sbuf: 0 = the launcher's choice, 1 = single score buffer / two CTAs per SM, 2 = double-buffered scores / one CTA per SM."""
    from uniaudio2_b200 import _lib

    _lib.check(_lib.lib().ua2_set_global_option(b"flash_sbuf", sbuf))
    try:
        _check_case(B, H, T, scale)
    finally:
        _lib.check(_lib.lib().ua2_set_global_option(b"flash_sbuf", 0))


def _check_case(B, H, T, scale):
    g = torch.Generator().manual_seed(B * 1000 + H * 10 + T)
    q = (torch.randn(B, H, T, 64, generator=g) * scale).bfloat16().cuda()
    k = (torch.randn(B, H, T, 64, generator=g) * scale).bfloat16().cuda()
    v = torch.randn(B, H, T, 64, generator=g).bfloat16().cuda()
    out = _run(q, k, v)
    ref = _ref(q, k, v)
    assert bool(torch.isfinite(out).all())
    err = float((out.cpu().double() - ref).abs().max())
    assert err <= 1e-2 * max(1.0, float(ref.abs().max())), err


def test_flash_attention_rows_are_normalised_and_deterministic():
    g = torch.Generator().manual_seed(5)
    B, H, T = 2, 5, 333
    q = (torch.randn(B, H, T, 64, generator=g) * 3).bfloat16().cuda()
    k = (torch.randn(B, H, T, 64, generator=g) * 3).bfloat16().cuda()
    v = torch.full((B, H, T, 64), 0.75).bfloat16().cuda()
    a, b = _run(q, k, v), _run(q, k, v)
    assert torch.equal(a, b)
    assert float((a - 0.75).abs().max()) < 2e-3


def test_flash_attention_rejects_other_head_sizes():
    from uniaudio2_b200 import _lib

    L, P = _lib.lib(), _lib.ptr
    q = torch.zeros(1, 1, 8, 32, dtype=torch.bfloat16, device="cuda")
    out = torch.zeros(1, 8, 32, device="cuda")
    assert L.ua2_flash_attn_bf16(P(q), P(q), P(q), P(out), 1, 8, 1, 32, None) != 0
