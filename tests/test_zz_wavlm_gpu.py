"""GPU: the WavLM encoder drop-in (csrc/ua2_wavlm.cu behind tools/tokenizer/ReasoningCodec_film/models/modeling_wavlm.py) against
hidden states of the real transformers.WavLMModel (tests/golden/frontend_golden.pt) and against the oracle (pinned to that class by
tests/test_wavlm_oracle.py) on fresh inputs, up to the checkpoint geometry (wavlm-base-plus, random weights).

Bar (floating point, fp32 class: 3xTF32 GEMMs, fp32 attention / convolution kernels): 2e-4 of the output scale for every hidden
state; the encoder's own kernels one at a time: 1e-5.  bf16 mode (the reference's autocast arithmetic: bf16 operands, fp32
accumulation, tensor-core attention with the bias added in the softmax): 5e-2 of the output scale and really different; the
tensor-core attention alone: 1e-2 against fp64 softmax of the same bf16 operands (the bar of tests/test_flash_gpu.py)."""
import ctypes as C
import math
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import wavlm_oracle as WO

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(ROOT, "tests", "golden", "frontend_golden.pt"), weights_only=False)


def _model(cfg, sd):
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.modeling_wavlm import WavLMConfig, WavLMModel

    m = WavLMModel(WavLMConfig(**cfg))
    m.load_state_dict(sd, strict=True)
    return m.to("cuda:0")


def _rel(a, b):
    return float((a - b).abs().max()) / max(1.0, float(b.abs().max()))


def _ops(op, a, b, c, d, y, *ints):
    from uniaudio2_b200 import _lib

    i = list(ints) + [0] * (5 - len(ints))
    _lib.check(_lib.lib().ua2_wavlm_ops_f32(op, _lib.ptr(a), _lib.ptr(b), _lib.ptr(c), _lib.ptr(d), _lib.ptr(y), *i, _lib.current_stream()), f"op {op}")
    torch.cuda.synchronize()
    return y.cpu()


@pytest.mark.parametrize("D,groups,K,T,B", [(768, 16, 128, 70, 2), (64, 4, 16, 45, 1), (96, 2, 8, 33, 3)])
def test_positional_convolution_operator(D, groups, K, T, B):
    g = torch.Generator().manual_seed(K + T)
    cg = D // groups
    h = torch.randn(B, T, D, generator=g)
    w = torch.randn(D, cg, K, generator=g) / math.sqrt(cg * K)
    bias = 0.1 * torch.randn(D, generator=g)
    ref = F.gelu(F.conv1d(h.transpose(1, 2), w, bias, padding=K // 2, groups=groups)[:, :, :-1]).transpose(1, 2)
    y = _ops(0, h.cuda(), w.cuda(), bias.cuda(), None, torch.full((B, T, D), float("nan"), device="cuda"), B, T, D, cg, K)
    assert _rel(y, ref) < 1e-5


def test_gate_operator():
    g = torch.Generator().manual_seed(4)
    B, T, H, hs = 2, 37, 12, 64
    h = torch.randn(B, T, H * hs, generator=g)
    sd = {"gru_rel_pos_linear.weight": torch.randn(8, hs, generator=g) / 8, "gru_rel_pos_linear.bias": 0.1 * torch.randn(8, generator=g),
          "gru_rel_pos_const": 1 + 0.3 * torch.randn(1, H, 1, 1, generator=g)}
    ref = WO.gate(sd, {"num_attention_heads": H}, "", h)
    y = _ops(1, h.cuda(), sd["gru_rel_pos_linear.weight"].cuda(), sd["gru_rel_pos_linear.bias"].cuda(), sd["gru_rel_pos_const"].reshape(H).contiguous().cuda(),
             torch.full((B, H, T), float("nan"), device="cuda"), B, T, H, hs)
    assert _rel(y, ref) < 1e-6


@pytest.mark.parametrize("hs,T,B,H", [(64, 150, 2, 3), (32, 45, 1, 2), (128, 33, 1, 2)])
def test_biased_attention_operator(hs, T, B, H):
    g = torch.Generator().manual_seed(T)
    q = torch.randn(B, T, H, hs, generator=g)
    k = torch.randn(B, H, T, hs, generator=g)
    v = torch.randn(B, H, T, hs, generator=g)
    gate = 1 + torch.rand(B, H, T, generator=g)
    tab = torch.randn(H, 2 * T - 1, generator=g)
    i = torch.arange(T)[:, None]
    j = torch.arange(T)[None, :]
    bias = gate[..., None] * tab[:, (j - i + T - 1)][None]
    ref = F.scaled_dot_product_attention(q.permute(0, 2, 1, 3), k, v, attn_mask=bias).permute(0, 2, 1, 3).reshape(B * T, H * hs)
    d = torch.cat([gate.reshape(-1), tab.reshape(-1)]).cuda()
    y = _ops(2, q.reshape(B * T, H * hs).contiguous().cuda(), k.contiguous().cuda(), v.contiguous().cuda(), d,
             torch.full((B * T, H * hs), float("nan"), device="cuda"), B, T, H, hs)
    assert _rel(y, ref) < 1e-5


@pytest.mark.parametrize("name", ["small", "mid"])
def test_wavlm_matches_transformers_golden(gold, name):
    c = gold["wavlm_" + name]
    m = _model(c["cfg"], WO.random_state_dict(c["cfg"], c["seed"]))
    out = m(c["wav16"].cuda(), output_hidden_states=True)
    assert len(out.hidden_states) == len(c["hidden_states"])
    for i, (a, b) in enumerate(zip(out.hidden_states, c["hidden_states"])):
        assert a.shape == b.shape and bool(torch.isfinite(a).all())
        assert _rel(a.cpu(), b) < 2e-4, (name, i, _rel(a.cpu(), b))
    assert torch.equal(out.last_hidden_state, out.hidden_states[-1])
    # the fused stack / slice / mean of get_wavlm_feature, and batch slicing
    n = c["cfg"]["num_hidden_layers"]
    mean = m.hidden_states_mean(c["wav16"].cuda(), 1, n + 1)
    want = torch.stack(c["hidden_states"], 1)[:, 1:n + 1].mean(1)
    assert _rel(mean.cpu(), want) < 2e-4
    assert torch.equal(m.hidden_states_mean(c["wav16"].cuda(), 1, n + 1), mean)  # deterministic
    m.MAX_BATCH = 1
    sliced = m(c["wav16"].cuda(), output_hidden_states=True)
    assert _rel(sliced.hidden_states[-1].cpu(), c["hidden_states"][-1]) < 2e-4 and _rel(sliced.hidden_states[0].cpu(), c["hidden_states"][0]) < 2e-4


def test_wavlm_checkpoint_geometry_vs_oracle():
    """wavlm-base-plus geometry (768 wide, 12 heads, 7 convolutions, k = 128 positional convolution in 16 groups), random weights, 3 s
    clips + the 160 appended zeros: hidden states 6..9 as AudioDiffusion1D.get_wavlm_feature takes them."""
    cfg = dict(WO.BASE_PLUS, num_hidden_layers=9)  # layers 10-12 never influence hidden_states[6:10]
    sd = WO.random_state_dict(cfg, 5)
    g = torch.Generator().manual_seed(8)
    wav16 = torch.randn(2, 48160, generator=g) * 0.2
    with torch.no_grad():
        hs = WO.hidden_states(sd, cfg, wav16)
    m = _model(cfg, sd)
    assert m.num_frames(48160) == hs[0].shape[1] == 150
    out = m(wav16.cuda(), output_hidden_states=True).hidden_states
    for i, (a, b) in enumerate(zip(out, hs)):
        assert _rel(a.cpu(), b) < 2e-4, (i, _rel(a.cpu(), b))
    mean = m.hidden_states_mean(wav16.cuda(), 6, 10).cpu()
    assert _rel(mean, torch.stack(hs, 1)[:, 6:10].mean(1)) < 2e-4


def test_get_wavlm_feature_end_to_end():
    """AudioDiffusion1D.get_wavlm_feature (:359-370): 24 kHz clip -> resample + 160 zeros -> mean of hidden states 6..9 -> (B, D, frames)."""
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.AudioDiffusion1D import AudioDiffusion1D

    cfg = dict(WO.BASE_PLUS, hidden_size=64, num_attention_heads=2, intermediate_size=64, num_hidden_layers=10, conv_dim=(64,) * 7,
               num_conv_pos_embeddings=16, num_conv_pos_embedding_groups=4)
    sd = WO.random_state_dict(cfg, 1)
    g = torch.Generator().manual_seed(2)
    wav24 = torch.randn(2, 1, 36000, generator=g) * 0.2  # 1.5 s -> 24160 samples at 16 kHz -> 75 frames
    with torch.no_grad():
        ref = WO.get_wavlm_feature(sd, cfg, wav24, len_semantic=30)
    host = AudioDiffusion1D.__new__(AudioDiffusion1D)  # the method reads wavlm_encoder / wavlm_transfer only
    torch.nn.Module.__init__(host)
    AudioDiffusion1D.attach_wavlm_encoder(host, _model(cfg, sd))
    out = AudioDiffusion1D.get_wavlm_feature(host, wav24.cuda(), 30)
    assert out.shape == ref.shape == (2, 64, 60)
    assert _rel(out.cpu(), ref) < 2e-4


def test_wavlm_interface_errors():
    from uniaudio2_b200 import _lib

    cfg = dict(hidden_size=64, num_attention_heads=2, intermediate_size=128, num_hidden_layers=2, conv_dim=(64, 64, 64), conv_kernel=(10, 3, 2),
               conv_stride=(5, 2, 2), conv_bias=False, num_conv_pos_embeddings=16, num_conv_pos_embedding_groups=4)
    m = _model(cfg, WO.random_state_dict(dict(cfg, num_buckets=320, max_bucket_distance=800, layer_norm_eps=1e-5), 3))
    with pytest.raises(ValueError):
        m(torch.zeros(1, 1, 4000, device="cuda"))           # (B, samples) only
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 20, device="cuda"))                # shorter than the receptive field
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 4000, device="cuda"), attention_mask=torch.ones(1, 4000, device="cuda"))
    with pytest.raises(ValueError):
        m.hidden_states_mean(torch.zeros(1, 4000, device="cuda"), 2, 9)   # range beyond the model's layers
    with pytest.raises((ValueError, _lib.Ua2Error)):
        m(torch.zeros(1, 200, device="cuda"))               # fewer than 32 frames in the batch


@pytest.mark.parametrize("B,H,T", [(2, 3, 150), (1, 2, 128), (1, 2, 300), (2, 12, 1500), (1, 1, 1)])
@pytest.mark.parametrize("sbuf", [0, 1, 2])
def test_biased_flash_attention_operator(B, H, T, sbuf):
    """flash_bf16_kernel<.., BIAS>: q / k / v as bf16 (B, H, T, 64), bias gate[b, h, i] * tab[h, j - i + T - 1] added in the softmax warps."""
    from uniaudio2_b200 import _lib

    g = torch.Generator().manual_seed(B * 1000 + H * 10 + T)
    q = (torch.randn(B, H, T, 64, generator=g) * 2).bfloat16()
    k = (torch.randn(B, H, T, 64, generator=g) * 2).bfloat16()
    v = torch.randn(B, H, T, 64, generator=g).bfloat16()
    gate = 1 + torch.rand(B, H, T, generator=g)
    tab = torch.randn(H, 2 * T - 1, generator=g) * 2
    i = torch.arange(T)[:, None]
    j = torch.arange(T)[None, :]
    bias = (gate[..., None] * tab[:, (j - i + T - 1)][None]).double()
    s_ = q.double() @ k.double().transpose(-1, -2) / 8.0 + bias
    ref = (torch.softmax(s_, dim=-1) @ v.double()).permute(0, 2, 1, 3).reshape(B, T, H * 64)
    d = torch.cat([gate.reshape(-1), tab.reshape(-1)]).cuda()
    _lib.check(_lib.lib().ua2_set_global_option(b"flash_sbuf", sbuf))
    try:
        y = _ops(3, q.cuda(), k.cuda(), v.cuda(), d, torch.full((B, T, H * 64), float("nan"), device="cuda"), B, T, H, 64)
    finally:
        _lib.check(_lib.lib().ua2_set_global_option(b"flash_sbuf", 0))
    assert bool(torch.isfinite(y).all())
    err = float((y.double() - ref).abs().max())
    assert err <= 1e-2 * max(1.0, float(ref.abs().max())), err


def test_wavlm_bf16_mode(gold):
    """The option that mirrors the reference's autocast: every hidden state close to the fp32 reference, not equal to the fp32-class
    result, and switching it off gives the fp32-class result back bit for bit.  `mid` fixture (head size 64) and checkpoint geometry."""
    c = gold["wavlm_mid"]
    cases = [(c["cfg"], WO.random_state_dict(c["cfg"], c["seed"]), c["wav16"], c["hidden_states"])]
    cfg = dict(WO.BASE_PLUS, num_hidden_layers=9)
    sd = WO.random_state_dict(cfg, 5)
    wav16 = torch.randn(2, 48160, generator=torch.Generator().manual_seed(8)) * 0.2
    with torch.no_grad():
        cases.append((cfg, sd, wav16, WO.hidden_states(sd, cfg, wav16)))
    for cfg, sd, wav, ref in cases:
        m = _model(cfg, sd)
        y32 = m(wav.cuda(), output_hidden_states=True).hidden_states
        m.set_option("bf16", 1)
        y16 = m(wav.cuda(), output_hidden_states=True).hidden_states
        m.set_option("bf16", 0)
        y32b = m(wav.cuda(), output_hidden_states=True).hidden_states
        for i, (a, b, r) in enumerate(zip(y16, y32, ref)):
            assert bool(torch.isfinite(a).all())
            err = _rel(a.cpu(), r)
            assert 1e-6 < err < 5e-2, (i, err)
            assert torch.equal(b, y32b[i])
    small = gold["wavlm_small"]  # head size 32: the tensor-core attention does not serve it
    m = _model(small["cfg"], WO.random_state_dict(small["cfg"], small["seed"]))
    m.set_option("bf16", 1)
    with pytest.raises(Exception):
        m(small["wav16"].cuda())
