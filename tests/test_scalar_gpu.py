"""GPU parity of the ScalarModel drop-in (SQ-codec wave encoder / decoder of ReasoningCodec_film) against golden vectors
produced by the UNMODIFIED reference class (scalar24k.py::ScalarModel, tests/golden/scalar_golden.pt) and the CPU oracle.
The production hyper-parameters (sqcodec_config.yaml) are not in the repository: configs here are assumed."""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import scalar_oracle as SO
from oracle.make_golden_scalar import scalar_cfgs

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(cfg, sd):
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.scalar24k import ScalarModel

    m = ScalarModel(cfg.num_bands, cfg.sample_rate, cfg.causal, cfg.num_samples, cfg.downsample_factors, cfg.downsample_kernel_sizes,
                    cfg.upsample_factors, cfg.upsample_kernel_sizes, cfg.latent_hidden_dim, cfg.default_kernel_size,
                    cfg.delay_kernel_size, cfg.init_channel, cfg.res_kernel_size, device="cuda")
    m.load_state_dict({k: v.cuda() for k, v in sd.items()}, strict=True)
    return m


@pytest.mark.parametrize("name", ["causal", "noncausal"])
def test_scalar_model_matches_reference_golden(name):
    fx = torch.load(os.path.join(ROOT, "tests", "golden", "scalar_golden.pt"), weights_only=False)[name]
    cfg = scalar_cfgs()[name]
    m = _build(cfg, fx["sd"])
    y = m.decode(fx["z"].cuda())
    e = m.encode(fx["wav"].cuda())
    torch.cuda.synchronize()
    assert y.shape == fx["decoded"].shape and e.shape == fx["encoded"].shape
    assert float((y.cpu() - fx["decoded"]).abs().max()) < 1e-4
    assert float((e.cpu() - fx["encoded"]).abs().max()) < 1e-4


def test_scalar_model_assumed_full_size_vs_oracle():
    """Assumed SQ-codec geometry (136-d latent @25 Hz -> 24 kHz, x960): 1.2 s of latent frames, decode vs the CPU oracle."""
    cfg = SO.ScalarCfg()
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.scalar24k import ScalarModel

    m = ScalarModel(cfg.num_bands, cfg.sample_rate, cfg.causal, cfg.num_samples, cfg.downsample_factors, cfg.downsample_kernel_sizes,
                    cfg.upsample_factors, cfg.upsample_kernel_sizes, cfg.latent_hidden_dim, cfg.default_kernel_size,
                    cfg.delay_kernel_size, cfg.init_channel, cfg.res_kernel_size)
    g = torch.Generator().manual_seed(1)
    sd = {}
    for k, v in m.state_dict().items():
        if k.endswith("weight_g"):
            sd[k] = 0.5 + torch.rand(v.shape, generator=g)
        elif "activation" in k:
            sd[k] = 0.1 + 0.3 * torch.rand(v.shape, generator=g)
        elif k.endswith("bias"):
            sd[k] = 0.05 * torch.randn(v.shape, generator=g)
        else:
            sd[k] = torch.randn(v.shape, generator=g)
    m = m.cuda()
    m.load_state_dict({k: v.cuda() for k, v in sd.items()})
    z = torch.rand(2, cfg.latent_hidden_dim, 30, generator=g) * 2 - 1
    with torch.no_grad():
        ref = SO.scalar_decode(z, sd, cfg)
    y = m.decode(z.cuda())
    torch.cuda.synchronize()
    assert y.shape == ref.shape == (2, 1, 30 * 960)
    scale = float(ref.abs().max())
    assert float((y.cpu() - ref).abs().max()) < 1e-4 * max(1.0, scale)


@pytest.mark.parametrize("B,Cin,Cout,T,K,stride,dil,pl,pr", [(2, 1024, 1024, 1500, 4, 4, 1, 0, 0), (1, 768, 768, 1500, 4, 4, 1, 0, 0),
                                                             (2, 1024, 1024, 750, 2, 2, 1, 0, 0), (1, 48, 48, 999, 7, 1, 9, 27, 27),
                                                             (1, 136, 1536, 50, 5, 1, 1, 2, 2)])
def test_general_conv1d(B, Cin, Cout, T, K, stride, dil, pl, pr):
    """Non-causal / non-overlapping strided nn.Conv1d forms: d_conv_whisper / d_conv_wavlm (k4 s4), d_conv_embedding_* (k2 s2)
    (AudioDiffusion1D.py:244-251, :515-518), ScalarModel's dilated k7 and delay convs - with bias + PReLU + residual fused."""
    from uniaudio2_b200 import _lib

    L = _lib.lib()
    g = torch.Generator().manual_seed(K + T)
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn(Cout, Cin, K, generator=g) / (Cin * K) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    slope = torch.tensor([0.2])
    ref = F.prelu(F.conv1d(F.pad(x, (pl, pr)), w, b, stride=stride, dilation=dil), slope)
    use_res = Cin == Cout and stride == 1 and ref.shape[-1] == T
    if use_res:
        ref = ref + x
    xd, wd, bd, sd_ = x.cuda(), w.cuda(), b.cuda(), slope.cuda()
    y = torch.empty(*ref.shape, device="cuda")
    _lib.check(L.ua2_conv1d_f32(_lib.ptr(xd), _lib.ptr(wd), _lib.ptr(bd), _lib.ptr(sd_), _lib.ptr(xd) if use_res else None, _lib.ptr(y), B,
                                Cin, Cout, T, K, stride, dil, pl, pr, None))
    torch.cuda.synchronize()
    assert float((y.cpu() - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))


def test_film_interp_linear_bias():
    """time_film, nearest x2.5 interpolation and the biased fusion Linear of AudioDiffusion1D.py (:428-438, :523, :278-280)."""
    from uniaudio2_b200 import _lib

    L = _lib.lib()
    g = torch.Generator().manual_seed(4)
    B, T, C, Kin = 3, 375, 768, 1024
    x = torch.randn(B, T, Kin, generator=g)
    W = torch.randn(2 * C, Kin, generator=g) / Kin ** 0.5
    bias = torch.randn(2 * C, generator=g) * 0.1
    feat = torch.randn(B, T, C, generator=g)
    mask = torch.tensor([0, 1, 0], dtype=torch.uint8)
    params = F.linear(x, W, bias)
    dg, beta = params.chunk(2, dim=-1)
    gamma = 1.0 + 0.1 * dg.tanh()
    mk = mask.float().view(B, 1, 1)
    ref = (gamma * (1 - mk) + 1.0 * mk) * feat + (beta * (1 - mk) + 0.0 * mk)
    xd, Wd, bd, fd, md = x.cuda(), W.cuda(), bias.cuda(), feat.cuda(), mask.cuda()
    pd = torch.empty(B, T, 2 * C, device="cuda")
    out = torch.empty(B, T, C, device="cuda")
    _lib.check(L.ua2_linear_bias_f32(_lib.ptr(xd), _lib.ptr(Wd), _lib.ptr(bd), _lib.ptr(pd), B * T, 2 * C, Kin, None))
    _lib.check(L.ua2_film_f32(_lib.ptr(pd), _lib.ptr(fd), _lib.ptr(md), _lib.ptr(out), B, T, C, 0.1, None))
    torch.cuda.synchronize()
    assert float((pd.cpu() - params).abs().max()) < 2e-5 * float(params.abs().max())
    assert float((out.cpu() - ref).abs().max()) < 1e-4
    r = torch.randn(2, 768, 150, generator=g)
    ref_i = F.interpolate(r, scale_factor=2.5, mode="nearest")
    rd = r.cuda()
    yi = torch.empty(2, 768, ref_i.shape[-1], device="cuda")
    _lib.check(L.ua2_interp_nearest_f32(_lib.ptr(rd), _lib.ptr(yi), 2, 768, 150, ref_i.shape[-1], 2.5, None))
    torch.cuda.synchronize()
    assert torch.equal(yi.cpu(), ref_i)
