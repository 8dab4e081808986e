"""GPU parity suite for the flow-matching decoder through the C ABI (ua2_dit_*) behind the
reference's interface (uniaudio2_b200...models.transformer_1d_flow.Transformer1DModel, ...models.AudioDiffusion1D.BASECFM).

Checked against the committed fixtures produced by the UNMODIFIED in-repo reference code (tests/golden/dit_golden.pt) and
against the CPU oracle on fresh seeded inputs.  Bar: 1e-4 max-abs relative to the tensor's scale (fp32-class arithmetic;
linears of >= 32 rows run as 3xTF32 tensor-core GEMMs, the rest as fp32 FMA chains)."""
import os

import pytest
import torch

from oracle import dit_oracle as DO
from oracle.make_golden_dit import dit_cfgs

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-4


@pytest.fixture(scope="module")
def dit_golden():
    return torch.load(os.path.join(ROOT, "tests", "golden", "dit_golden.pt"), weights_only=False)


def _rel(a, b):
    return float((a - b).abs().max()) / max(1.0, float(b.abs().max()))


def _product(cfg, sd):
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.transformer_1d_flow import Transformer1DModel

    m = Transformer1DModel(**cfg.ctor_kwargs())
    m.load_state_dict(sd, strict=True)
    return m.to("cuda:0")


@pytest.mark.parametrize("name", ["tiny", "mid"])
def test_estimator_matches_reference_golden(dit_golden, name):
    cfg = dit_cfgs()[name]
    m = _product(cfg, DO.random_state_dict(cfg, seed=909))
    for i, c in enumerate(dit_golden[name]["cases"]):
        y = m(c["x"].cuda(), timestep=c["t"].cuda(), added_cond_kwargs={"resolution": None, "aspect_ratio": None}).sample
        assert y.shape == c["y"].shape
        assert _rel(y.cpu(), c["y"]) < TOL, f"{name}: case {i} (B, T) = {tuple(c['x'].shape[:2])}"
    assert m.last_launch_count() > 0


@pytest.mark.parametrize("name", ["tiny", "mid"])
def test_euler_solver_matches_reference_golden(dit_golden, name):
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.AudioDiffusion1D import BASECFM

    cfg = dit_cfgs()[name]
    cfm = BASECFM(_product(cfg, DO.random_state_dict(cfg, seed=909)))
    for i, s in enumerate(dit_golden[name]["solves"]):
        t_span = torch.linspace(0, 1, s["steps"] + 1)
        out = cfm.solve_euler(s["z"].cuda(), s["incontext"].cuda(), s["incontext_length"], t_span, s["mu"].cuda(), None,
                              s["guidance_scale"])
        assert _rel(out.cpu(), s["out"]) < TOL, f"{name}: solve {i}"
    with pytest.raises(ValueError):
        cfm.solve_euler(s["z"].cuda(), s["incontext"].cuda(), 0, t_span, s["mu"].cuda(), None, 1.0)
    with pytest.raises(RuntimeError):
        cfm.solve_euler(s["z"].repeat(2, 1, 1).cuda(), s["incontext"].cuda(), 0, t_span, s["mu"].cuda(), None, 1.5)


@pytest.mark.parametrize("heads,hd,T,B", [(3, 64, 150, 2), (2, 128, 70, 1), (4, 32, 45, 3)])
def test_estimator_equals_oracle_many_rows_and_head_sizes(heads, hd, T, B):
    """>= 128 rows (tensor-core GEMMs for every linear), ragged last query / key tiles, head sizes 32 / 64 / 128."""
    cfg = DO.DitCfg(num_attention_heads=heads, attention_head_dim=hd, in_channels=56, out_channels=12, num_layers=2,
                    num_positional_embeddings=160)
    sd = DO.random_state_dict(cfg, seed=heads * 10 + hd)
    m = _product(cfg, sd)
    g = torch.Generator().manual_seed(T)
    x = torch.randn(B, T, cfg.in_channels, generator=g)
    t = torch.rand(B, generator=g)
    with torch.no_grad():
        ref = DO.DitOracle(cfg, sd).forward(x, t)
    y = m(x.cuda(), timestep=t.cuda()).sample
    assert _rel(y.cpu(), ref) < TOL


def test_solver_with_zero_estimator_property():
    """Size-independent property at a longer sequence: zero weights -> the solution keeps the noise outside the in-context
    rows and is the reference's blend inside them."""
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.AudioDiffusion1D import BASECFM

    cfg = DO.DitCfg(num_attention_heads=2, attention_head_dim=64, in_channels=40, out_channels=8, num_layers=1,
                    num_positional_embeddings=600)
    sd = {k: (v if k == "pos_embed.pe" else torch.zeros_like(v)) for k, v in DO.random_state_dict(cfg, seed=1).items()}
    cfm = BASECFM(_product(cfg, sd))
    g = torch.Generator().manual_seed(0)
    T, ic = 500, 125
    z, inc, mu = torch.randn(1, T, 8, generator=g), torch.randn(1, T, 8, generator=g), torch.randn(1, T, 24, generator=g)
    t_span = torch.linspace(0, 1, 11)
    out = cfm.solve_euler(z.cuda(), inc.cuda(), ic, t_span, mu.cuda(), None, 1.5).cpu()
    assert torch.equal(out[:, ic:], z[:, ic:])
    t_last = t_span[-2]
    assert torch.allclose(out[:, :ic], (1 - (1 - 1e-4) * t_last) * z[:, :ic] + t_last * inc[:, :ic], atol=1e-6)
