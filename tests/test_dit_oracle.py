"""CPU suite for the flow-matching decoder: oracle/dit_oracle.py reproduces the committed
outputs of the UNMODIFIED in-repo Transformer1DModel / BasicTransformerBlock / BASECFM.solve_euler bit-exactly
(tests/golden/dit_golden.pt, written by oracle/make_golden_dit.py); the product module keeps the reference's state-dict
keys; the C ABI validates without a GPU."""
import ctypes as C
import os

import pytest
import torch

from oracle import dit_oracle as DO
from oracle.make_golden_dit import THREADS, dit_cfgs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(autouse=True)
def _generator_thread_count():
    n = torch.get_num_threads()
    torch.set_num_threads(THREADS)  # bit-exact comparisons run with the generator's thread count (ATen partitions sums by it)
    yield
    torch.set_num_threads(n)


@pytest.fixture(scope="module")
def dit_golden():
    return torch.load(os.path.join(ROOT, "tests", "golden", "dit_golden.pt"), weights_only=False)


@pytest.mark.parametrize("name", ["tiny", "mid"])
def test_oracle_matches_reference(dit_golden, name):
    cfg = dit_cfgs()[name]
    sd = DO.random_state_dict(cfg, seed=909)
    assert {k: float(v.double().sum()) for k, v in sd.items()} == dit_golden[f"__checksum_{name}"]
    orc = DO.DitOracle(cfg, sd)
    with torch.no_grad():
        for c in dit_golden[name]["cases"]:
            assert torch.equal(orc.forward(c["x"], c["t"]), c["y"])
        for s in dit_golden[name]["solves"]:
            t_span = torch.linspace(0, 1, s["steps"] + 1)
            out = orc.solve_euler(s["z"].clone(), s["incontext"], s["incontext_length"], t_span, s["mu"], s["guidance_scale"])
            assert torch.equal(out, s["out"])
        with pytest.raises(ValueError):
            orc.solve_euler(s["z"].clone(), s["incontext"], 0, t_span, s["mu"], 1.0)


def test_solver_properties():
    """Size-independent checks of the solver restatement: with a zero estimator the solution keeps the noise outside the
    in-context rows and lands on (1 - (1 - sigma) t_last) noise + t_last * incontext inside them."""
    cfg = dit_cfgs()["tiny"]
    sd = {k: torch.zeros_like(v) for k, v in DO.random_state_dict(cfg, seed=1).items()}
    orc = DO.DitOracle(cfg, sd)
    g = torch.Generator().manual_seed(0)
    z, ic, mu = torch.randn(1, 9, 8, generator=g), torch.randn(1, 9, 8, generator=g), torch.randn(1, 9, 24, generator=g)
    t_span = torch.linspace(0, 1, 6)
    with torch.no_grad():
        out = orc.solve_euler(z.clone(), ic, 4, t_span, mu, 1.5)
    assert torch.equal(out[:, 4:], z[:, 4:])
    t_last = t_span[-2]
    assert torch.allclose(out[:, :4], (1 - (1 - 1e-4) * t_last) * z[:, :4] + t_last * ic[:, :4], atol=1e-6)


def test_product_state_dict_and_refusals():
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.AudioDiffusion1D import BASECFM
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.transformer_1d_flow import Transformer1DModel

    for name, cfg in dit_cfgs().items():
        m = Transformer1DModel(**cfg.ctor_kwargs())
        assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == DO.state_dict_shapes(cfg), name
        sd = DO.random_state_dict(cfg, seed=909)
        m.load_state_dict(sd, strict=True)
        assert torch.equal(m.pos_embed.pe, sd["pos_embed.pe"])  # the table the constructor builds = the reference buffer
        with pytest.raises(Exception):  # CPU parameters: refuse, no fallback
            m(torch.zeros(1, 4, cfg.in_channels), timestep=torch.zeros(1))
        with pytest.raises(Exception):
            BASECFM(m).solve_euler(torch.zeros(1, 4, cfg.out_channels), torch.zeros(1, 4, cfg.out_channels), 0, torch.linspace(0, 1, 3),
                                   torch.zeros(1, 4, cfg.in_channels - 2 * cfg.out_channels), None, 1.5)
    with pytest.raises(NotImplementedError):
        Transformer1DModel(num_attention_heads=2, attention_head_dim=64, in_channels=40, out_channels=8)  # geglu / layer_norm defaults


def test_production_config_file_matches_the_served_configuration():
    """models/model_config.json of the reference, restated (the file itself is not read at test time)."""
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.transformer_1d_flow import Transformer1DModel

    prod = dict(activation_fn="gelu-approximate", attention_bias=True, attention_head_dim=64, attention_type="default",
                cross_attention_dim=None, double_self_attention=False, dropout=0.0, in_channels=1040, norm_elementwise_affine=False,
                norm_eps=1e-06, norm_num_groups=32, norm_type="ada_norm_single", num_attention_heads=24, num_embeds_ada_norm=1000,
                num_layers=1, num_vector_embeds=None, only_cross_attention=False, out_channels=136, patch_size=1, sample_size=384,
                upcast_attention=False, use_linear_projection=False, _class_name="Transformer1DModel", _diffusers_version="0.22.0.dev0")
    m = Transformer1DModel(**prod)  # one layer instead of 32 keeps the CPU test light
    assert m.proj_in.ffn_1.weight.shape == (1536, 1040, 3) and m.proj_out.ffn_2.weight.shape == (136, 136)


def test_dit_cabi_validation():
    from uniaudio2_b200 import _lib

    L = _lib.lib()
    h = C.c_void_p()
    for bad in (_lib.DitCfg(2, 48, 40, 8, 1, 64, 512, 1e-6), _lib.DitCfg(2, 64, 42, 8, 1, 64, 512, 1e-6),
                _lib.DitCfg(2, 64, 40, 8, 0, 64, 512, 1e-6)):
        with pytest.raises(ValueError):
            _lib.check(L.ua2_dit_create(C.byref(bad), C.byref(h)))
    good = _lib.DitCfg(2, 64, 40, 8, 1, 64, 512, 1e-6)
    assert L.ua2_dit_create(C.byref(good), C.byref(h)) == 0
    s3 = (C.c_int64 * 3)(128, 40, 3)
    assert L.ua2_dit_load_weight(h, b"proj_in.ffn_1.weight", C.c_void_p(256), s3, 3) == 0
    with pytest.raises(ValueError):  # wrong shape
        _lib.check(L.ua2_dit_load_weight(h, b"proj_out.ffn_1.weight", C.c_void_p(256), s3, 3))
    s2 = (C.c_int64 * 2)(128, 128)
    assert L.ua2_dit_load_weight(h, b"transformer_blocks.0.attn1.to_q.weight", C.c_void_p(256), s2, 2) == 0
    with pytest.raises(ValueError):  # block index out of range
        _lib.check(L.ua2_dit_load_weight(h, b"transformer_blocks.1.attn1.to_q.weight", C.c_void_p(256), s2, 2))
    with pytest.raises(ValueError):
        _lib.check(L.ua2_dit_load_weight(h, b"transformer_blocks.0.attn2.to_q.weight", C.c_void_p(256), s2, 2))
    with pytest.raises(ValueError):  # parameters missing
        _lib.check(L.ua2_dit_finalize(h, None))
    with pytest.raises(ValueError):  # not finalized
        _lib.check(L.ua2_dit_forward(h, C.c_void_p(256), C.c_void_p(256), C.c_void_p(256), 1, 4, None))
    assert L.ua2_dit_destroy(h) == 0
