"""CPU suite for the Moshi-family streaming transformer + sampler:
  * oracle/moshi_oracle.py reproduces the committed reference outputs (tests/golden/moshi_golden.pt, written by
    oracle/make_golden_moshi.py from the UNMODIFIED llm_modules/transformer.py + llm_utils/sampling.py) bit-exactly;
  * the reference's own known-answer test for the sampler (llm_utils/sampling.py:156-174) holds for the restatement;
  * the product module keeps the reference's state-dict keys, and the C ABI validates configurations without a GPU."""
import ctypes as C
import os

import pytest
import torch

from oracle import moshi_oracle as MO
from oracle.make_golden_moshi import stx_cfgs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(autouse=True)
def _generator_thread_count():
    """ATen's CPU matmul / SDPA partition work by thread count, which changes the last bit of some sums: bit-exact
    comparisons against the fixtures run with the thread count oracle/make_golden_moshi.py used."""
    n = torch.get_num_threads()
    torch.set_num_threads(4)
    yield
    torch.set_num_threads(n)


@pytest.fixture(scope="module")
def moshi_golden():
    return torch.load(os.path.join(ROOT, "tests", "golden", "moshi_golden.pt"), weights_only=False)


@pytest.mark.parametrize("name", ["mimi_like", "lm_like", "dep_like", "sin_like"])
def test_oracle_matches_reference(moshi_golden, name):
    cfg = stx_cfgs()[name]
    fx = moshi_golden[name]
    sd = MO.random_state_dict(cfg, seed=2025)
    assert {k: float(v.double().sum()) for k, v in sd.items()} == moshi_golden[f"__checksum_{name}"]
    orc = MO.StxOracle(cfg, sd)
    with torch.no_grad():
        assert torch.equal(orc.forward(fx["x_nonstream"]), fx["y_nonstream"])
        orc.start_streaming(fx["batch"])
        n = len(fx["schedule"])
        for i, (x, y) in enumerate(zip(fx["xs"], fx["ys"])):
            if i == n:
                orc.reset_streaming()
            assert x.shape[1] == fx["schedule"][i % n]
            assert torch.equal(orc.forward(x), y), f"{name}: streaming call {i}"
        assert torch.equal(orc.state["kv"][-1].cache, fx["last_cache"]) and orc.state["kv"][-1].end_offset == fx["last_end"]
        orc.stop_streaming()
    with pytest.raises(ValueError):
        orc.reset_streaming()


def test_ring_positions_properties():
    """RingKVCache.complete position recovery (transformer.py:254-276), including the `delta <= 0` quirk."""
    cap = 5
    assert MO.ring_positions(cap, 0).tolist() == [-1] * 5
    assert MO.ring_positions(cap, 3).tolist() == [0, 1, 2, -1, -1]
    assert MO.ring_positions(cap, 5).tolist() == [5, 1, 2, 3, 4]  # slot 0 reports the FUTURE position 5: oldest key hidden
    assert MO.ring_positions(cap, 7).tolist() == [5, 6, 7, 3, 4]
    for E in range(cap, 40):
        pos = MO.ring_positions(cap, E)
        live = sorted(p for p in pos.tolist() if p < E)
        assert live == list(range(E - cap + 1, E))  # cap - 1 usable keys once wrapped


def test_sampler_oracle_matches_reference(moshi_golden):
    n = 0
    for key, fx in moshi_golden.items():
        if not key.startswith("sampler_"):
            continue
        kw = dict(fx["kwargs"])
        tok = MO.sample_token(fx["logits"], q=fx["q"], **kw)
        assert torch.equal(tok, fx["tokens"]), key
        n += 1
    assert n >= 9


def test_multinomial_frequency_kat():
    """The reference's own self-test, llm_utils/sampling.py:156-174."""
    torch.manual_seed(1234)
    ps = torch.tensor([5.0, 2.0, 12.0, 6.0, 8.0, 1.0, 0.0, 4.0])
    cnts = torch.zeros(ps.shape, dtype=torch.long)
    for _ in range(1000):
        cnts[MO.multinomial(ps)] += 1
    diff = cnts / cnts.sum() - ps / ps.sum()
    assert diff.abs().max().item() < 1.5e-2


def test_product_state_dict_keys_match_reference_layout():
    from uniaudio2_b200.llm_modules.transformer import StreamingTransformer

    for name, cfg in stx_cfgs().items():
        m = StreamingTransformer(d_model=cfg.d_model, num_heads=cfg.num_heads, num_layers=cfg.num_layers,
                                 dim_feedforward=cfg.dim_feedforward, causal=cfg.causal, context=cfg.context,
                                 positional_embedding=cfg.positional_embedding, max_period=cfg.max_period,
                                 positional_scale=cfg.positional_scale, norm=cfg.norm, layer_scale=cfg.layer_scale,
                                 gating=cfg.gating, weights_per_step=cfg.weights_per_step)
        shapes = MO.state_dict_shapes(cfg)
        got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        assert got == shapes, name
        m.load_state_dict(MO.random_state_dict(cfg, seed=1), strict=True)
        with pytest.raises(Exception):  # parameters on the CPU: the product must refuse, not fall back
            m(torch.zeros(1, 1, cfg.d_model))
    with pytest.raises(ValueError):
        StreamingTransformer(64, 2, 1, norm="batch_norm")


def test_sampler_product_refuses_cpu_tensors():
    from uniaudio2_b200.llm_utils.sampling import sample_token

    with pytest.raises(Exception):
        sample_token(torch.zeros(2, 8))


def test_stx_cabi_validation():
    from uniaudio2_b200 import _lib

    L = _lib.lib()

    def cfg(**kw):
        base = dict(d_model=128, num_heads=4, num_layers=1, causal=1, context=8, positional_embedding=2, norm=0, gating=0,
                    weights_per_step=0, layer_scale=0, max_period=10000.0, positional_scale=1.0)
        base.update(kw)
        ff = (C.c_int32 * 64)(*([base.pop("ff", 256)] + [0] * 63))
        return _lib.StxCfg(base["d_model"], base["num_heads"], base["num_layers"], base["causal"], base["context"],
                           base["positional_embedding"], base["norm"], base["gating"], base["weights_per_step"],
                           base["layer_scale"], ff, base["max_period"], base["positional_scale"])

    h = C.c_void_p()
    for bad in (cfg(num_heads=3), cfg(num_heads=8), cfg(norm=7), cfg(gating=0, weights_per_step=4), cfg(gating=1, ff=100)):
        with pytest.raises(ValueError):
            _lib.check(L.ua2_stx_create(C.byref(bad), C.byref(h)))
    good = cfg()
    assert L.ua2_stx_create(C.byref(good), C.byref(h)) == 0
    shape = (C.c_int64 * 2)(384, 128)
    assert L.ua2_stx_load_weight(h, b"layers.0.self_attn.in_proj_weight", C.c_void_p(256), shape, 2) == 0
    with pytest.raises(ValueError):  # wrong shape
        _lib.check(L.ua2_stx_load_weight(h, b"layers.0.self_attn.out_proj.weight", C.c_void_p(256), shape, 2))
    with pytest.raises(ValueError):  # unknown key / layer out of range
        _lib.check(L.ua2_stx_load_weight(h, b"layers.3.norm1.weight", C.c_void_p(256), shape, 2))
    with pytest.raises(ValueError):  # rms key on a layer_norm configuration
        _lib.check(L.ua2_stx_load_weight(h, b"layers.0.norm1.alpha", C.c_void_p(256), shape, 2))
    with pytest.raises(ValueError):  # parameters missing
        _lib.check(L.ua2_stx_finalize(h))
    with pytest.raises(ValueError):  # not finalized
        _lib.check(L.ua2_stx_start_streaming(h, 1, None))
    with pytest.raises(ValueError):  # streaming.py:118-121
        _lib.check(L.ua2_stx_reset_streaming(h))
    assert L.ua2_stx_destroy(h) == 0
    # sampler argument checks (no launch happens)
    with pytest.raises(ValueError):
        _lib.check(L.ua2_sample_token_f32(C.c_void_p(256), 1, 16, 1, 1.0, 32, 0.0, -1, None, 0, 0, C.c_void_p(256), None))
    with pytest.raises(ValueError):
        _lib.check(L.ua2_sample_token_f32(C.c_void_p(256), 1, 8192, 1, 1.0, 0, 0.5, -1, None, 0, 0, C.c_void_p(256), None))
