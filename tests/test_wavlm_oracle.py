"""CPU: oracle/wavlm_oracle.py pinned against transformers' WavLMModel (fixtures of tests/golden/frontend_golden.pt written by
oracle/make_golden_frontend.py from the real class, plus a live run when transformers is importable), and the product drop-in's
parameter surface against the real class's state dict."""
import os

import pytest
import torch

from oracle import wavlm_oracle as WO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(ROOT, "tests", "golden", "frontend_golden.pt"), weights_only=False)


@pytest.mark.parametrize("name", ["small", "mid"])
def test_wavlm_oracle_matches_transformers_fixture(gold, name):
    c = gold["wavlm_" + name]
    sd = WO.random_state_dict(c["cfg"], c["seed"])
    with torch.no_grad():
        hs = WO.hidden_states(sd, c["cfg"], c["wav16"])
    assert len(hs) == len(c["hidden_states"]) == c["cfg"]["num_hidden_layers"] + 1
    for a, b in zip(hs, c["hidden_states"]):
        assert a.shape == b.shape
        assert float((a - b).abs().max()) < 5e-6 * max(1.0, float(b.abs().max()))


def test_wavlm_oracle_matches_transformers_live():
    pytest.importorskip("transformers")
    from oracle.make_golden_frontend import SMALL, build_hf_wavlm

    m = build_hf_wavlm(SMALL, 77)  # strict load of the oracle's seeded dict: key set and shapes are the real class's
    sd = WO.random_state_dict(SMALL, 77)
    g = torch.Generator().manual_seed(3)
    wav = torch.randn(3, 2500, generator=g) * 0.5
    with torch.no_grad():
        ref = m(wav, output_hidden_states=True).hidden_states
        hs = WO.hidden_states(sd, SMALL, wav)
    for a, b in zip(hs, ref):
        assert float((a - b).abs().max()) < 5e-6 * max(1.0, float(b.abs().max()))


def test_get_wavlm_feature_restatement_shapes():
    """AudioDiffusion1D.get_wavlm_feature (:359-370) at the production frame geometry on a short clip."""
    cfg = dict(WO.BASE_PLUS, hidden_size=64, num_attention_heads=2, intermediate_size=64, num_hidden_layers=10, conv_dim=(16,) * 7,
               num_conv_pos_embeddings=16, num_conv_pos_embedding_groups=4)
    sd = WO.random_state_dict(cfg, 1)
    wav24 = torch.randn(1, 1, 24000)  # 1 s -> 16000 + 160 samples -> 50 frames
    with torch.no_grad():
        out = WO.get_wavlm_feature(sd, cfg, wav24, len_semantic=20)
    assert out.shape == (1, 64, 40)


def test_product_wavlm_parameter_surface():
    from uniaudio2_b200 import _lib
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.modeling_wavlm import WavLMConfig, WavLMModel
    from oracle.make_golden_frontend import MID

    sd = WO.random_state_dict(MID, 12)  # strict-loaded into transformers.WavLMModel by make_golden_frontend
    cfg = WavLMConfig(**{k: v for k, v in MID.items()})
    m = WavLMModel(cfg)
    have = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert have == {k: tuple(v.shape) for k, v in sd.items()}
    m.load_state_dict(sd, strict=True)
    assert WavLMModel(WavLMConfig()).num_frames(480160) == 1500  # 30 s + 160 zeros at the checkpoint geometry
    with pytest.raises(_lib.Ua2Error):  # no CPU fallback
        m(torch.zeros(1, 4000))
    with pytest.raises(NotImplementedError):
        WavLMModel(WavLMConfig(do_stable_layer_norm=True, feat_extract_norm="layer"))
