"""CPU: the restated Whisper encoder (oracle/whisper_oracle.py) reproduces the fixtures that
oracle/make_golden_whisper.py produced by executing the UNMODIFIED class source of the reference's modeling_whisper.py
(WhisperEncoder / WhisperEncoderLayer / WhisperAttention, :220-443, :723-867) - bit-equal."""
import os

import pytest
import torch

from oracle import whisper_oracle as WO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(autouse=True)
def _generator_thread_count():
    n = torch.get_num_threads()
    torch.set_num_threads(4)  # ATen partitions sums by thread count; the generator ran with 4
    yield
    torch.set_num_threads(n)


def test_whisper_encoder_oracle_matches_reference_golden():
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "whisper_golden.pt"), weights_only=False)
    assert len(gold["cases"]) >= 2
    for name, c in gold["cases"].items():
        cfg = WO.WhisperCfg(**c["cfg"])
        sd = WO.random_state_dict(cfg, c["param_seed"])
        g = torch.Generator().manual_seed(c["input_seed"])
        mel = torch.randn(c["batch"], cfg.num_mel_bins, 2 * cfg.max_source_positions, generator=g)
        with torch.no_grad():
            got = WO.WhisperEncoderOracle(cfg, sd).forward(mel)
        assert torch.equal(got, c["out"]), name


def test_state_keys_are_the_reference_modules():
    cfg = WO.WhisperCfg(d_model=64, encoder_attention_heads=1, encoder_ffn_dim=128, encoder_layers=2, max_source_positions=20)
    keys = WO.state_keys(cfg)
    assert "layers.1.self_attn.k_proj.weight" in keys and "layers.1.self_attn.k_proj.bias" not in keys  # modeling_whisper.py:240
    assert set(keys) == set(WO.state_shapes(cfg))
