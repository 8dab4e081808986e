"""GPU: the Whisper encoder drop-in (csrc/ua2_enc.cu behind tools/tokenizer/ReasoningCodec_film/models/modeling_whisper.py) against
the fixtures of the UNMODIFIED reference class source (tests/golden/whisper_golden.pt) and against the oracle on fresh inputs.

Bars (floating point): fp32 class (default: 3xTF32 linears, fp32 attention) 1e-4 of the output scale; bf16 mode (the reference's
autocast arithmetic: bf16 operands, fp32 accumulation, tensor-core attention) 3e-2 of the output scale and really different."""
import os

import pytest
import torch

from oracle import whisper_oracle as WO

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model(cfg, sd):
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.modeling_whisper import WhisperConfig, WhisperModel

    m = WhisperModel(WhisperConfig(d_model=cfg.d_model, encoder_attention_heads=cfg.encoder_attention_heads, encoder_ffn_dim=cfg.encoder_ffn_dim,
                                   encoder_layers=cfg.encoder_layers, max_source_positions=cfg.max_source_positions,
                                   num_mel_bins=cfg.num_mel_bins)).encoder
    assert sorted(m.state_dict().keys()) == sorted(WO.state_keys(cfg))
    m.load_state_dict(sd, strict=True)
    return m.to("cuda:0")


def _rel(a, b):
    return float((a - b).abs().max()) / max(1.0, float(b.abs().max()))


def test_whisper_encoder_matches_reference_golden():
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "whisper_golden.pt"), weights_only=False)
    for name, c in gold["cases"].items():
        cfg = WO.WhisperCfg(**c["cfg"])
        sd = WO.random_state_dict(cfg, c["param_seed"])
        g = torch.Generator().manual_seed(c["input_seed"])
        mel = torch.randn(c["batch"], cfg.num_mel_bins, 2 * cfg.max_source_positions, generator=g)
        m = _model(cfg, sd)
        y = m(mel.cuda(), return_dict=True).last_hidden_state.cpu()
        assert y.shape == c["out"].shape
        assert bool(torch.isfinite(y).all())
        assert _rel(y, c["out"]) < 1e-4, (name, _rel(y, c["out"]))


@pytest.mark.parametrize("d,heads,ffn,layers,P,B", [(128, 2, 512, 2, 100, 3), (256, 4, 1024, 2, 250, 2), (64, 2, 128, 1, 40, 1), (256, 2, 256, 1, 64, 1)])
def test_whisper_encoder_fresh_inputs_both_modes(d, heads, ffn, layers, P, B):
    cfg = WO.WhisperCfg(d_model=d, encoder_attention_heads=heads, encoder_ffn_dim=ffn, encoder_layers=layers, max_source_positions=P)
    sd = WO.random_state_dict(cfg, seed=d + P)
    g = torch.Generator().manual_seed(P)
    mel = torch.randn(B, cfg.num_mel_bins, 2 * P, generator=g)
    with torch.no_grad():
        ref = WO.WhisperEncoderOracle(cfg, sd).forward(mel)
    m = _model(cfg, sd)
    y32 = m(mel.cuda()).last_hidden_state.cpu()
    assert _rel(y32, ref) < 1e-4, _rel(y32, ref)
    if d // heads == 64:  # the tensor-core attention serves head size 64 (every Whisper checkpoint)
        m.set_option("bf16", 1)
        y16 = m(mel.cuda()).last_hidden_state.cpu()
        m.set_option("bf16", 0)
        y32b = m(mel.cuda()).last_hidden_state.cpu()
        assert torch.equal(y32, y32b)  # the option leaves the default path untouched, and the default path is deterministic
        err = _rel(y16, ref)
        assert 1e-6 < err < 3e-2, err
    else:
        m.set_option("bf16", 1)
        with pytest.raises(Exception):
            m(mel.cuda())
        m.set_option("bf16", 0)


def test_whisper_encoder_interface_errors():
    from uniaudio2_b200 import _lib

    cfg = WO.WhisperCfg(d_model=64, encoder_attention_heads=1, encoder_ffn_dim=128, encoder_layers=1, max_source_positions=40)
    sd = WO.random_state_dict(cfg, seed=1)
    m = _model(cfg, sd)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 80, 79, device="cuda"))  # the reference fails at inputs_embeds + embed_pos (modeling_whisper.py:811)
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 80, 80, device="cuda"), output_hidden_states=True)
    bad = dict(sd)
    bad["layers.0.fc1.weight"] = torch.zeros(3, 3)
    with pytest.raises(Exception):
        m.load_state_dict(bad, strict=True)
    with pytest.raises(_lib.Ua2Error):
        type(m)(m.config)(torch.zeros(1, 80, 80))  # CPU parameters: no fallback


def test_get_whisper_feature_mirrors_the_reference_cut():
    """AudioDiffusion1D.get_whisper_feature (:334-343): frames = max(int(n / 24000 * 50), 2 * len_semantic), result (B, D, T)."""
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.AudioDiffusion1D import AudioDiffusion1D
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.transformer_1d_flow import Transformer1DModel

    cfg = WO.WhisperCfg(d_model=128, encoder_attention_heads=2, encoder_ffn_dim=256, encoder_layers=1, max_source_positions=100)
    sd = WO.random_state_dict(cfg, seed=3)
    enc = _model(cfg, sd)
    est = Transformer1DModel(num_attention_heads=1, attention_head_dim=64, in_channels=56, out_channels=12, num_layers=1, norm_type="ada_norm_single",
                             activation_fn="gelu-approximate", attention_bias=True, norm_elementwise_affine=False, num_positional_embeddings=64)
    m = AudioDiffusion1D(est, codec_dim=32, codebook_size=16, codebook_dim=8, sq_codec_latent=12, whisper_dim=128, wavlm_dim=32, bestrq_dim=32)
    import pytest as _pt
    with _pt.raises(Exception):
        m.get_whisper_feature(torch.zeros(1, 80, 200), 24000, 10)
    m.attach_whisper_encoder(enc)
    g = torch.Generator().manual_seed(9)
    mel = torch.randn(2, 80, 200, generator=g)
    with torch.no_grad():
        ref = WO.WhisperEncoderOracle(cfg, sd).forward(mel)
    for n_samples, len_sem, want in ((24000, 10, 50), (24000, 40, 80), (12345, 1, 25)):
        y = m.get_whisper_feature(mel.cuda(), n_samples, len_sem)
        assert y.shape == (2, 128, want)
        assert _rel(y.cpu(), ref[:, :want].transpose(1, 2)) < 1e-4
