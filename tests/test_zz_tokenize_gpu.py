"""GPU: the tokenize direction of ReasoningCodec_film assembled - AudioDiffusion1D.fetch_codes_batch (models/AudioDiffusion1D.py:492-551):
Whisper encoder, WavLM encoder (device resampler + 160 zeros, hidden states 6..9), reasoning encoder + its residual VQ, and the
own-code chain (strided convolutions, fusion, FiLM, three residual VQs) in the reference's order, with the BEST-RQ features coming from a
provider the caller attaches (the one front-end this package does not build).  Every stage has its own parity suite; this one checks the
composition: shapes and frame bookkeeping of a 2 s window at the checkpoint widths (1024 / 768 / 1024 / 768), equality with the stages
called one by one, and agreement with the CPU oracles chained the same way."""
import pytest
import torch

from oracle import encode_oracle as EO
from oracle import thinking_oracle as TO
from oracle import wavlm_oracle as WLO
from oracle import whisper_oracle as WO

pytestmark = pytest.mark.gpu

B, SAMPLES, TB = 2, 48000, 50  # 2 s at 24 kHz: 100 Whisper / WavLM frames, 50 BEST-RQ frames, 10 query tokens, 25 code frames
W_CFG = WO.WhisperCfg(d_model=1024, encoder_attention_heads=16, encoder_ffn_dim=256, encoder_layers=1, max_source_positions=100)
L_CFG = dict(WLO.BASE_PLUS, intermediate_size=256, num_hidden_layers=9, conv_dim=(64,) * 7)
T_CFG = dict(TO.CFG, depth=1)


def _build(seed):
    from test_zz_encode_gpu import _product
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.audio_thinking import AudioThinking
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.modeling_wavlm import WavLMConfig, WavLMModel
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.modeling_whisper import WhisperConfig, WhisperModel

    p = EO.random_params(seed)
    m = _product(p)
    w_sd, l_sd, t_sd = WO.random_state_dict(W_CFG, seed + 1), WLO.random_state_dict(L_CFG, seed + 2), TO.random_state_dict(T_CFG, seed + 3)
    enc = WhisperModel(WhisperConfig(d_model=W_CFG.d_model, encoder_attention_heads=W_CFG.encoder_attention_heads, encoder_ffn_dim=W_CFG.encoder_ffn_dim,
                                     encoder_layers=W_CFG.encoder_layers, max_source_positions=W_CFG.max_source_positions)).encoder
    enc.load_state_dict(w_sd, strict=True)
    m.attach_whisper_encoder(enc.cuda())
    wl = WavLMModel(WavLMConfig(**L_CFG))
    wl.load_state_dict(l_sd, strict=True)
    m.attach_wavlm_encoder(wl.cuda())
    at = AudioThinking(dim=T_CFG["dim"], interval=T_CFG["interval"], encoder_depth=T_CFG["depth"], whisper_fea_dim=T_CFG["whisper_dim"], mu_dim=T_CFG["mu_dim"])
    g = torch.Generator().manual_seed(seed + 4)
    vq = {"project_in.weight": torch.randn(64, 768, generator=g) / 768 ** 0.5, "project_in.bias": 0.1 * torch.randn(64, generator=g),
          "project_out.weight": torch.randn(768, 64, generator=g) / 8, "project_out.bias": 0.1 * torch.randn(768, generator=g)}
    books = torch.stack([torch.randn(4096, 64, generator=g) * (0.7 ** i) for i in range(8)])
    for i in range(8):
        vq[f"layers.{i}._codebook.embed"] = books[i:i + 1].clone()
    at.load_state_dict({**t_sd, **{"reasoning_vq." + k: v for k, v in vq.items()}}, strict=True)
    m.attach_audio_thinking(at.cuda())
    p = dict(p)
    p.update({"reasoning_vq.project_in.weight": vq["project_in.weight"], "reasoning_vq.project_in.bias": vq["project_in.bias"],
              "reasoning_vq.project_out.weight": vq["project_out.weight"], "reasoning_vq.project_out.bias": vq["project_out.bias"],
              "reasoning_vq.codebooks": books})
    acoustic, semantic = torch.randn(B, 1024, TB, generator=g), torch.randn(B, 1024, TB, generator=g)

    class BestRQ:  # stand-in provider: the conformer is not part of this package
        def extract_continous_embeds_multiple(self, audios):
            assert tuple(audios.shape) == (B, 1, SAMPLES)
            return acoustic.cuda(), semantic.cuda()

    m.attach_bestrq(BestRQ())
    return m, p, (w_sd, l_sd, t_sd), (acoustic, semantic)


def test_fetch_codes_batch_composition():
    m, p, (w_sd, l_sd, t_sd), (acoustic, semantic) = _build(50)
    g = torch.Generator().manual_seed(60)
    audio = torch.randn(B, 1, SAMPLES, generator=g) * 0.2
    mels = torch.randn(B, 80, 200, generator=g)
    masks = [torch.tensor([0, 1], dtype=torch.uint8), torch.tensor([0, 0], dtype=torch.uint8), torch.tensor([1, 0], dtype=torch.uint8)]
    reason, rec, merge = m.fetch_codes_batch(audio.cuda(), mels.cuda(), additional_feats=[], return_reasoning_text=False, film_masks=masks)
    reason, rec, merge = reason[0], rec[0], merge[0]
    assert reason.shape == (B, 10, 8) and rec.shape == (B, 25, 8) and merge.shape == (B, 25, 768)
    assert reason.dtype == torch.int64 and rec.dtype == torch.int64 and bool(torch.isfinite(merge).all())
    # ---- the same stages one by one: identical results
    whisper = m.get_whisper_feature(mels.cuda(), SAMPLES, TB)
    wavlm = m.get_wavlm_feature(audio.cuda(), TB)
    assert whisper.shape == (B, 1024, 100) and wavlm.shape == (B, 768, 100)
    qr, rcodes, _ = m.encode_reasoning_part(whisper, semantic.cuda())
    codes2, merge2 = m.fetch_codes_from_features(whisper, wavlm, acoustic.cuda(), semantic.cuda(), qr, film_masks=masks)
    assert torch.equal(rcodes, reason) and torch.equal(codes2, rec) and torch.equal(merge2, merge)
    # ---- the CPU oracles chained the same way
    with torch.no_grad():
        o_whisper = WO.WhisperEncoderOracle(W_CFG, w_sd).forward(mels)[:, :100].transpose(1, 2)
        o_wavlm = WLO.get_wavlm_feature(l_sd, L_CFG, audio, TB)
        o_query = TO.encode(t_sd, T_CFG, o_whisper, semantic)
        o_qr, o_rcodes = EO.residual_vq_forward(o_query, p, "reasoning_vq", 8)
        o_codes, o_merge = EO.fetch_codes_from_features(p, whisper=o_whisper, wavlm=o_wavlm, bestrq_acoustic=acoustic, bestrq_semantic=semantic,
                                                        quantized_reasoning=o_qr, film_masks=[mk.float() for mk in masks])

    def rel(a, b):
        return float((a - b).abs().max()) / max(1.0, float(b.abs().max()))

    assert rel(whisper.cpu(), o_whisper) < 2e-4 and rel(wavlm.cpu(), o_wavlm) < 2e-4
    # codes are argmins over 4096 / 8192 entries of features that carry ~1e-4 of arithmetic difference: near-ties may resolve differently
    assert float((reason.cpu() == o_rcodes).float().mean()) >= 0.95, float((reason.cpu() == o_rcodes).float().mean())
    assert float((rec.cpu() == o_codes).float().mean()) >= 0.90, float((rec.cpu() == o_codes).float().mean())


def test_fetch_codes_batch_refusals():
    from uniaudio2_b200 import _lib

    m, *_ = _build(51)
    audio, mels = torch.zeros(B, 1, SAMPLES, device="cuda"), torch.zeros(B, 80, 200, device="cuda")
    with pytest.raises(NotImplementedError):
        m.fetch_codes_batch(audio, mels, return_reasoning_text=True)
    m.pretrained_model = None
    with pytest.raises(_lib.Ua2Error):
        m.fetch_codes_batch(audio, mels)
