"""GPU: the device-side waveform front-end (csrc/ua2_frontend.cu behind tools/tokenizer/ReasoningCodec_film/frontend.py) against
outputs of the real torchaudio.transforms.Resample / transformers.WhisperFeatureExtractor (tests/golden/frontend_golden.pt) and
against the oracle (pinned bit-exactly to those classes by tests/test_frontend_oracle.py) on fresh inputs.

Bars (floating point): resampler 2e-6 of the signal scale (same fp32 FMA chain, different summation order than torch's conv1d);
log-mel features 2e-4 absolute on values in [-1.5, 2] - the reference's torch.stft is an fp32 FFT, the kernel a direct DFT accumulated
in float64, so the difference is the reference's own rounding, amplified by log10 only on bins near the max - 8 floor."""
import os

import pytest
import torch

from oracle import frontend_oracle as FO

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(ROOT, "tests", "golden", "frontend_golden.pt"), weights_only=False)


def _fe():
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film import frontend

    return frontend


def test_resampler_matches_torchaudio_fixture(gold):
    r = gold["resample"]
    down = _fe().Resample(24000, 16000)
    y = down(r["x"].cuda())
    assert y.shape == r["y_24k_16k"].shape and y.is_cuda
    assert float((y.cpu() - r["y_24k_16k"]).abs().max()) < 2e-6
    up = _fe().Resample(16000, 24000)(r["x"][:, :1000].contiguous().cuda())
    assert float((up.cpu() - r["y_16k_24k"]).abs().max()) < 2e-6
    # pad / cut in the same launch: zeros behind the valid samples, identical samples in front
    n = y.shape[-1]
    yp = down(r["x"].cuda(), pad_to=n + 160)
    assert torch.equal(yp[:, :n], y) and bool((yp[:, n:] == 0).all())
    assert torch.equal(down(r["x"].cuda(), pad_to=1000), y[:, :1000])
    # leading dimensions are kept, like torchaudio
    y3 = down(r["x"][:, None].cuda())
    assert y3.shape == (3, 1, n) and torch.equal(y3[:, 0], y)


@pytest.mark.parametrize("L", [1, 2, 5, 720240])
def test_resampler_fresh_inputs_vs_oracle(L):
    g = torch.Generator().manual_seed(L)
    x = torch.randn(2, L, generator=g) * 0.3
    ref = FO.resample(x, 24000, 16000)
    y = _fe().Resample(24000, 16000)(x.cuda()).cpu()
    assert y.shape == ref.shape
    assert float((y - ref).abs().max()) < 2e-6 * max(1.0, float(ref.abs().max()))


def test_log_mel_matches_whisper_feature_extractor_fixture(gold):
    m = gold["logmel"]
    out = _fe().WhisperLogMel()(m["wav16"].cuda(), sampling_rate=16000)["input_features"]
    assert out.shape == (2, 80, 3000) and out.is_cuda and out.dtype == torch.float32
    out = out.cpu()
    assert bool(torch.isfinite(out).all())
    assert float((out[:, :, ::7] - m["features_strided"]).abs().max()) < 2e-4
    assert float((out[:, :, :340] - m["features_head"]).abs().max()) < 2e-4


def test_log_mel_full_windows_vs_oracle():
    """The reference's batch: 6 windows of 30 s (+ the 160 extra samples that the extractor cuts off)."""
    g = torch.Generator().manual_seed(30)
    wav = torch.randn(6, 480160, generator=g) * 0.1
    wav[3] *= torch.linspace(0, 1, 480160)  # fade-in: quiet frames near the floor
    ref = FO.whisper_log_mel(wav)
    lm = _fe().WhisperLogMel()
    out = lm(wav.cuda())["input_features"].cpu()
    assert float((out - ref).abs().max()) < 2e-4
    assert torch.equal(lm(wav.cuda())["input_features"].cpu(), out)  # deterministic
    with pytest.raises(ValueError):
        lm(wav.cuda(), sampling_rate=24000)


def test_get_whisper_features_end_to_end():
    """ReasoningTokenizer.get_whisper_features (reason_tokenizer.py:67-72): 24 kHz windows of 30 s + 240 samples -> (B, 80, 3000)."""
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.reason_tokenizer import ReasoningTokenizer

    g = torch.Generator().manual_seed(31)
    audio = torch.randn(2, 720240, generator=g) * 0.2
    ref = FO.whisper_features(audio)
    tok = ReasoningTokenizer(None, None, device=torch.device("cuda:0"))
    out = tok.get_whisper_features(audio, 24000)
    assert out.shape == (2, 80, 3000) and out.is_cuda
    assert float((out.cpu() - ref).abs().max()) < 2e-4
    with pytest.raises(ValueError):
        tok.get_whisper_features(audio, 22050)
