"""CPU: oracle/frontend_oracle.py (resampler + Whisper log-mel restatements) pinned against the real third-party classes the
reference calls (torchaudio.transforms.Resample, transformers.WhisperFeatureExtractor - run live here and stored in
tests/golden/frontend_golden.pt by oracle/make_golden_frontend.py), the product's host-side tables against the oracle's, and the
host-only bucket function of the C ABI against the WavLM formula."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import frontend_oracle as FO
from oracle import wavlm_oracle as WO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(ROOT, "tests", "golden", "frontend_golden.pt"), weights_only=False)


def test_resampler_oracle_matches_torchaudio_fixture(gold):
    r = gold["resample"]
    assert torch.equal(FO.resample(r["x"], 24000, 16000), r["y_24k_16k"])
    assert torch.equal(FO.resample(r["x"][:, :1000].contiguous(), 16000, 24000), r["y_16k_24k"])


def test_resampler_oracle_matches_torchaudio_live():
    torchaudio = pytest.importorskip("torchaudio")
    g = torch.Generator().manual_seed(5)
    for L in (1, 2, 3, 100, 7201):
        x = torch.randn(2, 1, L, generator=g)
        assert torch.equal(FO.resample(x, 24000, 16000), torchaudio.transforms.Resample(24000, 16000)(x))
    k, width, orig, new = FO.sinc_resample_kernel(24000, 16000)
    assert (width, orig, new, tuple(k.shape)) == (10, 3, 2, (2, 1, 23))
    assert torch.equal(k, torchaudio.transforms.Resample(24000, 16000).kernel)


def test_log_mel_oracle_matches_whisper_feature_extractor(gold):
    m = gold["logmel"]
    feats = FO.whisper_log_mel(m["wav16"])
    assert feats.shape == (2, FO.N_MELS, FO.N_FRAMES)
    assert torch.equal(feats[:, :, ::7], m["features_strided"])
    assert torch.equal(feats[:, :, :340], m["features_head"])
    # the float64-accumulated direct DFT the device kernel evaluates differs from torch's fp32 FFT by that FFT's own rounding only
    assert float((FO.log_mel_direct_f64(m["wav16"])[:, :, :340] - m["features_head"]).abs().max()) < 2e-4


def test_log_mel_oracle_matches_whisper_feature_extractor_live():
    transformers = pytest.importorskip("transformers")
    fe = transformers.WhisperFeatureExtractor()
    assert float(np.abs(fe.mel_filters - FO.mel_filter_bank()).max()) == 0.0
    g = torch.Generator().manual_seed(6)
    wav = torch.randn(2, 480160, generator=g) * 0.05  # longer than 30 s: truncated, like the reference's 30 s + 240-sample windows
    ref = fe(wav.numpy(), sampling_rate=16000, return_tensors="pt")["input_features"]
    assert torch.equal(FO.whisper_log_mel(wav), ref)


def test_product_host_tables_match_the_oracle():
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film import frontend as FE

    for orig, new in ((24000, 16000), (16000, 24000), (48000, 16000)):
        f, width, o, n = FE._polyphase_filters(orig, new)
        k, w2, o2, n2 = FO.sinc_resample_kernel(orig, new)
        assert (width, o, n) == (w2, o2, n2) and np.array_equal(f, k.reshape(n, -1).numpy())
    assert np.array_equal(FE._slaney_mel_filters(201, 80, 0.0, 8000.0, 16000), FO.mel_filter_bank())
    lm = FE.WhisperLogMel()
    assert (lm.n_samples, lm.nb_max_frames, lm.feature_size) == (FO.N_SAMPLES, FO.N_FRAMES, FO.N_MELS)
    with pytest.raises(Exception):  # no CPU fallback
        lm(torch.zeros(1, 16000))
    with pytest.raises(Exception):
        FE.Resample(24000, 16000)(torch.zeros(1, 100))


def test_c_abi_bucket_table_matches_the_wavlm_formula():
    from uniaudio2_b200 import _lib

    lib = _lib.lib()
    for T, nb, md in ((1, 320, 800), (7, 320, 800), (1500, 320, 800), (3001, 320, 800), (900, 32, 128)):
        out = (C.c_int32 * (2 * T - 1))()
        assert lib.ua2_wavlm_rel_bucket_table(T, nb, md, out) == 0
        want = WO.relative_positions_bucket(torch.arange(-(T - 1), T), nb, md)
        assert torch.equal(torch.tensor(list(out), dtype=torch.long), want)
    assert lib.ua2_wavlm_rel_bucket_table(4, 6, 800, (C.c_int32 * 7)()) != 0  # bad geometry is refused
