"""CPU: the host logic of the tokenize direction - ReasoningTokenizer.audio2token (reason_tokenizer.py:85-129: periodic extension, windows
of 30 s + 240 samples, batches of 6, cut back to the clip's own code lengths) - of the product against the UNMODIFIED reference method
source executed on a stand-in self, with the same scripted collaborators on both sides (a feature extractor that records its input and
a model whose codes are a function of the window contents): identical tokens AND the identical sequence of calls."""
import types

import pytest
import torch

from oracle.make_golden_detok import extract_method


def _collaborators(log):
    def get_whisper_features(audio, sr):
        log.append(("mel", tuple(audio.shape), sr, float(audio.double().sum())))
        return torch.full((audio.shape[0], 80, 3000), float(audio.shape[0]))

    class Model:
        def fetch_codes_batch(self, audios, mels, additional_feats=None, return_reasoning_text=False):
            log.append(("codes", tuple(audios.shape), tuple(mels.shape), additional_feats, return_reasoning_text))
            B = audios.shape[0]
            key = (audios[:, 0, ::997].double().sum(-1) * 1000).round().long()  # codes depend on the window's samples
            reason = (key[:, None, None] + torch.arange(150)[None, :, None] * 8 + torch.arange(8)[None, None, :]) % 4096
            rec = (key[:, None, None] * 3 + torch.arange(375)[None, :, None] * 8 + torch.arange(8)[None, None, :]) % 8192
            return [reason], [rec], [None]

    return get_whisper_features, Model()


def _reference_tokenizer(log):
    ns = {"torch": torch}
    fn = extract_method("reason_tokenizer.py", "ReasoningTokenizer", "audio2token", ns)
    feats, model = _collaborators(log)
    self_ = types.SimpleNamespace(device=torch.device("cpu"), sample_rate=24000, rec_frame_rate=12.5, reason_frame_rate=5, model=model,
                                  get_whisper_features=feats)
    return lambda *a, **kw: fn(self_, *a, **kw)


def _product_tokenizer(log):
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.reason_tokenizer import ReasoningTokenizer

    feats, model = _collaborators(log)
    tok = ReasoningTokenizer(model, None, device=torch.device("cpu"))
    tok.get_whisper_features = feats
    return tok.audio2token


@pytest.mark.parametrize("seconds,channels,ndim", [(3.94, 1, 2), (30.0, 1, 2), (30.02, 1, 3), (70.5, 1, 2), (200.0, 1, 2), (10.0, 2, 2)])
def test_audio2token_host_logic_matches_reference_source(seconds, channels, ndim):
    g = torch.Generator().manual_seed(int(seconds * 100))
    wav = torch.randn(channels, int(seconds * 24000), generator=g) * 0.1
    if ndim == 3:
        wav = wav[None]
    log_ref, log_mine = [], []
    with torch.autocast(device_type="cpu", enabled=False):
        # the reference enters torch.autocast(device_type="cuda") around fetch_codes_batch: harmless on a CPU-only build (a warning)
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            r_ref, s_ref = _reference_tokenizer(log_ref)(wav, 24000)
    r, s = _product_tokenizer(log_mine)(wav, 24000)
    assert r.shape == r_ref.shape and s.shape == s_ref.shape
    assert torch.equal(r, r_ref) and torch.equal(s, s_ref)
    assert log_mine == log_ref and len(log_ref) >= 2
    if channels == 1:
        n = wav.shape[-1]
        assert r.shape == (1, 8, int(n / 24000 * 5) + 1) and s.shape == (1, 8, int(n / 24000 * 12.5) + 1)


def test_tokenize_surface():
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.reason_tokenizer import ReasoningTokenizer

    tok = ReasoningTokenizer(None, None, device=torch.device("cpu"))
    t = torch.zeros(8, 5, dtype=torch.long)
    assert tok.tokenize(t) is t                      # a tensor passes through (reason_tokenizer.py:388-389)
    with pytest.raises(NotImplementedError):
        tok.tokenize(3)
    with pytest.raises(NotImplementedError):
        tok.audio2token(torch.zeros(1, 24000), 24000, return_reasoning_text=True)
    with pytest.raises(AssertionError):
        tok.audio2token(torch.zeros(24000), 24000)
