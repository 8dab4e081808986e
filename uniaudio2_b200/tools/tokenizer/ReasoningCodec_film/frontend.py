"""Device-side waveform front-end of the tokenize direction of tools/tokenizer/ReasoningCodec_film:

    self.transfer16k = torchaudio.transforms.Resample(24000, 16000).to(device)          # reason_tokenizer.py:37
    self.wav_processor = WhisperFeatureExtractor.from_pretrained(whisper_path)           # reason_tokenizer.py:36
    spectrogram = self.wav_processor(audio.detach().cpu().numpy(), sampling_rate=16000,
                                     return_tensors="pt")["input_features"]              # reason_tokenizer.py:71  (B, 80, 3000)
    self.wavlm_transfer = torchaudio.transforms.Resample(24000, 16000)                   # AudioDiffusion1D.py:227

`Resample` and `WhisperLogMel` are drop-ins for those two objects on CUDA tensors: the reference copies the batch to the host for the
feature extractor and back; here resampling, STFT, mel projection and normalisation run in libua2_b200.so (csrc/ua2_frontend.cu)
and the features never leave the device.  The small host-side tables (polyphase filter, hann window, Slaney mel filters) are built
here in float64 from the published formulas of torchaudio / transformers.audio_utils and uploaded once.  No torch / CPU fallback.
"""
import math

import numpy as np
import torch

from .... import _lib


def _polyphase_filters(orig_freq, new_freq, lowpass_filter_width=6, rolloff=0.99):
    """torchaudio's `sinc_interp_hann` filter bank: `new` windowed-sinc filters of 2 * width + orig taps, frequencies reduced by their
    gcd.  Filter p, tap k weighs input sample orig * m + k - width for output new * m + p."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    cutoff = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / cutoff)
    taps = np.arange(-width, width + orig, dtype=np.float64) / orig
    phase = (-np.arange(new, dtype=np.float32) / np.float32(new)).astype(np.float64)  # torchaudio evaluates p / new in fp32
    t = np.clip((phase[:, None] + taps[None, :]) * cutoff, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        sinc = np.where(t == 0, 1.0, np.sin(t) / t)
    return (sinc * window * (cutoff / orig)).astype(np.float32), width, orig, new


class Resample:
    """torchaudio.transforms.Resample(orig_freq, new_freq) for CUDA tensors (..., L) -> (..., ceil(new * L / orig)).
    `__call__(x, pad_to=n)` additionally zero-pads / cuts the result to n samples in the same launch (the 30 s window of the
    Whisper feature extractor; the 160 zeros AudioDiffusion1D.get_wavlm_feature appends)."""

    def __init__(self, orig_freq=24000, new_freq=16000):
        self.orig_freq, self.new_freq = int(orig_freq), int(new_freq)
        self._filters, self.width, self._orig, self._new = _polyphase_filters(orig_freq, new_freq)
        self._dev = {}

    def to(self, device):
        return self

    def _table(self, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = torch.from_numpy(self._filters).to(device).contiguous()
        return self._dev[key]

    def out_length(self, L):
        return int(math.ceil(self._new * L / self._orig))

    @torch.inference_mode()
    def __call__(self, waveform, pad_to=None):
        if waveform.device.type != "cuda":
            raise _lib.Ua2Error("uniaudio2_b200 Resample runs on CUDA tensors only (no CPU fallback)")
        if not waveform.is_floating_point():
            raise TypeError(f"Expected floating point type for waveform tensor, but received {waveform.dtype}.")  # torchaudio's message
        shape = waveform.shape
        x = waveform.to(torch.float32).reshape(-1, shape[-1]).contiguous()
        B, L = x.shape
        n = self.out_length(L)
        n_store = n if pad_to is None else int(pad_to)
        y = torch.empty(B, n_store, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().ua2_resample_f32(_lib.ptr(x), L, _lib.ptr(self._table(x.device)), _lib.ptr(y), n_store, B, L, min(n, n_store), n_store,
                                                   self._orig, self._new, self.width, _lib.current_stream()), "ua2_resample_f32")
        return y.reshape(shape[:-1] + (n_store,))


def _slaney_mel_filters(n_bins, n_mels, fmin, fmax, sampling_rate):
    """transformers.audio_utils.mel_filter_bank(norm='slaney', mel_scale='slaney'): triangular filters on the Slaney mel scale
    (linear below 1 kHz, logarithmic above), each divided by its band width.  (n_bins, n_mels) float64."""
    log_step = math.log(6.4) / 27.0

    def to_mel(f):
        return 15.0 + math.log(f / 1000.0) / log_step if f >= 1000.0 else 3.0 * f / 200.0

    edges_mel = np.linspace(to_mel(fmin), to_mel(fmax), n_mels + 2)
    edges_hz = np.where(edges_mel >= 15.0, 1000.0 * np.exp(log_step * (edges_mel - 15.0)), 200.0 * edges_mel / 3.0)
    bins_hz = np.linspace(0, sampling_rate // 2, n_bins)
    rise = (bins_hz[:, None] - edges_hz[None, :-2]) / (edges_hz[1:-1] - edges_hz[:-2])[None, :]
    fall = (edges_hz[None, 2:] - bins_hz[:, None]) / (edges_hz[2:] - edges_hz[1:-1])[None, :]
    tri = np.maximum(0.0, np.minimum(rise, fall))
    return tri * (2.0 / (edges_hz[2:] - edges_hz[:-2]))[None, :]


class WhisperLogMel:
    """`WhisperFeatureExtractor(...)(audio, sampling_rate=16000, return_tensors='pt')['input_features']` for CUDA tensors:
    (B, L) at 16 kHz -> (B, feature_size, 3000) fp32 on the same device.  Pass `wav_processor` (a real WhisperFeatureExtractor) to
    take n_fft / hop_length / mel_filters / n_samples from a checkpoint's preprocessor_config; the defaults are every Whisper
    checkpoint's values except large-v3's 128 mel bins."""

    def __init__(self, wav_processor=None, feature_size=80, sampling_rate=16000, hop_length=160, chunk_length=30, n_fft=400):
        if wav_processor is not None:
            feature_size, sampling_rate = wav_processor.feature_size, wav_processor.sampling_rate
            hop_length, n_fft, chunk_length = wav_processor.hop_length, wav_processor.n_fft, wav_processor.chunk_length
            if getattr(wav_processor, "dither", 0.0) != 0.0:
                raise NotImplementedError("dither != 0 is not served (no Whisper checkpoint sets it)")
        self.feature_size, self.sampling_rate, self.hop_length, self.n_fft = feature_size, sampling_rate, hop_length, n_fft
        self.n_samples = chunk_length * sampling_rate
        self.nb_max_frames = self.n_samples // hop_length
        filters = wav_processor.mel_filters if wav_processor is not None else _slaney_mel_filters(1 + n_fft // 2, feature_size, 0.0, 8000.0, sampling_rate)
        self.mel_filters = np.asarray(filters, dtype=np.float64)
        self._dev = {}

    def _tables(self, device):
        key = str(device)
        if key not in self._dev:
            window = torch.hann_window(self.n_fft)  # the reference's window tensor itself (periodic hann, evaluated by torch in fp32)
            self._dev[key] = (window.to(device), torch.from_numpy(self.mel_filters).to(torch.float32).to(device).contiguous())
        return self._dev[key]

    @torch.inference_mode()
    def __call__(self, audio, sampling_rate=None, return_tensors="pt"):
        if sampling_rate is not None and sampling_rate != self.sampling_rate:
            raise ValueError(f"The model corresponding to this feature extractor was trained using a sampling rate of {self.sampling_rate}, "
                             f"got {sampling_rate}.")  # the reference extractor's check
        if not isinstance(audio, torch.Tensor) or audio.device.type != "cuda":
            raise _lib.Ua2Error("uniaudio2_b200 WhisperLogMel takes CUDA tensors only (no CPU fallback)")
        x = audio.to(torch.float32)
        if x.dim() == 1:
            x = x[None]
        B, L = x.shape
        if L != self.n_samples:  # padding='max_length', truncation=True
            x = torch.nn.functional.pad(x, (0, self.n_samples - L)) if L < self.n_samples else x[:, :self.n_samples]
        x = x.contiguous()
        window, filters = self._tables(x.device)
        out = torch.empty(B, self.feature_size, self.nb_max_frames, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().ua2_whisper_logmel_f32(_lib.ptr(x), x.stride(0), _lib.ptr(window), _lib.ptr(filters), _lib.ptr(out), B, self.n_samples,
                                                         self.n_fft, self.hop_length, self.feature_size, self.nb_max_frames, _lib.current_stream()),
                       "ua2_whisper_logmel_f32")
        return {"input_features": out}
