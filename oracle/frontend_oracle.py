"""TEST INFRASTRUCTURE ONLY (imported by tests/ and oracle/make_golden_frontend.py; never by uniaudio2_b200/).

CPU restatement of the waveform front-end of ReasoningCodec_film's tokenize direction ("move the CPU
log-mel on-device"):

    tools/tokenizer/ReasoningCodec_film/reason_tokenizer.py
        :37      self.transfer16k = torchaudio.transforms.Resample(24000, 16000)
        :36      self.wav_processor = WhisperFeatureExtractor.from_pretrained(whisper_path)
        :67-72   get_whisper_features: transfer16k(audio) -> wav_processor(audio.cpu().numpy(), sampling_rate=16000)["input_features"]
    tools/tokenizer/ReasoningCodec_film/models/AudioDiffusion1D.py
        :227, :363-367   wavlm_transfer = Resample(24000, 16000); 160 zero samples appended before the WavLM encoder

Both leaves are third-party code that is absent from /root/reference: torchaudio (unpinned in pyproject.toml) and
transformers==4.57.0 (pyproject.toml:25; this image carries 5.5.0).  Both ARE installed in this image, so the restatement below
is pinned against the real classes: tests/test_frontend_oracle.py runs torchaudio.transforms.Resample and
transformers.WhisperFeatureExtractor live next to it, and oracle/make_golden_frontend.py stores their outputs as fixtures
(tests/golden/frontend_golden.pt) for the GPU tests.

  * sinc_resample_kernel / resample: torchaudio.functional.functional._get_sinc_resample_kernel / _apply_sinc_resample_kernel
    (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99 - the defaults of transforms.Resample)
  * mel_filter_bank: transformers.audio_utils.mel_filter_bank(norm="slaney", mel_scale="slaney") as WhisperFeatureExtractor.__init__
    calls it (201 bins, 80 filters, 0 - 8000 Hz, 16 kHz)
  * whisper_log_mel: WhisperFeatureExtractor._torch_extract_fbank_features (the path taken whenever torch is importable) after
    __call__'s pad / truncate to 30 s
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

N_FFT = 400
HOP = 160
N_MELS = 80
SAMPLING_RATE = 16000
N_SAMPLES = 30 * SAMPLING_RATE      # WhisperFeatureExtractor.n_samples
N_FRAMES = N_SAMPLES // HOP         # nb_max_frames = 3000


def sinc_resample_kernel(orig_freq, new_freq, lowpass_filter_width=6, rolloff=0.99):
    """(new, 1, 2 * width + orig) fp32 polyphase filters and `width`, frequencies already divided by their gcd."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = torch.arange(-width, width + orig, dtype=torch.float64)[None, None] / orig
    t = torch.arange(0, -new, -1)[:, None, None] / new + idx  # (dtype=None in torchaudio: the phase term p / new is rounded to fp32)
    t = (t * base).clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    kern = torch.where(t == 0, torch.tensor(1.0, dtype=torch.float64), t.sin() / t) * window * (base / orig)
    return kern.to(torch.float32), width, orig, new


def resample(wav, orig_freq, new_freq):
    """wav (..., L) fp32 -> (..., ceil(new * L / orig))."""
    kern, width, orig, new = sinc_resample_kernel(orig_freq, new_freq)
    shape = wav.shape
    x = wav.reshape(-1, shape[-1])
    L = x.shape[-1]
    y = F.conv1d(F.pad(x, (width, width + orig))[:, None], kern, stride=orig)
    y = y.transpose(1, 2).reshape(x.shape[0], -1)
    n = int(math.ceil(new * L / orig))
    return y[..., :n].reshape(shape[:-1] + (n,))


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    mels = 3.0 * f / 200.0
    logstep = 27.0 / np.log(6.4)
    return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-300) / 1000.0) * logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    logstep = np.log(6.4) / 27.0
    return np.where(m >= 15.0, 1000.0 * np.exp(logstep * (m - 15.0)), 200.0 * m / 3.0)


def mel_filter_bank(n_bins=1 + N_FFT // 2, n_mels=N_MELS, fmin=0.0, fmax=8000.0, sr=SAMPLING_RATE):
    """(n_bins, n_mels) float64, Slaney scale + Slaney area normalisation."""
    mel_pts = np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2)
    hz_pts = _mel_to_hz(mel_pts)
    fft_freqs = np.linspace(0, sr // 2, n_bins)
    diff = np.diff(hz_pts)
    slopes = hz_pts[None, :] - fft_freqs[:, None]
    down = -slopes[:, :-2] / diff[:-1]
    up = slopes[:, 2:] / diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    fb *= (2.0 / (hz_pts[2:n_mels + 2] - hz_pts[:n_mels]))[None, :]
    return fb


def pad_or_trim(wav16, n=N_SAMPLES):
    """WhisperFeatureExtractor.__call__: truncation=True, padding='max_length' with zeros."""
    L = wav16.shape[-1]
    return wav16[..., :n] if L >= n else F.pad(wav16, (0, n - L))


def whisper_log_mel(wav16):
    """wav16 (B, L) fp32 at 16 kHz -> (B, 80, 3000) fp32 input_features."""
    x = pad_or_trim(wav16.to(torch.float32))
    stft = torch.stft(x, N_FFT, HOP, window=torch.hann_window(N_FFT), return_complex=True)
    mag = stft[..., :-1].abs() ** 2
    mel = torch.from_numpy(mel_filter_bank()).to(torch.float32).T @ mag
    log_spec = torch.clamp(mel, min=1e-10).log10()
    mx = log_spec.max(dim=2, keepdim=True)[0].max(dim=1, keepdim=True)[0]
    log_spec = torch.maximum(log_spec, mx - 8.0)
    return (log_spec + 4.0) / 4.0


def whisper_features(audio_24k):
    """reason_tokenizer.py:67-72 for sr = 24000: (B, L) -> (B, 80, 3000)."""
    return whisper_log_mel(resample(audio_24k, 24000, 16000))


def log_mel_direct_f64(wav16, n_frames=None):
    """The same features from the formula the device kernel evaluates (no FFT): reflect-centred frames, hann window, a direct DFT
    with float64 accumulation, float64 mel projection.  Used by the CPU-shim test to separate 'formula' from 'kernel' errors."""
    x = pad_or_trim(wav16.to(torch.float32)).double()
    B, L = x.shape
    n_frames = L // HOP if n_frames is None else n_frames
    xp = F.pad(x[:, None], (N_FFT // 2, N_FFT // 2), mode="reflect")[:, 0]
    frames = xp.unfold(1, N_FFT, HOP)[:, :n_frames]                      # (B, n_frames, 400)
    win = torch.hann_window(N_FFT, dtype=torch.float32).double()
    n = torch.arange(N_FFT, dtype=torch.float64)
    k = torch.arange(N_FFT // 2 + 1, dtype=torch.float64)
    ang = 2 * math.pi * ((k[:, None] * n[None, :]) % N_FFT) / N_FFT
    fw = frames * win
    re = fw @ torch.cos(ang).T
    im = fw @ torch.sin(ang).T
    power = (re * re + im * im).float()                                   # (B, n_frames, 201)
    mel = power.double() @ torch.from_numpy(mel_filter_bank()).float().double()
    log_spec = torch.log10(torch.clamp(mel.float(), min=1e-10).double()).float().transpose(1, 2)
    mx = log_spec.amax(dim=(1, 2), keepdim=True)
    return (torch.maximum(log_spec, mx - 8.0) + 4.0) / 4.0
