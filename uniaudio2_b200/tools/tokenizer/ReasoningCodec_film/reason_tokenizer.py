"""Drop-in for the detokenize surface of the reference's tools/tokenizer/ReasoningCodec_film/reason_tokenizer.py::ReasoningTokenizer:

    wave = tok.detokenize_no_reason(rec_codec (8, T2), return_reasoning_text=False, steps=10)        # :399-404
    wave = tok.token2audio_no_reason(rec_codec (B, 8, T2), False, duration=20, num_steps=20)         # :228-306

i.e. the codes -> waveform caller of `--stage all` (multi_task_inference.py:546).  What the reference does, and this does too:
the code sequence is made periodic and cut into windows of `duration` seconds that advance by 3/4 of a window; every window is
one AudioDiffusion1D.inference_codes call (a flow-matching solve) whose first latent frames are pinned to the tail of the previous
window's latents; each window's latents go through the SQ-codec decoder, and consecutive waveforms are joined by a linear
cross-fade over the shared quarter (on the host, in float64, like the reference).  tests/test_detok_oracle.py checks this host
logic bit-exactly against fixtures produced by the unmodified reference source; `model` and `SQCodec` are the uniaudio2_b200
drop-ins (GPU only).  Of the tokenize direction this class carries `get_whisper_features` (:67-72), `audio2token`
(:85-129) and `tokenize` (:377-391); the encoders hang off AudioDiffusion1D (Whisper, WavLM, AudioThinking; BEST-RQ features come from a
provider the caller attaches).  tests/test_tokenize_host_cpu.py checks the windowing bit-exactly against the unmodified reference source.
"""
import math
from dataclasses import dataclass

import numpy as np
import torch

LATENT_DIM = 136          # SQ-codec latent width, hard-wired at reason_tokenizer.py:234
GUIDANCE_IN_LOOP = 1.5    # the reference ignores its guidance_scale argument inside the window loop (:273, :282)


@dataclass
class _Windows:
    """Window geometry in code frames (12.5 Hz) and in waveform samples (24 kHz); the reference's `*_samples` variables."""

    codes: int          # code frames per window
    hop: int            # window advance, 3/4 of a window (integer arithmetic of the reference: codes // 4 * 3)
    overlap: int        # codes - hop
    latents: int        # latent frames per window (25 Hz)
    wave: int           # samples per window
    wave_overlap: int   # samples shared by consecutive windows

    @staticmethod
    def of(duration, rec_frame_rate, sq_codec_hz, sample_rate):
        codes = int(duration * rec_frame_rate)
        hop = codes // 4 * 3
        wave = int(duration * sample_rate)
        return _Windows(codes, hop, codes - hop, int(duration * sq_codec_hz), wave, wave - wave // 4 * 3)


def _make_periodic(rec_codec, w: _Windows):
    """Repeat the code sequence (by doubling, like the reference) until it covers a whole number of hops plus one overlap."""
    def tile_to(x, n):
        while x.shape[-1] < n:
            x = torch.cat([x, x], -1)
        return x[:, :, :n]

    if rec_codec.shape[-1] < w.codes:
        rec_codec = tile_to(rec_codec, w.codes)
    n = rec_codec.shape[-1]
    if (n - w.overlap) % w.hop > 0:
        rec_codec = tile_to(rec_codec, math.ceil((n - w.overlap) / float(w.hop)) * w.hop + w.overlap)
    return rec_codec


def _cross_fade(joined, nxt, n):
    """Blend the last n samples of `joined` into the first n of `nxt` with a linear ramp (float64 ramp from numpy, as the
    reference builds it) and append the rest."""
    ramp = torch.from_numpy(np.linspace(0, 1, n)[None, :])
    joined[:, -n:] = joined[:, -n:] * (1 - ramp) + nxt[:, :n] * ramp
    return torch.cat([joined, nxt[:, n:]], -1)


class ReasoningTokenizer:
    def __init__(self, model, SQCodec, device=torch.device("cuda")):
        self.sample_rate = 24000
        self.device = device
        self.n_codebook = 8
        self.sq_codec_hz = 25        # latent frames per second
        self.rec_frame_rate = 12.5   # reconstruction-code frames per second
        self.reason_frame_rate = 5
        self.model = model
        self.SQCodec = SQCodec
        # The reference runs the window loop under torch.autocast(device_type='cuda', dtype=torch.bfloat16) (reason_tokenizer.py:265):
        # the flow decoder's linears see bf16 operands with fp32 accumulation.  True mirrors that (the estimator's "bf16" option: the
        # hand-written tcgen05 kind::f16 mainloop); False keeps the fp32-class 3xTF32 arithmetic that the 1e-4 parity tests compare
        # with the fp32 CPU oracle.
        self.autocast_bf16 = True

    @torch.inference_mode()
    def get_whisper_features(self, audio, sr):
        """reason_tokenizer.py:67-72: audio (B, samples) -> Whisper input features (B, 80, 3000).  The reference resamples on the device,
        copies the batch to the host for WhisperFeatureExtractor and copies the features back; here the resampler (with the 30 s pad /
        cut), the STFT, the mel projection and the normalisation are three launches of csrc/ua2_frontend.cu and nothing leaves the device."""
        from .frontend import Resample, WhisperLogMel

        if getattr(self, "wav_processor", None) is None:
            self.wav_processor = WhisperLogMel()
            self.transfer16k = Resample(24000, 16000)
        audio = audio.to(self.device)
        if sr != 16000:
            if sr != 24000:
                raise ValueError(f"sample rate {sr}: the reference's transfer16k resamples from 24000 Hz only")
            audio = self.transfer16k(audio, pad_to=self.wav_processor.n_samples)
        return self.wav_processor(audio, sampling_rate=16000, return_tensors="pt")["input_features"]

    @torch.no_grad()
    def audio2token(self, orig_samples, sr, return_reasoning_text=False, task_name="speech_reasoning", min_duration=30, batch_size=6):
        """reason_tokenizer.py:85-129: (channels, samples) at 24 kHz -> (reason codes (1, 8, Tr), reconstruction codes (1, 8, Ts)).
        The clip is made periodic (doubled until it covers one window, then once more), cut into windows of min_duration s + 240 samples
        and encoded `batch_size` windows at a time by AudioDiffusion1D.fetch_codes_batch; the code sequences are cut back to the clip's
        own length (12.5 and 5 frames per second, plus one)."""
        if return_reasoning_text:
            raise NotImplementedError("the reasoning-text head is not on this path")
        if orig_samples.ndim == 2:
            audios = orig_samples.unsqueeze(0).to(self.device)
        elif orig_samples.ndim == 3:
            audios = orig_samples.to(self.device)
        else:
            raise AssertionError(tuple(orig_samples.shape))  # the reference asserts ndim in (2, 3)
        audios = audios.squeeze(0)
        n_orig = audios.shape[-1]
        window = int(min_duration * self.sample_rate) + 240  # 240 extra samples so that the last frame of a window is complete
        n_rec = int(n_orig / float(self.sample_rate) * self.rec_frame_rate) + 1
        n_reason = int(n_orig / float(self.sample_rate) * self.reason_frame_rate) + 1
        while audios.shape[-1] < window:
            audios = torch.cat([audios, audios], -1)
        n_windows = audios.shape[-1] // (window - 240) + 1
        audios = torch.cat([audios, audios], -1)[:, :int(n_windows * window)]
        windows = audios.reshape(1, -1, window).permute(1, 0, 2).reshape(-1, 1, window)
        # the reference encodes under torch.autocast(device_type='cuda', dtype=torch.bfloat16) (:114-118): the two big encoders follow it
        # through their "bf16" option (bf16 operands, fp32 accumulation, tensor-core attention) when autocast_bf16 is set
        for enc in (getattr(self.model, "whisper_encoder", None), getattr(self.model, "wavlm_encoder", None)):
            if enc is not None and hasattr(enc, "set_option"):
                enc.set_option("bf16", 1 if self.autocast_bf16 else 0)
        reason, rec = [], []
        for i in range(0, windows.shape[0], batch_size):
            chunk = windows[i:i + batch_size]
            mels = self.get_whisper_features(chunk[:, 0, :], 24000).to(self.device)
            reasoning_codes, rec_codes, _ = self.model.fetch_codes_batch(chunk, mels, additional_feats=[], return_reasoning_text=False)
            reason.append(torch.cat(reasoning_codes, 1))
            rec.append(torch.cat(rec_codes, 1))
        reason = torch.cat(reason, 0).reshape(-1, 8).unsqueeze(0)[:, :n_reason, :].transpose(1, 2)
        rec = torch.cat(rec, 0).reshape(-1, 8).unsqueeze(0)[:, :n_rec, :].transpose(1, 2)
        return reason, rec

    def tokenize(self, wav, return_reasoning_text=False, task_name="asr", min_duration=30):
        """reason_tokenizer.py:377-391: a path -> (reason codes (8, Tr), reconstruction codes (8, Ts)); a tensor is returned as it is."""
        if isinstance(wav, str):
            import torchaudio  # file decoding only

            prompt_audio, fs = torchaudio.load(wav)
            if prompt_audio.shape[0] == 2:
                prompt_audio = prompt_audio.mean(0, keepdim=True)
            if fs != self.sample_rate:
                prompt_audio = torchaudio.functional.resample(prompt_audio, fs, self.sample_rate)
                fs = self.sample_rate
            reason_codec, rec_codec = self.audio2token(prompt_audio, fs, return_reasoning_text, task_name=task_name)
            return reason_codec.squeeze(0), rec_codec.squeeze(0)
        if isinstance(wav, torch.Tensor):
            return wav
        raise NotImplementedError

    def _randn(self, *shape):
        """Noise the reference draws on the CPU generator and then moves to the device (reason_tokenizer.py:234, :279)."""
        return torch.randn(*shape)

    def _solve_window(self, codes, prior, n_pinned, latent_frames, num_steps, disable_progress):
        return self.model.inference_codes([codes], None, prior, latent_frames, n_pinned, additional_feats=[], guidance_scale=GUIDANCE_IN_LOOP,
                                          num_steps=num_steps, disable_progress=disable_progress, scenario="other_seg")

    @torch.no_grad()
    def token2audio_no_reason(self, rec_codec, return_reasoning_text, duration=20, guidance_scale=1.5, num_steps=20, disable_progress=False):
        w = _Windows.of(duration, self.rec_frame_rate, self.sq_codec_hz, self.sample_rate)
        rec_codec = rec_codec.to(self.device)
        n_out = int(rec_codec.shape[-1] / 12.5 * self.sample_rate)  # samples the caller gets back: the ORIGINAL code length
        prior = self._randn(rec_codec.shape[0], int(duration * 25), LATENT_DIM).to(self.device)  # drawn before the codes are tiled
        rec_codec = _make_periodic(rec_codec, w)
        # ---- latents, window by window; from the second window on the first overlap/2 latent frames continue the previous window
        carried = w.overlap // 2
        latents = []
        estimator = getattr(getattr(self.model, "cfm_wrapper", None), "estimator", None)
        if estimator is not None and hasattr(estimator, "set_option"):
            estimator.set_option("bf16", 1 if self.autocast_bf16 else 0)
        for start in range(0, rec_codec.shape[-1] - w.hop, w.hop):
            n_pinned = 0
            if latents:
                tail = latents[-1][:, -carried:, :]
                n_pinned = tail.shape[1]
                fresh = self._randn(tail.shape[0], w.latents - n_pinned, tail.shape[-1]).to(self.device)
                prior = torch.cat([tail, fresh], 1)
            latents.append(self._solve_window(rec_codec[:, :, start:start + w.codes], prior, n_pinned, w.latents, num_steps, disable_progress))
        # ---- waveforms, joined on the host
        joined = None
        for lat in latents:
            wave = self.SQCodec.decode(lat.float().transpose(1, 2)).squeeze(0)[:, :w.wave].detach().cpu()
            joined = wave if joined is None else _cross_fade(joined, wave, w.wave_overlap)
        return joined[:, :n_out]

    def detokenize_no_reason(self, rec_codec, return_reasoning_text, min_duration=30, steps=50, guidance_scale=1.5, disable_progress=False):
        """rec_codec (8, T2) -> wave (1, samples); 20 s windows (the default of token2audio_no_reason)."""
        return self.token2audio_no_reason(rec_codec.unsqueeze(0), return_reasoning_text=return_reasoning_text, guidance_scale=guidance_scale,
                                          num_steps=steps, disable_progress=disable_progress)
