"""Drop-in for the detokenize surface of the reference's tools/tokenizer/ReasoningCodec_film/reason_tokenizer.py::ReasoningTokenizer:

    wave = tok.detokenize_no_reason(rec_codec (8, T2), return_reasoning_text=False, steps=10)        # :399-404
    wave = tok.token2audio_no_reason(rec_codec (B, 8, T2), False, duration=20, num_steps=20)         # :228-306

i.e. the codes -> waveform caller of `--stage all` (multi_task_inference.py:546): windows of `duration` seconds with a 3/4 hop,
every window = AudioDiffusion1D.inference_codes (flow-matching solve with the previous window's tail as in-context latents) ->
ScalarModel.decode, linear cross-fade of the overlaps on the host (the reference does the cross-fade on the CPU in float64 too).
The host logic below is the reference's, statement for statement (tests/test_detok_oracle.py runs it against fixtures produced
by the unmodified reference source); `model` and `SQCodec` are the uniaudio2_b200 drop-ins (GPU only).
The tokenize direction (Whisper / WavLM / BEST-RQ front-ends) is not on this path.
"""
import math

import numpy as np
import torch


class ReasoningTokenizer:
    def __init__(self, model, SQCodec, device=torch.device("cuda")):
        self.sample_rate = 24000
        self.device = device
        self.n_codebook = 8
        self.sq_codec_hz = 25        # the frame-rate of SQCodec
        self.rec_frame_rate = 12.5
        self.reason_frame_rate = 5
        self.model = model
        self.SQCodec = SQCodec

    def _randn(self, *shape):
        """The reference draws these on the CPU generator and moves them to the device (reason_tokenizer.py:234, :279)."""
        return torch.randn(*shape)

    @torch.no_grad()
    def token2audio_no_reason(self, rec_codec, return_reasoning_text, duration=20, guidance_scale=1.5, num_steps=20, disable_progress=False):
        rec_codec = rec_codec.to(self.device)
        first_latent = self._randn(rec_codec.shape[0], int(duration * 25), 136).to(self.device)
        first_latent_length = 0
        first_latent_codes_length = 0
        min_samples = int(duration * self.rec_frame_rate)
        hop_samples = min_samples // 4 * 3
        ovlp_samples = min_samples - hop_samples
        ovlp_frames = ovlp_samples // 2
        rec_codes_len = rec_codec.shape[-1]
        target_len = int((rec_codes_len - first_latent_codes_length) / 12.5 * self.sample_rate)
        if rec_codes_len < min_samples:
            while rec_codec.shape[-1] < min_samples:
                rec_codec = torch.cat([rec_codec, rec_codec], -1)
            rec_codec = rec_codec[:, :, 0:min_samples]
        rec_codes_len = rec_codec.shape[-1]
        if (rec_codes_len - ovlp_samples) % hop_samples > 0:
            len_codes = math.ceil((rec_codes_len - ovlp_samples) / float(hop_samples)) * hop_samples + ovlp_samples
            while rec_codec.shape[-1] < len_codes:
                rec_codec = torch.cat([rec_codec, rec_codec], -1)
            rec_codec = rec_codec[:, :, 0:len_codes]
        latent_length = int(duration * self.sq_codec_hz)
        latent_list = []
        spk_embeds = None
        for sinx in range(0, rec_codec.shape[-1] - hop_samples, hop_samples):
            codes_input = [rec_codec[:, :, sinx:sinx + min_samples]]
            if sinx == 0:
                incontext_length = first_latent_length
                latents = self.model.inference_codes(codes_input, spk_embeds, first_latent, latent_length, incontext_length,
                                                     additional_feats=[], guidance_scale=1.5, num_steps=num_steps,
                                                     disable_progress=disable_progress, scenario="other_seg")
            else:
                true_latent = latent_list[-1][:, -ovlp_frames:, :]
                len_add_to_latent = latent_length - true_latent.shape[1]
                incontext_length = true_latent.shape[1]
                true_latent = torch.cat([true_latent, self._randn(true_latent.shape[0], len_add_to_latent, true_latent.shape[-1]).to(self.device)], 1)
                latents = self.model.inference_codes(codes_input, spk_embeds, true_latent, latent_length, incontext_length,
                                                     additional_feats=[], guidance_scale=1.5, num_steps=num_steps,
                                                     disable_progress=disable_progress, scenario="other_seg")
            latent_list.append(latents)
        latent_list = [l.float() for l in latent_list]
        latent_list[0] = latent_list[0][:, first_latent_length:, :]
        min_samples = int(duration * self.sample_rate)
        hop_samples = min_samples // 4 * 3
        ovlp_samples = min_samples - hop_samples
        output = None
        for latent in latent_list:
            cur_output = self.SQCodec.decode(latent.transpose(1, 2)).squeeze(0)
            cur_output = cur_output[:, 0:min_samples].detach().cpu()  # B, T
            if output is None:
                output = cur_output
            else:
                ov_win = torch.from_numpy(np.linspace(0, 1, ovlp_samples)[None, :])
                ov_win = torch.cat([ov_win, 1 - ov_win], -1)
                output[:, -ovlp_samples:] = output[:, -ovlp_samples:] * ov_win[:, -ovlp_samples:] + cur_output[:, 0:ovlp_samples] * ov_win[:, 0:ovlp_samples]
                output = torch.cat([output, cur_output[:, ovlp_samples:]], -1)
        return output[:, 0:target_len]

    def detokenize_no_reason(self, rec_codec, return_reasoning_text, min_duration=30, steps=50, guidance_scale=1.5, disable_progress=False):
        """rec_codec: (8, T2)"""
        return self.token2audio_no_reason(rec_codec.unsqueeze(0), return_reasoning_text=return_reasoning_text, guidance_scale=guidance_scale,
                                          num_steps=steps, disable_progress=disable_progress)
