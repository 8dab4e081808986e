"""CPU oracle for the codes -> waveform caller of the flow-matching decoder (the boundary
`ReasoningTokenizer.detokenize_no_reason`; section 8(f) rank 1).  TEST INFRASTRUCTURE ONLY - never imported by the product path.

Restated from /root/reference/tools/tokenizer/ReasoningCodec_film (paths relative to it):
    reason_tokenizer.py   token2audio_no_reason :228-306   windows of `duration` s (hop 3/4), in-context continuation from the
                                                           previous window's tail, ScalarModel decode, linear cross-fade
                          detokenize_no_reason  :399-404
    models/AudioDiffusion1D.py  inference_codes :553-624 (branch without reasoning codes), prepare_latents :652-655
and, from the un-vendored dependency vector_quantize_pytorch==1.27.15 (pyproject.toml:31), the published algorithm of
    ResidualVQ.get_output_from_indices: sum over quantizers of codebook rows, then project_out (Linear codebook_dim -> dim).

Parity status: the in-repo parts are PINNED - oracle/make_golden_detok.py executes the unmodified source text of the three
methods (their modules cannot be imported here: omegaconf, whisper, peft, fairseq ... are absent) on stand-in objects and
asserts bit-identical waveforms; fixtures in tests/golden/detok_golden.pt.  ResidualVQ is "parity UNPINNED" (package absent:
the stand-in used by the generator is this file's own restatement).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List

import numpy as np
import torch
import torch.nn.functional as F

from oracle import dit_oracle as DO


def residual_vq_output_from_indices(codebooks: torch.Tensor, proj_w: torch.Tensor, proj_b: torch.Tensor, indices: torch.Tensor):
    """vector_quantize_pytorch.ResidualVQ.get_output_from_indices: codebooks (q, K, d), indices (B, T, q) -> (B, T, dim)."""
    q = codebooks.shape[0]
    codes = torch.stack([codebooks[i][indices[..., i]] for i in range(q)], dim=0)  # (q, B, T, d)
    return F.linear(codes.sum(dim=0), proj_w, proj_b)


class DetokOracle:
    """params: flat dict with the AudioDiffusion1D state-dict names that this path touches
         vq_{pronunciation_semantic,structure_semantic,acoustic}.{codebooks (q, K, d), project_out.weight, project_out.bias}
         cond_feature_emb.{weight,bias}, zero_cond_embedding1
       dit: DitOracle (cfm_wrapper.estimator); sq_decode: latent (B, 136, T) -> wave (B, 1, T * 960)  (SQCodec.decode)."""

    def __init__(self, params: Dict[str, torch.Tensor], dit: DO.DitOracle, sq_decode: Callable, sample_rate=24000, sq_codec_hz=25,
                 rec_frame_rate=12.5, sq_codec_latent=136):
        self.p, self.dit, self.sq_decode = params, dit, sq_decode
        self.sample_rate, self.sq_codec_hz, self.rec_frame_rate, self.lat = sample_rate, sq_codec_hz, rec_frame_rate, sq_codec_latent

    def _vq(self, name, idx):
        p = self.p
        return residual_vq_output_from_indices(p[f"{name}.codebooks"], p[f"{name}.project_out.weight"], p[f"{name}.project_out.bias"], idx)

    def inference_codes(self, codes: torch.Tensor, true_latents, latent_length, incontext_length, guidance_scale, num_steps,
                        randn: Callable):
        """AudioDiffusion1D.inference_codes (:553-624), `len(codes) == 1`, scenario 'other_seg', no speaker embedding."""
        p = self.p
        codes_phone, codes_semantic, codes_acoustic = codes[:, 0:1, :], codes[:, 1:2, :], codes[:, 2:, :]
        batch_size = codes_phone.shape[0]
        q = (self._vq("vq_pronunciation_semantic", codes_phone.transpose(1, 2)) + self._vq("vq_structure_semantic", codes_semantic.transpose(1, 2))
             + self._vq("vq_acoustic", codes_acoustic.transpose(1, 2)))
        merge = F.linear(q, p["cond_feature_emb.weight"], p["cond_feature_emb.bias"])
        merge = F.interpolate(merge.permute(0, 2, 1), scale_factor=2, mode="nearest").permute(0, 2, 1)
        num_frames = merge.shape[1]
        latents = randn((batch_size, num_frames, self.lat))  # prepare_latents
        latent_masks = torch.zeros(latents.shape[0], latents.shape[1], dtype=torch.int64)
        latent_masks[:, 0:latent_length] = 2
        latent_masks[:, 0:incontext_length] = 1  # scenario == 'other_seg'
        merge = (latent_masks > 0.5).unsqueeze(-1) * merge + (latent_masks < 0.5).unsqueeze(-1) * p["zero_cond_embedding1"].unsqueeze(0)
        incontext_latents = true_latents * ((latent_masks > 0.5) * (latent_masks < 1.5)).unsqueeze(-1).float()
        incontext_length = int(((latent_masks > 0.5) * (latent_masks < 1.5)).sum(-1)[0])
        t_span = torch.linspace(0, 1, num_steps + 1)
        latents = self.dit.solve_euler(latents * 1.0, incontext_latents, incontext_length, t_span, merge, guidance_scale)
        latents[:, 0:incontext_length, :] = incontext_latents[:, 0:incontext_length, :]
        return latents

    def token2audio_no_reason(self, rec_codec: torch.Tensor, duration=20, num_steps=20, randn: Callable = None):
        """ReasoningTokenizer.token2audio_no_reason (:228-306): rec_codec (B, 8, T2) int64 -> wave (B, samples) on the host.
        (The reference pins guidance_scale = 1.5 in its inference_codes calls, :273/:282.)"""
        randn = randn or (lambda shape: torch.randn(*shape))
        first_latent = randn((rec_codec.shape[0], int(duration * 25), self.lat))
        first_latent_length = 0
        min_samples = int(duration * self.rec_frame_rate)
        hop_samples = min_samples // 4 * 3
        ovlp_samples = min_samples - hop_samples
        ovlp_frames = ovlp_samples // 2
        rec_codes_len = rec_codec.shape[-1]
        target_len = int((rec_codes_len - 0) / 12.5 * self.sample_rate)
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
        latent_list: List[torch.Tensor] = []
        for sinx in range(0, rec_codec.shape[-1] - hop_samples, hop_samples):
            codes_input = rec_codec[:, :, sinx:sinx + min_samples]
            if sinx == 0:
                latents = self.inference_codes(codes_input, first_latent, latent_length, first_latent_length, 1.5, num_steps, randn)
            else:
                true_latent = latent_list[-1][:, -ovlp_frames:, :]
                len_add = latent_length - true_latent.shape[1]
                incontext_length = true_latent.shape[1]
                true_latent = torch.cat([true_latent, randn((true_latent.shape[0], len_add, true_latent.shape[-1]))], 1)
                latents = self.inference_codes(codes_input, true_latent, latent_length, incontext_length, 1.5, num_steps, randn)
            latent_list.append(latents)
        latent_list = [l.float() for l in latent_list]
        latent_list[0] = latent_list[0][:, first_latent_length:, :]
        min_samples = int(duration * self.sample_rate)
        hop_samples = min_samples // 4 * 3
        ovlp_samples = min_samples - hop_samples
        output = None
        for latent in latent_list:
            cur_output = self.sq_decode(latent.transpose(1, 2)).squeeze(0)
            cur_output = cur_output[:, 0:min_samples].detach().cpu()
            if output is None:
                output = cur_output
            else:
                ov_win = torch.from_numpy(np.linspace(0, 1, ovlp_samples)[None, :])
                ov_win = torch.cat([ov_win, 1 - ov_win], -1)
                output[:, -ovlp_samples:] = output[:, -ovlp_samples:] * ov_win[:, -ovlp_samples:] + cur_output[:, 0:ovlp_samples] * ov_win[:, 0:ovlp_samples]
                output = torch.cat([output, cur_output[:, ovlp_samples:]], -1)
        return output[:, 0:target_len]
