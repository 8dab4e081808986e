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
        (The reference pins guidance_scale = 1.5 in its inference_codes calls, :273/:282.)  Written with explicit index
        arithmetic: the reference's repeated `cat([x, x])` + crop is periodic indexing, first with the period of the input
        (when it is shorter than a window), then with the period of that result (to reach a whole number of hops)."""
        randn = randn or (lambda shape: torch.randn(*shape))
        n_in = rec_codec.shape[-1]
        win = int(duration * self.rec_frame_rate)      # code frames per window
        hop = win // 4 * 3
        ov = win - hop
        n_lat = int(duration * self.sq_codec_hz)       # latent frames per window
        prior = randn((rec_codec.shape[0], int(duration * 25), self.lat))
        idx = torch.arange(max(n_in, win)) % n_in      # :247-250
        if (len(idx) - ov) % hop > 0:                  # :252-256
            n_full = math.ceil((len(idx) - ov) / float(hop)) * hop + ov
            idx = idx[torch.arange(n_full) % len(idx)]
        codes = rec_codec[:, :, idx]
        lat_windows: List[torch.Tensor] = []
        for k, start in enumerate(range(0, codes.shape[-1] - hop, hop)):
            pinned = 0
            if k > 0:                                  # :276-283: the last ov // 2 latent frames of the previous window lead
                tail = lat_windows[-1][:, -(ov // 2):, :]
                pinned = tail.shape[1]
                prior = torch.cat([tail, randn((tail.shape[0], n_lat - pinned, tail.shape[-1]))], 1)
            lat_windows.append(self.inference_codes(codes[:, :, start:start + win], prior, n_lat, pinned, 1.5, num_steps, randn))
        n_wave = int(duration * self.sample_rate)
        wave_ov = n_wave - n_wave // 4 * 3
        fade_in = torch.from_numpy(np.linspace(0, 1, wave_ov)[None, :])  # float64, :299-301
        out = None
        for lat in lat_windows:
            cur = self.sq_decode(lat.float().transpose(1, 2)).squeeze(0)[:, :n_wave].detach().cpu()
            if out is None:
                out = cur
                continue
            blended = out[:, -wave_ov:] * (1 - fade_in) + cur[:, :wave_ov] * fade_in
            out = torch.cat([out[:, :-wave_ov], blended.to(out.dtype), cur[:, wave_ov:]], -1)
        return out[:, :int(n_in / 12.5 * self.sample_rate)]
