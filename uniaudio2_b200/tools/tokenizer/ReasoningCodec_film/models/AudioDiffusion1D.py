"""Drop-in for the solver of the reference's tools/tokenizer/ReasoningCodec_film/models/AudioDiffusion1D.py::BASECFM
(:62-129): `solve_euler` with classifier-free guidance, the whole loop (in-context blend, CFG batch assembly, estimator,
guidance mix, Euler update) on the device through ua2_dit_solve_euler - no host synchronisation between steps."""
import ctypes as C

import math

import torch
import torch.nn as nn

from ..... import _lib
from .transformer_1d_flow import Transformer1DModel


class BASECFM(nn.Module):
    def __init__(self, estimator: Transformer1DModel):
        super().__init__()
        if not isinstance(estimator, Transformer1DModel):
            raise TypeError("estimator must be the uniaudio2_b200 Transformer1DModel")
        self.sigma_min = 1e-4
        self.estimator = estimator

    @torch.inference_mode()
    def solve_euler(self, x, incontext_x, incontext_length, t_span, mu, added_cond_kwargs=None, guidance_scale=1.5):
        """x (1, T, latent) noise, incontext_x (1, T, latent), t_span (n + 1,), mu (1, T, cond) -> (1, T, latent).
        The solution is returned as a new tensor; the caller's `x` is left untouched (the reference overwrites its in-context
        rows in place, AudioDiffusion1D.py:106 - no caller reads them afterwards)."""
        est = self.estimator
        h = est._ensure()
        dev = est.device
        if x.dim() != 3 or x.shape[0] != 1:
            raise RuntimeError("solve_euler runs batch 1 (the reference repeats the timestep twice, AudioDiffusion1D.py:114)")
        if not guidance_scale > 1.0:
            raise ValueError("solve_euler is served with classifier-free guidance (guidance_scale > 1) only")
        T, lat = x.shape[1], x.shape[2]
        cond = est.in_channels - 2 * est.out_channels
        if lat != est.out_channels or tuple(incontext_x.shape) != (1, T, lat) or tuple(mu.shape) != (1, T, cond):
            raise ValueError(f"expected x / incontext_x (1, T, {est.out_channels}) and mu (1, T, {cond})")
        xs = x.to(device=dev, dtype=torch.float32).contiguous().clone()
        ic = incontext_x.to(device=dev, dtype=torch.float32).contiguous()
        m = mu.to(device=dev, dtype=torch.float32).contiguous()
        ts = [float(v) for v in torch.as_tensor(t_span, dtype=torch.float32).cpu().tolist()]
        arr = (C.c_float * len(ts))(*ts)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().ua2_dit_solve_euler(h, _lib.ptr(xs), _lib.ptr(ic), int(incontext_length), arr, len(ts), _lib.ptr(m), T,
                                                      float(guidance_scale), float(self.sigma_min), _lib.current_stream()), "solve_euler")
        return xs


class ResidualVQ(nn.Module):
    """Decode surface of vector_quantize_pytorch.ResidualVQ (1.27.15, pyproject.toml:31) as the reference uses it at
    AudioDiffusion1D.py:577-579: get_output_from_indices(indices (B, T, q)) = project_out(sum_q codebook_q[indices[..., q]]).
    State-dict names follow that package: project_out.{weight,bias}, layers.{i}._codebook.embed (1, K, d); its other entries
    (project_in, EMA statistics) are accepted and ignored.  The lookup + sum runs in ua2_rvq_decode_f32."""

    def __init__(self, dim, codebook_size, codebook_dim, num_quantizers, device=None, **unused_training_kwargs):
        super().__init__()
        self.dim, self.codebook_size, self.codebook_dim, self.num_quantizers = dim, codebook_size, codebook_dim, num_quantizers
        self.project_in = nn.Module()  # dim -> codebook_dim (the package inserts it because dim != codebook_dim); encode side only
        self.project_in.weight = nn.Parameter(torch.zeros(codebook_dim, dim, device=device), requires_grad=False)
        self.project_in.bias = nn.Parameter(torch.zeros(codebook_dim, device=device), requires_grad=False)
        self.project_out = nn.Module()
        self.project_out.weight = nn.Parameter(torch.empty(dim, codebook_dim, device=device), requires_grad=False)
        self.project_out.bias = nn.Parameter(torch.zeros(dim, device=device), requires_grad=False)
        self.layers = nn.ModuleList()
        for _ in range(num_quantizers):
            layer = nn.Module()
            layer._codebook = nn.Module()
            layer._codebook.embed = nn.Parameter(torch.empty(1, codebook_size, codebook_dim, device=device), requires_grad=False)
            self.layers.append(layer)
        self._emb = None
        self._sqnorm = None

    def load_state_dict(self, sd, strict=True, **kw):
        keep = set(self.state_dict().keys())
        self._emb = self._sqnorm = None
        sd = {k: v for k, v in sd.items() if k in keep}
        for k in ("project_in.weight", "project_in.bias"):  # decode-only checkpoints / fixtures carry no project_in
            if k not in sd:
                sd[k] = self.state_dict()[k]
        return super().load_state_dict(sd, strict=strict, **kw)

    def _apply(self, fn, *a, **kw):
        self._emb = self._sqnorm = None
        return super()._apply(fn, *a, **kw)

    @torch.inference_mode()
    def encode(self, x, codes, q_off):
        """Eval-mode forward of the package's ResidualVQ on x (B, T, dim): project_in, then per quantizer the nearest code vector
        (Euclidean, lowest index on ties) of the running residual - ua2_rvq_encode_gemm_f32 - writing codes[b, q_off + q, t] of a
        (B, n_q_total, T) tensor; returns project_out(sum of the chosen code vectors) (B, T, dim)."""
        emb = self.codebooks()
        if self._sqnorm is None:
            self._sqnorm = (emb * emb).sum(-1).contiguous()
        B, T, _ = x.shape
        L = _lib.lib()
        with torch.cuda.device(emb.device):
            h = torch.empty(B, T, self.codebook_dim, device=emb.device, dtype=torch.float32)
            _lib.check(L.ua2_linear_bias_f32(_lib.ptr(x), _lib.ptr(self.project_in.weight.detach().float().contiguous()),
                                             _lib.ptr(self.project_in.bias.detach().float().contiguous()), _lib.ptr(h), B * T, self.codebook_dim,
                                             self.dim, _lib.current_stream()), "project_in")
            r = h.clone()  # running residual, frame-major (B * T, codebook_dim), updated in place
            S = torch.empty(B * T, self.codebook_size, device=emb.device, dtype=torch.float32)
            _lib.check(L.ua2_rvq_encode_gemm_f32(_lib.ptr(r), _lib.ptr(emb), _lib.ptr(self._sqnorm), _lib.ptr(S), _lib.ptr(codes), B,
                                                 self.codebook_dim, T, self.codebook_size, self.num_quantizers, codes.shape[1], q_off,
                                                 _lib.current_stream()), "rvq_encode")
            q = (h - r).contiguous()  # sum of the chosen code vectors = input minus the final residual
            out = torch.empty(B, T, self.dim, device=emb.device, dtype=torch.float32)
            _lib.check(L.ua2_linear_bias_f32(_lib.ptr(q), _lib.ptr(self.project_out.weight.detach().float().contiguous()),
                                             _lib.ptr(self.project_out.bias.detach().float().contiguous()), _lib.ptr(out), B * T, self.dim,
                                             self.codebook_dim, _lib.current_stream()), "project_out")
        return out

    def codebooks(self):
        if self._emb is None:  # (q, K, d) contiguous, the layout ua2_rvq_decode_f32 reads
            self._emb = torch.cat([l._codebook.embed.detach() for l in self.layers], 0).float().contiguous()
        return self._emb

    @torch.inference_mode()
    def lookup_sum(self, codes, q_off):
        """codes (B, n_q_total, T) int64 on the device -> sum over this VQ's quantizers [q_off, q_off + q) of the code vectors,
        (B, codebook_dim, T)."""
        emb = self.codebooks()
        if not codes.is_cuda or codes.device != emb.device:
            raise _lib.Ua2Error("ResidualVQ lookup runs on the CUDA device of its codebooks (no CPU fallback)")
        B, nq_total, T = codes.shape
        out = torch.empty(B, self.codebook_dim, T, device=emb.device, dtype=torch.float32)
        with torch.cuda.device(emb.device):
            _lib.check(_lib.lib().ua2_rvq_decode_f32(_lib.ptr(codes), _lib.ptr(emb), _lib.ptr(out), B, self.codebook_dim, T, self.codebook_size,
                                                     self.num_quantizers, nq_total, q_off, _lib.current_stream()), "rvq_decode")
        return out


class AudioDiffusion1D(nn.Module):
    """Inference surface of the reference's AudioDiffusion1D that turns reconstruction codes into SQ-codec latents:
    `inference_codes` (AudioDiffusion1D.py:553-624, the branch without reasoning codes, scenario 'other_seg') and
    `prepare_latents` (:652-655).  Sub-module names follow the reference (vq_pronunciation_semantic, vq_structure_semantic,
    vq_acoustic, cond_feature_emb, zero_cond_embedding1, cfm_wrapper.estimator) so that its checkpoint keys load.
    Arithmetic: code lookups + sums (ua2_rvq_decode_f32), the three project_out and cond_feature_emb linears
    (ua2_linear_bias_f32; the three projections are one GEMM over the concatenated code vectors), the flow-matching solve
    (ua2_dit_solve_euler); torch only moves data (concatenate, repeat frames x2, fill the masked tail)."""

    def __init__(self, estimator: Transformer1DModel, codec_dim=768, codebook_size=8192, codebook_dim=32, sq_codec_latent=136,
                 device=None, whisper_dim=1024, wavlm_dim=768, bestrq_dim=1024):
        super().__init__()
        self.codec_dim, self.sq_codec_latent = codec_dim, sq_codec_latent
        self.max_t_len = 30 * 50
        self.vq_pronunciation_semantic = ResidualVQ(codec_dim, codebook_size, codebook_dim, 1, device=device)
        self.vq_structure_semantic = ResidualVQ(codec_dim, codebook_size, codebook_dim, 1, device=device)
        self.vq_acoustic = ResidualVQ(codec_dim, codebook_size, codebook_dim, 6, device=device)
        self.cond_feature_emb = nn.Module()
        self.cond_feature_emb.weight = nn.Parameter(torch.empty(codec_dim, codec_dim, device=device), requires_grad=False)
        self.cond_feature_emb.bias = nn.Parameter(torch.zeros(codec_dim, device=device), requires_grad=False)
        self.zero_cond_embedding1 = nn.Parameter(torch.zeros(codec_dim, device=device), requires_grad=False)
        self.cfm_wrapper = BASECFM(estimator)
        self._proj = None
        # ---- encode side (fetch_codes_batch :515-551): strided down-samplers of the SSL features, fusion linears, FiLM heads
        self.gamma = 0.1

        def lin(out_f, in_f):
            m = nn.Module()
            m.weight = nn.Parameter(torch.zeros(out_f, in_f, device=device), requires_grad=False)
            m.bias = nn.Parameter(torch.zeros(out_f, device=device), requires_grad=False)
            return m

        def conv(ch, k):
            m = nn.Module()
            m.weight = nn.Parameter(torch.zeros(ch, ch, k, device=device), requires_grad=False)
            m.bias = nn.Parameter(torch.zeros(ch, device=device), requires_grad=False)
            return m

        self.d_conv_whisper, self.d_conv_wavlm = conv(whisper_dim, 4), conv(wavlm_dim, 4)
        self.d_conv_embedding_semantic, self.d_conv_embedding_acoustic = conv(bestrq_dim, 2), conv(bestrq_dim, 2)
        self.cond_fusion_layer_semantic = lin(codec_dim, bestrq_dim)
        self.cond_fusion_layer_acoustic = lin(codec_dim, bestrq_dim + whisper_dim)
        self.cond_fusion_layer_phone = lin(codec_dim, wavlm_dim)
        self.time_film_phone, self.time_film_semantic, self.time_film_acoustic = (lin(2 * codec_dim, codec_dim) for _ in range(3))
        self.reason_adaptor = lin(codec_dim, codec_dim)
        # SSL front-ends (AudioDiffusion1D.py:222-236).  The Whisper encoder (modeling_whisper.py) and the WavLM encoder
        # (modeling_wavlm.py) exist in this package; they are attached by the caller because their sizes come from the checkpoints'
        # configs, not from this class's arguments.  The BEST-RQ conformer is not built (its features are inputs).
        self.whisper_encoder = None
        self.wavlm_encoder = None
        self.wavlm_transfer = None
        self.audio_thinking = None  # the reasoning encoder (models/audio_thinking.py), attached like the SSL encoders
        self.pretrained_model = None  # BEST-RQ feature provider (attach_bestrq); the conformer itself is not built here

    def attach_whisper_encoder(self, encoder):
        """`self.whisper_encoder = WhisperModel.from_pretrained(whisper_path).encoder` (AudioDiffusion1D.py:223): the caller builds the
        drop-in WhisperEncoder (models/modeling_whisper.py), loads the checkpoint's encoder.* tensors into it and hands it over."""
        self.whisper_encoder = encoder
        return encoder

    @torch.inference_mode()
    def get_whisper_feature(self, mels, n_len, len_semantic):
        """AudioDiffusion1D.py:334-343: encoder output cut to the clip's frames (50 Hz), at least twice the BEST-RQ frames, as (B, D, T)."""
        if self.whisper_encoder is None:
            raise _lib.Ua2Error("no Whisper encoder attached (attach_whisper_encoder)")
        n_len = int((n_len / 24000) * 50)
        n_len = max(n_len, len_semantic * 2)
        whisper_embeds = self.whisper_encoder(mels, return_dict=True).last_hidden_state
        return whisper_embeds[:, :n_len, :].transpose(1, 2)

    def attach_bestrq(self, pretrained_model):
        """`self.pretrained_model = BESTRQ_Model(...)` (AudioDiffusion1D.py:228-229): any object with
        `extract_continous_embeds_multiple(input_audios (B, 1, samples)) -> (acoustic (B, 1024, Tb), semantic (B, 1024, Tb))`.  The BEST-RQ
        MusicFM conformer itself is not built in this package: the caller supplies it."""
        self.pretrained_model = pretrained_model
        return pretrained_model

    @torch.inference_mode()
    def fetch_codes_batch(self, input_audios, spectrograms, additional_feats=None, return_reasoning_text=False, film_masks=None):
        """AudioDiffusion1D.py:492-551: input_audios (B, 1, samples) at 24 kHz + Whisper input features (B, 80, 3000) ->
        ([reasoning codes (B, Tq, 8)], [reconstruction codes (B, T, 8)], [merge features (B, T, codec_dim)]) - the BEST-RQ provider,
        the Whisper encoder, the WavLM encoder, the reasoning encoder and the own-code chain of this class in the reference's order."""
        if return_reasoning_text:
            raise NotImplementedError("the reasoning-text head (llama tokenizer + LoRA LLM) is not on this path")
        if getattr(self, "pretrained_model", None) is None:
            raise _lib.Ua2Error("no BEST-RQ feature provider attached (attach_bestrq)")
        bestrq_emb_acoustic, bestrq_emb_semantic = self.pretrained_model.extract_continous_embeds_multiple(input_audios.clone())
        len_semantic = bestrq_emb_semantic.shape[2]
        whisper_embeds = self.get_whisper_feature(spectrograms, input_audios.shape[-1], len_semantic)
        wavlm_embeds = self.get_wavlm_feature(input_audios, len_semantic)
        quantized_reasoning, reasoning_codes, _ = self.encode_reasoning_part(whisper_embeds, bestrq_emb_semantic)
        merge_codes, merge_features = self.fetch_codes_from_features(whisper_embeds, wavlm_embeds, bestrq_emb_acoustic, bestrq_emb_semantic,
                                                                     quantized_reasoning, film_masks=film_masks)
        return [reasoning_codes], [merge_codes], [merge_features]

    def attach_audio_thinking(self, audio_thinking):
        """`self.audio_thinking = AudioThinking(dim=self.codec_dim, interval=5, encoder_depth=5, ...)` (AudioDiffusion1D.py:303)."""
        self.audio_thinking = audio_thinking
        return audio_thinking

    @torch.inference_mode()
    def encode_reasoning_part(self, whisper_embeds, muencoder_embeds):
        """AudioDiffusion1D.py:372-390: (B, 1024, Tw) Whisper features + (B, 1024, Tb) BEST-RQ features -> (quantized reasoning features
        (B, T / 5, codec_dim), reasoning codes (B, T / 5, 8), None)."""
        if self.audio_thinking is None:
            raise _lib.Ua2Error("no reasoning encoder attached (attach_audio_thinking)")
        return self.audio_thinking.encode_reasoning_part(whisper_embeds, muencoder_embeds)

    def attach_wavlm_encoder(self, encoder):
        """`self.wavlm_encoder = AutoModel.from_pretrained(wav_lm_path)` + `self.wavlm_transfer = Resample(24000, 16000)`
        (AudioDiffusion1D.py:226-227): the caller builds the drop-in WavLMModel (models/modeling_wavlm.py), loads the checkpoint into it
        and hands it over; the resampler is the device kernel of ../frontend.py."""
        from ..frontend import Resample

        self.wavlm_encoder = encoder
        self.wavlm_transfer = Resample(24000, 16000)
        return encoder

    @torch.inference_mode()
    def get_wavlm_feature(self, wav_24k, len_semantic):
        """AudioDiffusion1D.py:359-370: wav_24k (B, 1, T) -> (B, wavlm_dim, frames) = mean of hidden states 6..9 of the WavLM encoder on the
        16 kHz waveform with 160 zero samples appended, cut to twice the BEST-RQ frames.  Resampling and the zero tail are one launch; the
        stack / slice / mean of the reference is fused into the encoder call, which stops after the ninth layer."""
        if self.wavlm_encoder is None:
            raise _lib.Ua2Error("no WavLM encoder attached (attach_wavlm_encoder)")
        wav = wav_24k.squeeze(1)
        wav_16k = self.wavlm_transfer(wav, pad_to=self.wavlm_transfer.out_length(wav.shape[-1]) + 160)
        target = self.wavlm_encoder.hidden_states_mean(wav_16k, 6, 10).transpose(1, 2)
        n_len = min(target.shape[-1], len_semantic * 2)
        return target[:, :, :n_len]

    @property
    def device(self):
        return self.zero_cond_embedding1.device

    def _apply(self, fn, *a, **kw):
        self._proj = None
        return super()._apply(fn, *a, **kw)

    def load_state_dict(self, sd, strict=True, **kw):
        self._proj = None
        for vq in (self.vq_pronunciation_semantic, self.vq_structure_semantic, self.vq_acoustic):
            vq._emb = None
        self.cfm_wrapper.estimator._destroy()  # its native handle holds the old parameter addresses
        return super().load_state_dict(sd, strict=strict, **kw)

    def _linear_bias(self, x, w, b):
        M, K = x.shape[0] * x.shape[1], x.shape[2]
        y = torch.empty(x.shape[0], x.shape[1], w.shape[0], device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().ua2_linear_bias_f32(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), M, w.shape[0], K, _lib.current_stream()),
                       "linear_bias")
        return y

    def _draw_zero_cond(self, B):
        """The reference's per-call `torch.rand(B, 1, 1) < 0.2` zero-condition draw of time_film (:435 - yes, at inference too)."""
        return (torch.rand(B, 1, 1, device=self.device) < 0.2).view(-1).to(torch.uint8)

    @torch.inference_mode()
    def fetch_codes_from_features(self, whisper_embeds, wavlm_embeds, bestrq_emb_acoustic, bestrq_emb_semantic, quantized_reasoning,
                                  film_masks=None):
        """Everything fetch_codes_batch (:492-551) does after the SSL front-ends (Whisper, WavLM, BEST-RQ) and the reasoning encoder
        have produced their features - the own-code chain of the codec's encode direction:
          d_conv_* (strided Conv1d k4 s4 / k2 s2) -> cond_fusion_layer_* -> time_film (FiLM from the up-sampled reasoning features)
          -> ResidualVQ (1 + 1 + 6 quantizers of 8192 x 32) -> codes (B, T, 8) = [phone | semantic | 6 x acoustic] and
          merge_features = cond_feature_emb(sum of the three quantized outputs).
        whisper / wavlm (B, C, Tw), bestrq_* (B, 1024, Tw / 2), quantized_reasoning (B, Tq, 768) with Tw / 4 == 2.5 Tq.
        film_masks: three (B,) uint8 tensors replacing the reference's random zero-condition draws (None = draw like the reference)."""
        dev = self.device
        if dev.type != "cuda":
            raise _lib.Ua2Error("AudioDiffusion1D runs on a CUDA device only (no CPU fallback): call .to('cuda') first")
        L = _lib.lib()
        st = _lib.current_stream

        def dconv(x, m, k):
            x = x.to(device=dev, dtype=torch.float32).contiguous()
            B, C, T = x.shape
            T_out = (T - k) // k + 1
            y = torch.empty(B, C, T_out, device=dev, dtype=torch.float32)
            _lib.check(L.ua2_conv1d_f32(_lib.ptr(x), _lib.ptr(m.weight.detach().float().contiguous()), _lib.ptr(m.bias.detach().float().contiguous()),
                                        None, None, _lib.ptr(y), B, C, C, T, k, k, 1, 0, 0, st()), "d_conv")
            return y

        def p32(m):
            return m.weight.detach().float().contiguous(), m.bias.detach().float().contiguous()

        with torch.cuda.device(dev):
            whisper_rec = dconv(whisper_embeds, self.d_conv_whisper, 4)
            wavlm_feat = dconv(wavlm_embeds, self.d_conv_wavlm, 4)
            sem_rec = dconv(bestrq_emb_semantic, self.d_conv_embedding_semantic, 2)
            acoustic = dconv(bestrq_emb_acoustic, self.d_conv_embedding_acoustic, 2)
            qr = quantized_reasoning.to(device=dev, dtype=torch.float32).contiguous()
            B, Tq, D = qr.shape
            ra = self._linear_bias(qr, *p32(self.reason_adaptor))                       # (B, Tq, D)
            T = int(math.floor(Tq * 2.5))
            ra_ct = ra.transpose(1, 2).contiguous()                                      # F.interpolate works on (B, C, T)
            up = torch.empty(B, D, T, device=dev, dtype=torch.float32)
            _lib.check(L.ua2_interp_nearest_f32(_lib.ptr(ra_ct), _lib.ptr(up), B, D, Tq, T, 2.5, st()), "interp")
            reasoning = up.transpose(1, 2).contiguous()                                  # (B, T, D)
            if film_masks is None:
                film_masks = [self._draw_zero_cond(B) for _ in range(3)]
            codes = torch.zeros(B, 8, T, dtype=torch.int64, device=dev)
            total = None
            n = min(acoustic.shape[-1], whisper_rec.shape[-1])
            branches = ((wavlm_feat, self.cond_fusion_layer_phone, self.time_film_phone, self.vq_pronunciation_semantic, 0),
                        (sem_rec, self.cond_fusion_layer_semantic, self.time_film_semantic, self.vq_structure_semantic, 1),
                        (torch.cat([acoustic[:, :, :n], whisper_rec[:, :, :n]], dim=1), self.cond_fusion_layer_acoustic, self.time_film_acoustic,
                         self.vq_acoustic, 2))
            for (feat_bct, fusion, film, vq, q_off), mask in zip(branches, film_masks):
                f = self._linear_bias(feat_bct.transpose(1, 2).contiguous(), *p32(fusion))   # (B, T, D)
                if f.shape[1] != T:
                    raise ValueError(f"feature frames {f.shape[1]} != reasoning frames {T} (the reference's time_film broadcasts)")
                params = self._linear_bias(reasoning, *p32(film))                             # (B, T, 2D)
                out = torch.empty_like(f)
                _lib.check(L.ua2_film_f32(_lib.ptr(params), _lib.ptr(f), _lib.ptr(mask.to(device=dev, dtype=torch.uint8).contiguous()), _lib.ptr(out),
                                          B, T, D, float(self.gamma), st()), "time_film")
                q = vq.encode(out, codes, q_off)
                total = q if total is None else total + q
            merge = self._linear_bias(total.contiguous(), *p32(self.cond_feature_emb))
        return codes.transpose(1, 2).contiguous(), merge

    def prepare_latents(self, batch_size, num_frames, dtype, device):
        return torch.randn((batch_size, num_frames, self.sq_codec_latent), device=device, dtype=dtype)  # randn_tensor, :652-655

    @torch.inference_mode()
    def inference_codes(self, codes, spk_embeds, true_latents, latent_length, incontext_length, additional_feats, guidance_scale=2,
                        num_steps=20, disable_progress=True, scenario="start_seg"):
        if len(codes) != 1 or spk_embeds is not None or additional_feats:
            raise NotImplementedError("served: codes = [reconstruction codes], no speaker embedding (reason_tokenizer.py:268-283)")
        dev = self.device
        if dev.type != "cuda":
            raise _lib.Ua2Error("AudioDiffusion1D.inference_codes runs on a CUDA device only (no CPU fallback): call .to('cuda') first")
        rec = codes[0].to(device=dev, dtype=torch.int64).contiguous()  # (B, 8, T): phone | semantic | 6 x acoustic
        B, _, Tc = rec.shape
        if self._proj is None:  # [P_phone | P_semantic | P_acoustic] and the summed bias: the three project_out as one linear
            vqs = (self.vq_pronunciation_semantic, self.vq_structure_semantic, self.vq_acoustic)
            self._proj = (torch.cat([v.project_out.weight.detach() for v in vqs], 1).float().contiguous(),
                          (vqs[0].project_out.bias.detach() + vqs[1].project_out.bias.detach() + vqs[2].project_out.bias.detach()).float().contiguous())
        summed = [self.vq_pronunciation_semantic.lookup_sum(rec, 0), self.vq_structure_semantic.lookup_sum(rec, 1),
                  self.vq_acoustic.lookup_sum(rec, 2)]
        x = torch.cat(summed, 1).transpose(1, 2).contiguous()  # (B, T, 3 * codebook_dim)
        quantized = self._linear_bias(x, *self._proj)
        merge = self._linear_bias(quantized, self.cond_feature_emb.weight.detach().float().contiguous(),
                                  self.cond_feature_emb.bias.detach().float().contiguous())
        merge = merge.repeat_interleave(2, dim=1)  # F.interpolate(scale_factor=2, mode='nearest') over time, :589
        num_frames = merge.shape[1]
        latents = self.prepare_latents(B, num_frames, torch.float32, dev)
        # latent_masks (:609-616): frames < latent_length keep the condition, the rest get zero_cond_embedding1; frames
        # < incontext_length (scenario 'other_seg') carry the in-context latents
        n_ic = int(incontext_length) if scenario == "other_seg" else 0
        n_ic = max(0, min(n_ic, num_frames))
        merge[:, latent_length:] = self.zero_cond_embedding1.detach().float()
        tl = true_latents.to(device=dev, dtype=torch.float32)
        incontext = torch.zeros_like(tl)
        incontext[:, :n_ic] = tl[:, :n_ic]
        t_span = torch.linspace(0, 1, num_steps + 1)
        out = self.cfm_wrapper.solve_euler(latents, incontext, n_ic, t_span, merge.contiguous(), None, guidance_scale)
        out[:, :n_ic] = incontext[:, :n_ic]
        return out
