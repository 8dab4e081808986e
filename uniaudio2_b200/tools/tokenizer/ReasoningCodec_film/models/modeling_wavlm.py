"""Drop-in for the WavLM encoder that the reference's AudioDiffusion1D uses as the second SSL front-end of tokenize:

    self.wavlm_encoder = AutoModel.from_pretrained(wav_lm_path).to(device)                                  # AudioDiffusion1D.py:226
    target = self.wavlm_encoder(wav_16k, output_hidden_states=True).hidden_states                            # AudioDiffusion1D.py:366
    target = torch.stack(target, dim=1)[:, 6:10, :].mean(1).transpose(1, 2)                                  # AudioDiffusion1D.py:367-368

The model class behind AutoModel is transformers' WavLMModel (transformers==4.57.0 in the reference's pyproject.toml:25; the code is
not in the reference tree).  `WavLMModel` here has the same state-dict keys (feature_extractor.conv_layers.{i}.conv.weight,
feature_extractor.conv_layers.0.layer_norm.*, feature_projection.*, encoder.pos_conv_embed.conv.{bias, parametrizations.weight.
original0 / original1}, encoder.layer_norm.*, encoder.layers.{i}.{attention.{q,k,v,out}_proj.*, attention.gru_rel_pos_linear.*,
attention.gru_rel_pos_const, layer_norm.*, feed_forward.{intermediate,output}_dense.*, final_layer_norm.*},
encoder.layers.0.attention.rel_attn_embed.weight, masked_spec_embed), so `load_state_dict(hf_model.state_dict())` works, and the
same call: `model(wav_16k, output_hidden_states=True).hidden_states` is the tuple of num_hidden_layers + 1 tensors (B, T, hidden).
`hidden_states_mean(wav_16k, lo, hi)` is the fused form of the reference's stack + slice + mean: only the first hi - 1 layers run
and no per-layer tensor is materialised.

Served configuration: feat_extract_norm='group', do_stable_layer_norm=False, 'gelu' activations (wavlm-base, wavlm-base-plus - the
768-wide models the reference's `wavlm_fea_dim = 768` implies), inference, no attention mask.  fp32 class arithmetic by default;
`set_option('bf16', 1)` switches to the reference's autocast arithmetic (bf16 operands, fp32 accumulation, tensor-core attention).
All arithmetic runs in libua2_b200.so (csrc/ua2_wavlm.cu).  No torch / CPU fallback."""
import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Optional, Tuple

import torch
import torch.nn as nn

from ..... import _lib
from .modeling_whisper import BaseModelOutput, _P, _linear, _norm


@dataclass
class WavLMConfig:
    """The fields of transformers' WavLMConfig that the model reads (defaults: microsoft/wavlm-base-plus)."""
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    hidden_act: str = "gelu"
    layer_norm_eps: float = 1e-5
    feat_extract_norm: str = "group"
    feat_extract_activation: str = "gelu"
    conv_dim: Tuple[int, ...] = (512, 512, 512, 512, 512, 512, 512)
    conv_stride: Tuple[int, ...] = (5, 2, 2, 2, 2, 2, 2)
    conv_kernel: Tuple[int, ...] = (10, 3, 3, 3, 3, 2, 2)
    conv_bias: bool = False
    num_conv_pos_embeddings: int = 128
    num_conv_pos_embedding_groups: int = 16
    num_buckets: int = 320
    max_bucket_distance: int = 800
    do_stable_layer_norm: bool = False

    @property
    def num_feat_extract_layers(self):
        return len(self.conv_dim)


class WavLMModel(nn.Module):
    MAX_BATCH = 8  # clips per native call (the reference's tokenize batches 6); larger batches are served in slices

    def __init__(self, config: WavLMConfig = None, device=None):
        super().__init__()
        c = config if config is not None else WavLMConfig()
        if c.feat_extract_norm != "group" or c.do_stable_layer_norm or c.hidden_act != "gelu" or c.feat_extract_activation != "gelu":
            raise NotImplementedError("served configuration: feat_extract_norm='group', do_stable_layer_norm=False, gelu activations "
                                      "(wavlm-base / wavlm-base-plus)")
        if not (len(c.conv_dim) == len(c.conv_kernel) == len(c.conv_stride)):
            raise ValueError("conv_dim, conv_kernel and conv_stride must have the same length")  # the reference config's own check
        self.config = c
        D, Fi, H = c.hidden_size, c.intermediate_size, c.num_attention_heads
        self.masked_spec_embed = nn.Parameter(torch.rand(D, device=device), requires_grad=False)  # training-time masking only; unused
        fe = nn.Module()
        fe.conv_layers = nn.ModuleList()
        for i, (co, k) in enumerate(zip(c.conv_dim, c.conv_kernel)):
            ci = 1 if i == 0 else c.conv_dim[i - 1]
            L = nn.Module()
            w = torch.empty(co, ci, k, device=device)
            nn.init.kaiming_normal_(w)
            L.conv = _P(weight=w, bias=torch.zeros(co, device=device)) if c.conv_bias else _P(weight=w)
            if i == 0:
                L.layer_norm = _norm(co, device)  # GroupNorm(num_groups = co)
            fe.conv_layers.append(L)
        self.feature_extractor = fe
        fp = nn.Module()
        fp.layer_norm = _norm(c.conv_dim[-1], device)
        fp.projection = _linear(D, c.conv_dim[-1], device)
        self.feature_projection = fp
        enc = nn.Module()
        enc.pos_conv_embed = nn.Module()
        conv = nn.Module()
        cg, K = D // c.num_conv_pos_embedding_groups, c.num_conv_pos_embeddings
        v = torch.randn(D, cg, K, device=device) * (2 * math.sqrt(4 / (K * D)))
        conv.bias = nn.Parameter(torch.zeros(D, device=device), requires_grad=False)
        conv.parametrizations = nn.Module()
        conv.parametrizations.weight = _P(original0=v.norm(dim=(0, 1), keepdim=True), original1=v)  # weight_norm(dim=2): g (1, 1, K), v
        enc.pos_conv_embed.conv = conv
        enc.layer_norm = _norm(D, device)
        enc.layers = nn.ModuleList()
        for i in range(c.num_hidden_layers):
            L = nn.Module()
            a = nn.Module()
            a.k_proj, a.v_proj, a.q_proj, a.out_proj = (_linear(D, D, device) for _ in range(4))
            a.gru_rel_pos_const = nn.Parameter(torch.ones(1, H, 1, 1, device=device), requires_grad=False)
            a.gru_rel_pos_linear = _linear(8, D // H, device)
            if i == 0:
                a.rel_attn_embed = _P(weight=torch.randn(c.num_buckets, H, device=device))
            L.attention = a
            L.layer_norm = _norm(D, device)
            L.feed_forward = nn.Module()
            L.feed_forward.intermediate_dense = _linear(Fi, D, device)
            L.feed_forward.output_dense = _linear(D, Fi, device)
            L.final_layer_norm = _norm(D, device)
            enc.layers.append(L)
        self.encoder = enc
        self._h = None
        self._keep = []
        self._bf16 = 0

    # ------------------------------------------------------------------ native handle
    def _destroy(self):
        if getattr(self, "_h", None) is not None:
            _lib.lib().ua2_wavlm_destroy(self._h)
            self._h = None
            self._keep = []

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    def load_state_dict(self, sd, strict=True, **kw):
        self._destroy()
        return super().load_state_dict(sd, strict=strict, **kw)

    def _apply(self, fn, *a, **kw):
        self._destroy()
        return super()._apply(fn, *a, **kw)

    @property
    def device(self):
        return self.encoder.layer_norm.weight.device

    def _ensure(self):
        if self._h is not None:
            return self._h
        L = _lib.lib()
        dev = self.device
        if dev.type != "cuda":
            raise _lib.Ua2Error("uniaudio2_b200 WavLMModel runs on a CUDA device only (no CPU fallback): call .to('cuda') first")
        c = self.config
        n = c.num_feat_extract_layers
        if n > 8:
            raise ValueError("at most 8 feature-encoder layers")
        arr = lambda v: (C.c_int32 * 8)(*(list(v) + [0] * (8 - n)))
        cfg = _lib.WavLMCfg(c.hidden_size, c.num_attention_heads, c.intermediate_size, c.num_hidden_layers, n, arr(c.conv_dim), arr(c.conv_kernel),
                            arr(c.conv_stride), 1 if c.conv_bias else 0, c.num_conv_pos_embeddings, c.num_conv_pos_embedding_groups, c.num_buckets,
                            c.max_bucket_distance, c.layer_norm_eps)
        h = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(L.ua2_wavlm_create(C.byref(cfg), C.byref(h)), "ua2_wavlm_create")
            keep = []
            try:
                sd = {k: t.detach() for k, t in self.state_dict().items()}
                pre = "encoder.pos_conv_embed.conv."
                g, v = sd.pop(pre + "parametrizations.weight.original0"), sd.pop(pre + "parametrizations.weight.original1")
                sd[pre + "weight"] = g * (v / v.norm(dim=(0, 1), keepdim=True))  # the weight nn.utils.parametrizations.weight_norm(dim=2) yields
                sd.pop("masked_spec_embed", None)
                for key, t in sd.items():
                    if t.dtype != torch.float32:
                        raise _lib.Ua2Error(f"{key} has dtype {t.dtype}; this path takes fp32 parameters")
                    t = t.contiguous()
                    keep.append(t)
                    shape = (C.c_int64 * t.dim())(*t.shape)
                    _lib.check(L.ua2_wavlm_load_weight(h, key.encode(), _lib.ptr(t), shape, t.dim()), f"load_weight({key})")
                _lib.check(L.ua2_wavlm_finalize(h, _lib.current_stream()), "ua2_wavlm_finalize")
                _lib.check(L.ua2_wavlm_set_option(h, b"bf16", self._bf16), "set_option(bf16)")
            except Exception:
                L.ua2_wavlm_destroy(h)
                raise
        self._h, self._keep = h, keep
        return h

    def set_option(self, name: str, value: int):
        """'bf16' (0/1, default 0): the reference's arithmetic for this call (torch.autocast(bfloat16) around fetch_codes_batch,
        reason_tokenizer.py:114-118): bf16 operands with fp32 accumulation on tensor cores, attention included (head size 64); the default
        is fp32 class (3xTF32 GEMMs, fp32 attention)."""
        if name == "bf16":
            self._bf16 = 1 if value else 0
        _lib.check(_lib.lib().ua2_wavlm_set_option(self._ensure(), name.encode(), int(value)), f"set_option({name})")

    def last_launch_count(self) -> int:
        return int(_lib.lib().ua2_wavlm_last_launch_count(self._h)) if self._h is not None else 0

    def num_frames(self, n_samples: int) -> int:
        """Encoder frames for a clip of n_samples (transformers' _get_feat_extract_output_lengths)."""
        t = n_samples
        for k, s in zip(self.config.conv_kernel, self.config.conv_stride):
            t = (t - k) // s + 1
        return t

    # ------------------------------------------------------------------ forward
    def _run(self, input_values, lo, hi, want_all):
        if input_values.dim() != 2:
            raise ValueError(f"expected input_values of shape (B, samples), got {tuple(input_values.shape)}")
        c = self.config
        if not (0 <= lo < hi <= c.num_hidden_layers + 1):
            raise ValueError(f"hidden-state range [{lo}, {hi}) outside [0, {c.num_hidden_layers + 1})")
        h = self._ensure()
        dev = self.device
        x = input_values.to(device=dev, dtype=torch.float32).contiguous()
        B, L = x.shape
        T = self.num_frames(L)
        if T < 1:
            raise RuntimeError(f"clip of {L} samples is shorter than the feature encoder's receptive field")  # the reference's conv1d fails here
        out = torch.empty(B, T, c.hidden_size, device=dev, dtype=torch.float32)
        allh = torch.empty(hi, B, T, c.hidden_size, device=dev, dtype=torch.float32) if want_all else None
        with torch.cuda.device(dev):
            for b0 in range(0, B, self.MAX_BATCH):
                b1 = min(B, b0 + self.MAX_BATCH)
                if want_all:  # a batch slice of (hi, B, T, D) is not contiguous: stage it
                    stage = allh if (b0 == 0 and b1 == B) else torch.empty(hi, b1 - b0, T, c.hidden_size, device=dev, dtype=torch.float32)
                _lib.check(_lib.lib().ua2_wavlm_forward(h, _lib.ptr(x[b0:b1]), L, b1 - b0, L, lo, hi, _lib.ptr(out[b0:b1]),
                                                        _lib.ptr(stage) if want_all else None, _lib.current_stream()), "ua2_wavlm_forward")
                if want_all and stage is not allh:
                    allh[:, b0:b1] = stage
        return out, allh

    @torch.inference_mode()
    def hidden_states_mean(self, input_values, lo=6, hi=10):
        """mean of hidden_states[lo:hi] as (B, T, hidden): `torch.stack(hidden_states, 1)[:, lo:hi].mean(1)` of the reference."""
        return self._run(input_values, lo, hi, False)[0]

    @torch.inference_mode()
    def forward(self, input_values, attention_mask=None, mask_time_indices=None, output_attentions=None, output_hidden_states=None,
                return_dict=None):
        if attention_mask is not None or mask_time_indices is not None or output_attentions:
            raise NotImplementedError("attention_mask / mask_time_indices / output_attentions are not served (the reference passes none)")
        n = self.config.num_hidden_layers
        last, allh = self._run(input_values, n, n + 1, bool(output_hidden_states))
        hs = tuple(allh[i] for i in range(n + 1)) if output_hidden_states else None
        if return_dict is False:
            return (last,) + ((hs,) if hs is not None else ())
        return BaseModelOutput(last_hidden_state=last, hidden_states=hs)
