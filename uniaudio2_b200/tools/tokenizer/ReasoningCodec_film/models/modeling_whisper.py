"""Drop-in for the ENCODER of the reference's tools/tokenizer/ReasoningCodec_film/models/modeling_whisper.py (its own fork of the
transformers Whisper code), which AudioDiffusion1D uses as the first SSL front-end of tokenize:

    self.whisper_encoder = WhisperModel.from_pretrained(whisper_path).encoder            # AudioDiffusion1D.py:223
    whisper_embeds = self.whisper_encoder(mels, return_dict=True).last_hidden_state      # AudioDiffusion1D.py:340

Same state-dict keys as `WhisperEncoder` (conv1.*, conv2.*, embed_positions.weight, layers.{i}.{self_attn.{q,k,v,out}_proj.*,
self_attn_layer_norm.*, fc1.*, fc2.*, final_layer_norm.*}, layer_norm.*; k_proj has no bias), same call and result type.  Inference
surface only: no attention / head masks, no output_attentions (the reference passes none), output_hidden_states is not served.
All arithmetic runs in libua2_b200.so (csrc/ua2_enc.cu).  No torch / CPU fallback.

`WhisperModel(config).encoder` is provided so that `WhisperModel(...).encoder` call sites keep working; the decoder half of
Whisper is not on the tokenize path and is not built."""
import ctypes as C
import math
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

from ..... import _lib


@dataclass
class WhisperConfig:
    """The fields of transformers' WhisperConfig that the encoder reads (defaults: openai/whisper-medium)."""
    d_model: int = 1024
    encoder_attention_heads: int = 16
    encoder_ffn_dim: int = 4096
    encoder_layers: int = 24
    max_source_positions: int = 1500
    num_mel_bins: int = 80
    activation_function: str = "gelu"
    scale_embedding: bool = False


@dataclass
class BaseModelOutput:
    last_hidden_state: torch.Tensor
    hidden_states: Optional[tuple] = None
    attentions: Optional[tuple] = None


class _P(nn.Module):
    def __init__(self, **tensors):
        super().__init__()
        for name, t in tensors.items():
            setattr(self, name, nn.Parameter(t, requires_grad=False))


def _linear(n_out, n_in, device, bias=True):
    w = torch.empty(n_out, n_in, device=device)
    nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    if not bias:
        return _P(weight=w)
    return _P(weight=w, bias=torch.empty(n_out, device=device).uniform_(-1 / math.sqrt(n_in), 1 / math.sqrt(n_in)))


def _norm(d, device):
    return _P(weight=torch.ones(d, device=device), bias=torch.zeros(d, device=device))


class WhisperEncoder(nn.Module):
    def __init__(self, config: WhisperConfig, device=None):
        super().__init__()
        if config.activation_function != "gelu" or config.scale_embedding:
            raise NotImplementedError("served configuration: activation_function='gelu', scale_embedding=False (every Whisper checkpoint)")
        self.config = config
        d, f = config.d_model, config.encoder_ffn_dim
        w1 = torch.empty(d, config.num_mel_bins, 3, device=device)
        w2 = torch.empty(d, d, 3, device=device)
        nn.init.kaiming_uniform_(w1, a=math.sqrt(5))
        nn.init.kaiming_uniform_(w2, a=math.sqrt(5))
        self.conv1 = _P(weight=w1, bias=torch.zeros(d, device=device))
        self.conv2 = _P(weight=w2, bias=torch.zeros(d, device=device))
        self.embed_positions = _P(weight=torch.randn(config.max_source_positions, d, device=device) * 0.02)
        layers = nn.ModuleList()
        for _ in range(config.encoder_layers):
            L = nn.Module()
            L.self_attn = nn.Module()
            L.self_attn.k_proj = _linear(d, d, device, bias=False)  # modeling_whisper.py:240
            L.self_attn.v_proj = _linear(d, d, device)
            L.self_attn.q_proj = _linear(d, d, device)
            L.self_attn.out_proj = _linear(d, d, device)
            L.self_attn_layer_norm = _norm(d, device)
            L.fc1 = _linear(f, d, device)
            L.fc2 = _linear(d, f, device)
            L.final_layer_norm = _norm(d, device)
            layers.append(L)
        self.layers = layers
        self.layer_norm = _norm(d, device)
        self._h = None
        self._keep = []
        self._bf16 = 0

    # ------------------------------------------------------------------ native handle
    def _destroy(self):
        if getattr(self, "_h", None) is not None:
            _lib.lib().ua2_whisper_destroy(self._h)
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
        return self.layer_norm.weight.device

    def _ensure(self):
        if self._h is not None:
            return self._h
        L = _lib.lib()
        dev = self.device
        if dev.type != "cuda":
            raise _lib.Ua2Error("uniaudio2_b200 WhisperEncoder runs on a CUDA device only (no CPU fallback): call .to('cuda') first")
        c = self.config
        cfg = _lib.WhisperCfg(c.d_model, c.encoder_attention_heads, c.encoder_ffn_dim, c.encoder_layers, c.max_source_positions, c.num_mel_bins)
        h = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(L.ua2_whisper_create(C.byref(cfg), C.byref(h)), "ua2_whisper_create")
            keep = []
            try:
                for key, t in self.state_dict().items():
                    t = t.detach()
                    if t.dtype != torch.float32:
                        raise _lib.Ua2Error(f"{key} has dtype {t.dtype}; this path takes fp32 parameters")
                    t = t.contiguous()
                    keep.append(t)
                    shape = (C.c_int64 * t.dim())(*t.shape)
                    _lib.check(L.ua2_whisper_load_weight(h, key.encode(), _lib.ptr(t), shape, t.dim()), f"load_weight({key})")
                _lib.check(L.ua2_whisper_finalize(h, _lib.current_stream()), "ua2_whisper_finalize")
                _lib.check(L.ua2_whisper_set_option(h, b"bf16", self._bf16), "set_option(bf16)")
            except Exception:
                L.ua2_whisper_destroy(h)
                raise
        self._h, self._keep = h, keep
        return h

    def set_option(self, name: str, value: int):
        """'bf16' (0/1, default 0): the reference's arithmetic for this call (torch.autocast(bfloat16), reason_tokenizer.py:114-118):
        bf16 operands with fp32 accumulation on tensor cores, attention included; the default is fp32 class (3xTF32)."""
        if name == "bf16":
            self._bf16 = 1 if value else 0
        _lib.check(_lib.lib().ua2_whisper_set_option(self._ensure(), name.encode(), int(value)), f"set_option({name})")

    def last_launch_count(self) -> int:
        return int(_lib.lib().ua2_whisper_last_launch_count(self._h)) if self._h is not None else 0

    # ------------------------------------------------------------------ forward
    @torch.inference_mode()
    def forward(self, input_features, attention_mask=None, head_mask=None, output_attentions=None, output_hidden_states=None,
                return_dict=None):
        if head_mask is not None or output_attentions or output_hidden_states:
            raise NotImplementedError("head_mask / output_attentions / output_hidden_states are not served (no reference call site uses them)")
        c = self.config
        if input_features.dim() != 3 or input_features.shape[1] != c.num_mel_bins or input_features.shape[2] != 2 * c.max_source_positions:
            # the reference fails at `inputs_embeds + embed_pos` (modeling_whisper.py:811) for any other length
            raise RuntimeError(f"expected input_features of shape (B, {c.num_mel_bins}, {2 * c.max_source_positions}), got {tuple(input_features.shape)}")
        h = self._ensure()
        dev = self.device
        x = input_features.to(device=dev, dtype=torch.float32).contiguous()
        B = x.shape[0]
        out = torch.empty(B, c.max_source_positions, c.d_model, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().ua2_whisper_forward(h, _lib.ptr(x), _lib.ptr(out), B, _lib.current_stream()), "forward")
        if return_dict is False:
            return (out,)
        return BaseModelOutput(last_hidden_state=out)


class WhisperModel(nn.Module):
    """`WhisperModel(config).encoder` like the reference's use (AudioDiffusion1D.py:223); encoder only."""

    def __init__(self, config: WhisperConfig, device=None):
        super().__init__()
        self.config = config
        self.encoder = WhisperEncoder(config, device=device)

    def get_encoder(self):
        return self.encoder
