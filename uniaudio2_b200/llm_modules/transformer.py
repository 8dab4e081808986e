"""Drop-in for the reference's llm_modules/transformer.py::StreamingTransformer (inference surface, fp32).

Same constructor arguments (transformer.py:616-669 plus the StreamingTransformerLayer keywords :449-470 it forwards),
same state-dict key names (layers.{i}.self_attn.in_proj_weight, .self_attn.out_proj.weight, .norm1/.norm2 (weight, bias |
alpha), .linear1/.linear2.weight | .gating[.{s}].linear_in/.linear_out.weight, .layer_scale_{1,2}.scale), same streaming
API as llm_modules/streaming.py::StreamingModule:

    with model.streaming(batch_size):        # ring KV cache of `context` slots per layer (transformer.py:337-352)
        y = model(x)                          # x (B, T, d_model) -> (B, T, d_model), offsets advance by T
        model.reset_streaming()
    model.streaming_forever(batch_size); model.is_streaming

All arithmetic runs in libua2_b200.so (csrc/ua2_stream.cu + the skinny-linear family); this module owns the parameters and
the native handle.  No torch / CPU fallback.  Served: causal or non-causal attention (streaming needs causal, like the
reference), norm in {layer_norm, layer_norm_f32, rms_norm, rms_norm_f32}, gating in {none (GELU feed-forward), silu},
positional_embedding in {sin, rope, sin_rope, none}, layer_scale, weights_per_step with a per-step dim_feedforward list.
"""
import ctypes as C
import math
from contextlib import contextmanager

import torch
import torch.nn as nn

from .. import _lib

_NORMS = {"layer_norm": 0, "layer_norm_f32": 1, "rms_norm": 2, "rms_norm_f32": 3}
_POS = {"none": 0, "sin": 1, "rope": 2, "sin_rope": 3}


class _P(nn.Module):
    def __init__(self, **tensors):
        super().__init__()
        for name, t in tensors.items():
            setattr(self, name, nn.Parameter(t, requires_grad=False))


def _gating_hidden(dim: int, dim_feedforward: int) -> int:
    """ActivationGating.__init__, gating.py:40-43."""
    return (21 * dim) // 8 if dim_feedforward == 4 * dim else (2 * dim_feedforward) // 3


class StreamingTransformer(nn.Module):
    def __init__(self, d_model: int, num_heads: int, num_layers: int, dim_feedforward=2048, causal: bool = False, context=None,
                 positional_embedding: str = "sin", max_period: float = 10_000, positional_scale: float = 1.0, betas=None,
                 layer_class=None, device=None, dtype=None, norm: str = "layer_norm", layer_scale=None, gating: str = "none",
                 weights_per_step: int = 0, activation=None, skip_self_attn: bool = False):
        super().__init__()
        assert d_model % num_heads == 0
        assert positional_embedding in {"sin", "rope", "sin_rope", "none"}
        if norm not in _NORMS:
            raise ValueError(f"Unknown norm type: {norm}")  # create_norm_fn, transformer.py:122-123
        if dtype not in (None, torch.float32):
            raise _lib.Ua2Error("this path computes in fp32")
        if layer_class is not None or skip_self_attn:
            raise NotImplementedError("custom layer_class / skip_self_attn are not on this path")
        if gating not in ("none", "silu"):
            raise NotImplementedError(f"gating '{gating}': only 'none' and 'silu' are served")
        if gating == "none":
            assert not weights_per_step, "weights_per_step without gating not supported for now."
            assert not isinstance(dim_feedforward, list), "List dim_feedforward without gating not supported for now."
            if activation is not None and activation is not torch.nn.functional.gelu:
                raise NotImplementedError("the ungated feed-forward is served with F.gelu (the reference default)")
        if isinstance(dim_feedforward, list):
            assert dim_feedforward
            assert len(dim_feedforward) == weights_per_step, (
                "Length of dim_feedforward must match weights_per_step,"
                f" got {len(dim_feedforward)} != {weights_per_step}")
        self.d_model, self.num_heads, self.num_layers = d_model, num_heads, num_layers
        self.causal, self.context = causal, context
        self.positional_embedding, self.max_period, self.positional_scale = positional_embedding, max_period, positional_scale
        self.betas = betas
        self.norm, self.gating_name, self.weights_per_step = norm, gating, weights_per_step
        self.layer_scale = layer_scale
        steps = weights_per_step if weights_per_step else 1
        self.ff_dims = list(dim_feedforward) if isinstance(dim_feedforward, list) else [dim_feedforward] * steps
        D, mult = d_model, (weights_per_step if weights_per_step else 1)
        layers = nn.ModuleList()
        for _ in range(num_layers):
            lay = nn.Module()
            lay.self_attn = _P(in_proj_weight=torch.empty(mult * 3 * D, D, device=device))
            lay.self_attn.out_proj = _P(weight=torch.empty(mult * D, D, device=device))
            for n in ("norm1", "norm2"):
                if norm.startswith("layer_norm"):
                    setattr(lay, n, _P(weight=torch.ones(D, device=device), bias=torch.zeros(D, device=device)))
                else:
                    setattr(lay, n, _P(alpha=torch.ones(1, 1, D, device=device)))
            if gating == "none":
                lay.linear1 = _P(weight=torch.empty(self.ff_dims[0], D, device=device))
                lay.linear2 = _P(weight=torch.empty(D, self.ff_dims[0], device=device))
            elif weights_per_step:
                gs = nn.ModuleList()
                for ff in self.ff_dims:
                    g = nn.Module()
                    g.linear_in = _P(weight=torch.empty(2 * _gating_hidden(D, ff), D, device=device))
                    g.linear_out = _P(weight=torch.empty(D, _gating_hidden(D, ff), device=device))
                    gs.append(g)
                lay.gating = gs
            else:
                g = nn.Module()
                g.linear_in = _P(weight=torch.empty(2 * _gating_hidden(D, self.ff_dims[0]), D, device=device))
                g.linear_out = _P(weight=torch.empty(D, _gating_hidden(D, self.ff_dims[0]), device=device))
                lay.gating = g
            if layer_scale is not None:
                lay.layer_scale_1 = _P(scale=torch.full((D,), float(layer_scale), device=device))
                lay.layer_scale_2 = _P(scale=torch.full((D,), float(layer_scale), device=device))
            layers.append(lay)
        self.layers = layers
        with torch.no_grad():
            for name, p in self.named_parameters():  # nn.Linear's default init, so a fresh module is usable
                if p.dim() == 2:
                    nn.init.kaiming_uniform_(p, a=math.sqrt(5))
        self._h = None
        self._keep = []
        self._streaming_batch = None

    # ------------------------------------------------------------------ native handle
    def _destroy(self):
        if self._h is not None:
            _lib.lib().ua2_stx_destroy(self._h)
            self._h = None
            self._keep = []
            self._streaming_batch = None

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    def load_state_dict(self, sd, strict=True, **kw):
        self._destroy()
        return super().load_state_dict(sd, strict=strict, **kw)

    def _apply(self, fn, *a, **kw):  # .to(device) / .cuda(): parameters move, the handle is rebuilt lazily
        self._destroy()
        return super()._apply(fn, *a, **kw)

    def _device(self):
        return self.layers[0].self_attn.in_proj_weight.device

    def _ensure(self):
        if self._h is not None:
            return
        L = _lib.lib()
        dev = self._device()
        if dev.type != "cuda":
            raise _lib.Ua2Error("uniaudio2_b200 StreamingTransformer runs on a CUDA device only (no CPU fallback): call .to('cuda') first")
        ff = (C.c_int32 * 64)(*(self.ff_dims + [0] * (64 - len(self.ff_dims))))
        cfg = _lib.StxCfg(self.d_model, self.num_heads, self.num_layers, int(bool(self.causal)), int(self.context or 0),
                          _POS[self.positional_embedding], _NORMS[self.norm], 1 if self.gating_name == "silu" else 0,
                          int(self.weights_per_step), 0 if self.layer_scale is None else 1, ff, float(self.max_period),
                          float(self.positional_scale))
        h = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(L.ua2_stx_create(C.byref(cfg), C.byref(h)), "ua2_stx_create")
            keep = []
            tensors = {k: v.detach() for k, v in self.state_dict().items()}
            hd = self.d_model // self.num_heads
            if self.positional_embedding in ("rope", "sin_rope"):  # freqs of apply_rope, rope.py:37-38, evaluated like the reference
                ds = torch.arange(hd // 2, device=dev, dtype=torch.float32)
                tensors["rope_freqs"] = torch.exp(ds * (-math.log(self.max_period) * 2 / hd))
            if self.positional_embedding in ("sin", "sin_rope"):  # divisors of create_sin_embedding, transformer.py:147-151
                half = self.d_model // 2
                adim = torch.arange(half, device=dev, dtype=torch.float32)
                tensors["sin_denoms"] = torch.full([], self.max_period, device=dev, dtype=torch.float32) ** (adim / (half - 1))
            for key, t in tensors.items():
                if t.dtype != torch.float32:
                    raise _lib.Ua2Error(f"{key} has dtype {t.dtype}; this path computes in fp32")
                t = t.contiguous()
                keep.append(t)
                shape = (C.c_int64 * t.dim())(*t.shape)
                _lib.check(L.ua2_stx_load_weight(h, key.encode(), _lib.ptr(t), shape, t.dim()), f"load_weight({key})")
            _lib.check(L.ua2_stx_finalize(h), "ua2_stx_finalize")
        self._h, self._keep = h, keep

    # ------------------------------------------------------------------ StreamingModule API (llm_modules/streaming.py)
    @property
    def is_streaming(self):
        return self._streaming_batch is not None

    def _start_streaming(self, batch_size: int):
        if self.context is None and not self.weights_per_step:
            raise RuntimeError("Cannot create a streaming KVCache without a context to estimate capacity.")  # transformer.py:341-343
        self._ensure()
        with torch.cuda.device(self._device()):
            _lib.check(_lib.lib().ua2_stx_start_streaming(self._h, int(batch_size), _lib.current_stream()), "start_streaming")
        self._streaming_batch = int(batch_size)

    def _stop_streaming(self):
        if self._h is not None:
            _lib.check(_lib.lib().ua2_stx_stop_streaming(self._h), "stop_streaming")
        self._streaming_batch = None

    def streaming_forever(self, batch_size: int):
        self._start_streaming(batch_size)

    @contextmanager
    def streaming(self, batch_size: int):
        """Context manager to enter streaming mode. Reset streaming state on exit."""
        self._start_streaming(batch_size)
        try:
            yield
        finally:
            self._stop_streaming()

    def reset_streaming(self):
        if not self.is_streaming:
            raise ValueError("Trying to reset streaming, but  wasn't streaming.")  # streaming.py:118-121
        _lib.check(_lib.lib().ua2_stx_reset_streaming(self._h), "reset_streaming")

    def set_option(self, name: str, value: int):
        """'graph' / 'pdl' (0/1): how streaming calls are launched (include/ua2_b200.h, ua2_stx_set_option)."""
        self._ensure()
        _lib.check(_lib.lib().ua2_stx_set_option(self._h, name.encode(), int(value)), f"set_option({name})")

    def last_launch_count(self) -> int:
        return int(_lib.lib().ua2_stx_last_launch_count(self._h)) if self._h is not None else 0

    def streaming_kv(self, layer: int):
        """(k, v, end_offset): zero-copy views of a layer's ring buffers (batch, H, capacity, head_dim) - RingKVCache.cache[0/1]."""
        from ..llm_models.model_new import _from_ptr

        if not self.is_streaming:
            raise ValueError("not streaming")
        k, v, end, cap = C.c_void_p(), C.c_void_p(), C.c_int64(), C.c_int()
        _lib.check(_lib.lib().ua2_stx_get_kv(self._h, layer, C.byref(k), C.byref(v), C.byref(end), C.byref(cap)), "get_kv")
        shape = (self._streaming_batch, self.num_heads, cap.value, self.d_model // self.num_heads)
        return _from_ptr(k.value, shape, self._device()), _from_ptr(v.value, shape, self._device()), int(end.value)

    # ------------------------------------------------------------------ forward
    @torch.inference_mode()
    def forward(self, x: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        self._ensure()
        if x.dim() != 3 or x.shape[2] != self.d_model:
            raise ValueError(f"expected x of shape (B, T, {self.d_model})")
        dev = self._device()
        xin = x.to(device=dev, dtype=torch.float32).contiguous()
        B, T, _ = xin.shape
        if self.is_streaming:
            assert self.causal, "Streaming only available for causal"  # transformer.py:381
        y = torch.empty_like(xin)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().ua2_stx_forward(self._h, _lib.ptr(xin), _lib.ptr(y), B, T, _lib.current_stream()), "forward")
        return y
