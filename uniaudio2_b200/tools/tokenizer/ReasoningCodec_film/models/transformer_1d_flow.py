"""Drop-in for the reference's tools/tokenizer/ReasoningCodec_film/models/transformer_1d_flow.py::Transformer1DModel
(the DiT estimator of the flow-matching decoder; inference surface, fp32-class arithmetic).

Same constructor keywords as the reference / models/model_config.json, same state-dict keys (proj_in.ffn_{1,2}.*,
pos_embed.pe, transformer_blocks.{i}.{scale_shift_table, attn1.to_{q,k,v}.*, attn1.to_out.0.*, ff.net.0.proj.*, ff.net.2.*},
scale_shift_table, proj_out.ffn_{1,2}.*, adaln_single.emb.timestep_embedder.linear_{1,2}.*, adaln_single.linear.*), same call:

    out = model(hidden_states (B, T, in_channels), timestep=t (B,), added_cond_kwargs={...}).sample   # (B, T, out_channels)

Only the configuration the reference ships is served: norm_type='ada_norm_single', activation_fn='gelu-approximate',
attention_bias=True, no cross attention (transformer_1d_flow.py:213-231 with model_config.json).  All arithmetic runs in
libua2_b200.so (csrc/ua2_dit.cu; linears on the tcgen05 3xTF32 path for >= 32 rows).  No torch / CPU fallback.
"""
import ctypes as C
import math
from dataclasses import dataclass

import torch
import torch.nn as nn

from ..... import _lib


class _P(nn.Module):
    def __init__(self, **tensors):
        super().__init__()
        for name, t in tensors.items():
            if name.startswith("buf_"):
                self.register_buffer(name[4:], t)
            else:
                setattr(self, name, nn.Parameter(t, requires_grad=False))


def _linear(n_out, n_in, device):
    w = torch.empty(n_out, n_in, device=device)
    nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    b = torch.empty(n_out, device=device).uniform_(-1 / math.sqrt(n_in), 1 / math.sqrt(n_in))
    return _P(weight=w, bias=b)


def _project_layer(n_in, n_out, device):
    """ProjectLayer(hidden_size, filter_size, kernel_size=3): Conv1d(k3, 'same') -> * 3 ** -0.5 -> Linear (:19-34)."""
    m = nn.Module()
    w = torch.empty(n_out, n_in, 3, device=device)
    nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    m.ffn_1 = _P(weight=w, bias=torch.zeros(n_out, device=device))
    m.ffn_2 = _linear(n_out, n_out, device)
    return m


@dataclass
class Transformer1DModelOutput:
    sample: torch.Tensor


class Transformer1DModel(nn.Module):
    def __init__(self, num_attention_heads: int = 16, attention_head_dim: int = 88, in_channels=None, out_channels=None,
                 num_layers: int = 1, dropout: float = 0.0, norm_num_groups: int = 32, num_positional_embeddings: int = 3000,
                 cross_attention_dim=None, attention_bias: bool = False, sample_size=None, num_vector_embeds=None, patch_size=None,
                 activation_fn: str = "geglu", num_embeds_ada_norm=None, use_linear_projection: bool = False,
                 only_cross_attention: bool = False, double_self_attention: bool = False, upcast_attention: bool = False,
                 norm_type: str = "layer_norm", norm_elementwise_affine: bool = True, norm_eps: float = 1e-5,
                 attention_type: str = "default", caption_channels=None, device=None, **unused_config_keys):
        super().__init__()
        if (norm_type != "ada_norm_single" or activation_fn != "gelu-approximate" or not attention_bias or norm_elementwise_affine
                or cross_attention_dim is not None or only_cross_attention or double_self_attention or attention_type != "default"):
            raise NotImplementedError("served configuration: norm_type='ada_norm_single', activation_fn='gelu-approximate', "
                                      "attention_bias=True, norm_elementwise_affine=False, self-attention only (model_config.json)")
        self.num_attention_heads, self.attention_head_dim = num_attention_heads, attention_head_dim
        self.in_channels, self.out_channels, self.num_layers = in_channels, out_channels, num_layers
        self.norm_eps, self.num_positional_embeddings = norm_eps, num_positional_embeddings
        self.flow_t_size = 512  # PixArtAlphaCombinedFlowEmbeddings.flow_t_size, transformer_1d_flow.py:47
        D = num_attention_heads * attention_head_dim
        self.proj_in = _project_layer(in_channels, D, device)
        # diffusers SinusoidalPositionalEmbedding: pe[0, :, 0::2] = sin, pe[0, :, 1::2] = cos
        position = torch.arange(num_positional_embeddings).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, D, 2) * (-math.log(10000.0) / D))
        pe = torch.zeros(1, num_positional_embeddings, D)
        pe[0, :, 0::2] = torch.sin(position * div_term)
        pe[0, :, 1::2] = torch.cos(position * div_term)
        self.pos_embed = _P(buf_pe=pe.to(device) if device is not None else pe)
        blocks = nn.ModuleList()
        for _ in range(num_layers):
            b = nn.Module()
            b.scale_shift_table = nn.Parameter(torch.randn(6, D, device=device) / D ** 0.5, requires_grad=False)
            b.attn1 = nn.Module()
            b.attn1.to_q, b.attn1.to_k, b.attn1.to_v = _linear(D, D, device), _linear(D, D, device), _linear(D, D, device)
            b.attn1.to_out = nn.ModuleList([_linear(D, D, device)])
            b.ff = nn.Module()
            proj = nn.Module()
            proj.proj = _linear(4 * D, D, device)
            b.ff.net = nn.ModuleList([proj, nn.Identity(), _linear(D, 4 * D, device)])
            blocks.append(b)
        self.transformer_blocks = blocks
        self.scale_shift_table = nn.Parameter(torch.randn(2, D, device=device) / D ** 0.5, requires_grad=False)
        self.proj_out = _project_layer(D, out_channels, device)
        ada = nn.Module()
        ada.emb = nn.Module()
        ada.emb.timestep_embedder = nn.Module()
        ada.emb.timestep_embedder.linear_1 = _linear(D, self.flow_t_size, device)
        ada.emb.timestep_embedder.linear_2 = _linear(D, D, device)
        ada.linear = _linear(6 * D, D, device)
        self.adaln_single = ada
        self._h = None
        self._keep = []

    # ------------------------------------------------------------------ native handle
    def _destroy(self):
        if self._h is not None:
            _lib.lib().ua2_dit_destroy(self._h)
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
        return self.scale_shift_table.device

    def _ensure(self):
        if self._h is not None:
            return self._h
        L = _lib.lib()
        dev = self.device
        if dev.type != "cuda":
            raise _lib.Ua2Error("uniaudio2_b200 Transformer1DModel runs on a CUDA device only (no CPU fallback): call .to('cuda') first")
        cfg = _lib.DitCfg(self.num_attention_heads, self.attention_head_dim, self.in_channels, self.out_channels, self.num_layers,
                          self.num_positional_embeddings, self.flow_t_size, float(self.norm_eps))
        h = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(L.ua2_dit_create(C.byref(cfg), C.byref(h)), "ua2_dit_create")
            tensors = {k: v.detach() for k, v in self.state_dict().items()}
            half = self.flow_t_size // 2  # freqs of timestep_embedding, evaluated like the reference (transformer_1d_flow.py:67)
            tensors["tfreqs"] = torch.exp(-math.log(10000) * torch.arange(start=0, end=half, device=dev) / half).float()
            keep = []
            for key, t in tensors.items():
                if t.dtype != torch.float32:
                    raise _lib.Ua2Error(f"{key} has dtype {t.dtype}; this path computes in fp32")
                t = t.contiguous()
                keep.append(t)
                shape = (C.c_int64 * t.dim())(*t.shape)
                _lib.check(L.ua2_dit_load_weight(h, key.encode(), _lib.ptr(t), shape, t.dim()), f"load_weight({key})")
            _lib.check(L.ua2_dit_finalize(h, _lib.current_stream()), "ua2_dit_finalize")
        self._h, self._keep = h, keep
        return h

    def set_option(self, name: str, value: int):
        """'bf16' (0/1, default 0): run the many-row linears on bf16 operands with fp32 accumulation, like the reference under
        torch.autocast(bfloat16) (reason_tokenizer.py:265); the default is fp32-class (3xTF32)."""
        _lib.check(_lib.lib().ua2_dit_set_option(self._ensure(), name.encode(), int(value)), f"set_option({name})")

    def last_launch_count(self) -> int:
        return int(_lib.lib().ua2_dit_last_launch_count(self._h)) if self._h is not None else 0

    # ------------------------------------------------------------------ forward
    @torch.inference_mode()
    def forward(self, hidden_states, encoder_hidden_states=None, timestep=None, added_cond_kwargs=None, class_labels=None,
                cross_attention_kwargs=None, attention_mask=None, encoder_attention_mask=None, return_dict: bool = True):
        if encoder_hidden_states is not None or attention_mask is not None or encoder_attention_mask is not None:
            raise NotImplementedError("unmasked self-attention only (every reference call site, AudioDiffusion1D.py:108-121)")
        h = self._ensure()
        dev = self.device
        if hidden_states.dim() != 3 or hidden_states.shape[2] != self.in_channels:
            raise ValueError(f"expected hidden_states of shape (B, T, {self.in_channels})")
        x = hidden_states.to(device=dev, dtype=torch.float32).contiguous()
        B, T, _ = x.shape
        t = torch.as_tensor(timestep, device=dev, dtype=torch.float32).reshape(-1).contiguous()
        if t.numel() != B:
            raise RuntimeError(f"timestep has {t.numel()} entries for batch {B}")  # the reshape(batch, 6, -1) failure of attention.py:314
        out = torch.empty(B, T, self.out_channels, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().ua2_dit_forward(h, _lib.ptr(x), _lib.ptr(t), _lib.ptr(out), B, T, _lib.current_stream()), "forward")
        return Transformer1DModelOutput(sample=out) if return_dict else (out,)
