"""Drop-in for the inference surface of `AudioThinking` (tools/tokenizer/ReasoningCodec_film/models/AudioDiffusion1D.py:169-188) - the
reasoning encoder of tokenize - and of the method that drives it, AudioDiffusion1D.encode_reasoning_part (:372-390):

    quantized_features, indices, commitment_loss = self.encode_reasoning_part(whisper_embeds (B, 1024, 1500), muencoder_embeds (B, 1024, 750))

Sub-module and parameter names follow the reference so that its checkpoint keys load: cls_token, down_sampling_layer_whisper.*,
semantic_merge_proj.*, encoder_transformers.{i}.{self_attn.to_qkv / to_out (weight-normed: parametrizations.weight.original0 / original1),
self_attn.q_norm / k_norm, self_attn_scale.scale, ff.ff.0.proj (weight-normed, bias), ff.ff.2 (weight-normed, bias), ff_scale.scale,
rope.inv_freq}, reasoning_vq.* (ResidualVQ: dim -> 64, 8 x 4096 codes).  The LLM head that turns the query tokens into reasoning TEXT
(return_reasoning_text=True, llama tokenizer + LoRA model) is not part of this path.

Arithmetic: libua2_b200.so (csrc/ua2_thinking.cu for the encoder, ua2_linear_bias_f32 / ua2_rvq_encode_gemm_f32 for the quantiser).
The weight normalisation g * v / ||v|| is folded once when the native handle is built.  No torch / CPU fallback."""
import ctypes as C
import math

import torch
import torch.nn as nn

from ..... import _lib
from .modeling_whisper import _P


def _weight_normed(n_out, n_in, device, bias):
    """nn.utils.parametrizations.weight_norm(nn.Linear(n_in, n_out)) as the state dict sees it."""
    m = nn.Module()
    v = torch.empty(n_out, n_in, device=device)
    nn.init.kaiming_uniform_(v, a=math.sqrt(5))
    m.parametrizations = nn.Module()
    m.parametrizations.weight = _P(original0=v.norm(dim=1, keepdim=True), original1=v)
    if bias:
        m.bias = nn.Parameter(torch.zeros(n_out, device=device), requires_grad=False)
    return m


class AudioThinking(nn.Module):
    def __init__(self, dim=768, interval=5, encoder_depth=5, whisper_fea_dim=1024, mu_dim=1024, dim_heads=128, ff_mult=4, device=None):
        super().__init__()
        from .AudioDiffusion1D import ResidualVQ

        if dim_heads != 128:
            raise NotImplementedError("dim_heads is 128 in the reference (AudioDiffusion1D.py:178)")
        self.dim, self.interval, self.whisper_fea_dim, self.mu_dim, self.dim_heads, self.ff_mult = dim, interval, whisper_fea_dim, mu_dim, dim_heads, ff_mult
        self.cls_token = nn.Parameter(torch.randn(1, dim, device=device), requires_grad=False)
        blocks = nn.ModuleList()
        rot = max(dim_heads // 2, 32)
        for _ in range(encoder_depth):
            b = nn.Module()
            b.self_attn = nn.Module()
            b.self_attn.to_qkv = _weight_normed(3 * dim, dim, device, bias=False)
            b.self_attn.to_out = _weight_normed(dim, dim, device, bias=False)
            b.self_attn.q_norm = _P(weight=torch.ones(dim_heads, device=device), bias=torch.zeros(dim_heads, device=device))
            b.self_attn.k_norm = _P(weight=torch.ones(dim_heads, device=device), bias=torch.zeros(dim_heads, device=device))
            b.self_attn_scale = _P(scale=torch.full([dim], 1e-2, device=device))
            b.ff = nn.Module()
            b.ff.ff = nn.ModuleList([nn.Module(), nn.Identity(), _weight_normed(dim, dim * ff_mult, device, bias=True), nn.Identity()])
            b.ff.ff[0].proj = _weight_normed(2 * dim * ff_mult, dim, device, bias=True)
            b.ff_scale = _P(scale=torch.full([dim], 1e-2, device=device))
            b.rope = nn.Module()
            b.rope.register_buffer("inv_freq", (1.0 / (10000 ** (torch.arange(0, rot, 2).float() / rot))).to(device))
            blocks.append(b)
        self.encoder_transformers = blocks
        self.semantic_merge_proj = _P(weight=torch.empty(dim, whisper_fea_dim + mu_dim, device=device).uniform_(-0.02, 0.02), bias=torch.zeros(dim, device=device))
        self.reasoning_vq = ResidualVQ(dim, 4096, 64, 8, device=device)
        self.down_sampling_layer_whisper = _P(weight=torch.empty(whisper_fea_dim, whisper_fea_dim, 2, device=device).uniform_(-0.02, 0.02),
                                              bias=torch.zeros(whisper_fea_dim, device=device))
        self._h = None
        self._keep = []

    # ------------------------------------------------------------------ native handle
    def _destroy(self):
        if getattr(self, "_h", None) is not None:
            _lib.lib().ua2_thinking_destroy(self._h)
            self._h = None
            self._keep = []

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    def load_state_dict(self, sd, strict=True, **kw):
        self._destroy()
        own = {k: v for k, v in sd.items() if not k.startswith("reasoning_vq.")}
        vq = {k[len("reasoning_vq."):]: v for k, v in sd.items() if k.startswith("reasoning_vq.")}
        if vq:
            self.reasoning_vq.load_state_dict(vq, strict=strict)
        own.update({"reasoning_vq." + k: v for k, v in self.reasoning_vq.state_dict().items()})
        return super().load_state_dict(own, strict=strict, **kw)

    def _apply(self, fn, *a, **kw):
        self._destroy()
        return super()._apply(fn, *a, **kw)

    @property
    def device(self):
        return self.cls_token.device

    def _ensure(self):
        if self._h is not None:
            return self._h
        L = _lib.lib()
        dev = self.device
        if dev.type != "cuda":
            raise _lib.Ua2Error("uniaudio2_b200 AudioThinking runs on a CUDA device only (no CPU fallback): call .to('cuda') first")
        cfg = _lib.ThinkingCfg(self.dim, self.dim_heads, len(self.encoder_transformers), self.interval, self.whisper_fea_dim, self.mu_dim, self.ff_mult)
        h = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(L.ua2_thinking_create(C.byref(cfg), C.byref(h)), "ua2_thinking_create")
            keep = []
            try:
                sd = {k: t.detach() for k, t in self.state_dict().items() if not k.startswith("reasoning_vq.")}
                folded = {}
                for k, t in sd.items():
                    if k.endswith(".parametrizations.weight.original0"):
                        pre = k[:-len(".parametrizations.weight.original0")]
                        v = sd[pre + ".parametrizations.weight.original1"]
                        folded[pre + ".weight"] = t * (v / v.norm(dim=1, keepdim=True))  # weight_norm, dim = 0
                    elif not k.endswith(".parametrizations.weight.original1"):
                        folded[k] = t
                for key, t in folded.items():
                    if t.dtype != torch.float32:
                        raise _lib.Ua2Error(f"{key} has dtype {t.dtype}; this path takes fp32 parameters")
                    t = t.contiguous()
                    keep.append(t)
                    shape = (C.c_int64 * t.dim())(*t.shape)
                    _lib.check(L.ua2_thinking_load_weight(h, key.encode(), _lib.ptr(t), shape, t.dim()), f"load_weight({key})")
                _lib.check(L.ua2_thinking_finalize(h, _lib.current_stream()), "ua2_thinking_finalize")
            except Exception:
                L.ua2_thinking_destroy(h)
                raise
        self._h, self._keep = h, keep
        return h

    def last_launch_count(self) -> int:
        return int(_lib.lib().ua2_thinking_last_launch_count(self._h)) if self._h is not None else 0

    # ------------------------------------------------------------------ forward
    @torch.inference_mode()
    def query_tokens(self, whisper_embeds, muencoder_embeds):
        """AudioDiffusion1D.encode_reasoning_part (:372-387) up to `query_tokens`: (B, whisper_dim, Tw), (B, mu_dim, Tb) -> (B, T / interval, dim)."""
        if whisper_embeds.dim() != 3 or muencoder_embeds.dim() != 3 or whisper_embeds.shape[1] != self.whisper_fea_dim or muencoder_embeds.shape[1] != self.mu_dim:
            raise ValueError(f"expected (B, {self.whisper_fea_dim}, Tw) and (B, {self.mu_dim}, Tb), got {tuple(whisper_embeds.shape)} and {tuple(muencoder_embeds.shape)}")
        h = self._ensure()
        dev = self.device
        w = whisper_embeds.to(device=dev, dtype=torch.float32).contiguous()
        mu = muencoder_embeds.to(device=dev, dtype=torch.float32).contiguous()
        B, _, Tw = w.shape
        Tb = mu.shape[-1]
        rows = int(_lib.lib().ua2_thinking_rows(h, Tw, Tb))
        if rows <= 0:  # the reference's set_masking fails in its reshape
            raise RuntimeError(f"min(Tw // 2, Tb) = {min(Tw // 2, Tb)} frames cannot be split into groups of {self.interval}")
        out = torch.empty(B, rows, self.dim, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().ua2_thinking_encode(h, _lib.ptr(w), _lib.ptr(mu), B, Tw, Tb, _lib.ptr(out), _lib.current_stream()), "ua2_thinking_encode")
        return out[:, self.interval::self.interval + 1]  # extract_mask_positions (:479-486)

    @torch.inference_mode()
    def encode_reasoning_part(self, whisper_embeds, muencoder_embeds):
        """-> (quantized_features (B, Tq, dim), indices (B, Tq, 8), commitment_loss = None in inference)."""
        q = self.query_tokens(whisper_embeds, muencoder_embeds).contiguous()
        B, Tq, _ = q.shape
        codes = torch.empty(B, self.reasoning_vq.num_quantizers, Tq, device=q.device, dtype=torch.int64)
        quantized = self.reasoning_vq.encode(q, codes, 0)
        return quantized, codes.transpose(1, 2), None
