"""TEST INFRASTRUCTURE ONLY - never imported by the product path.

Minimal stand-in for the `diffusers` package (pyproject.toml:38 pins `diffusers>=0.25.0`; model_config.json was written by
0.22.0.dev0; the package is NOT installed in this image and not vendored under /root/reference), just enough to import the
UNMODIFIED in-repo files of the flow-matching decoder:

    tools/tokenizer/ReasoningCodec_film/models/transformer_1d_flow.py   (Transformer1DModel, ProjectLayer, adaLN-single)
    tools/tokenizer/ReasoningCodec_film/models/attention.py             (BasicTransformerBlock, FeedForward)

The leaf classes those files take from diffusers are RESTATED here from the published diffusers 0.25 sources
(src/diffusers/models/{attention_processor,activations,embeddings}.py):
    Attention (AttnProcessor2_0 path: to_q/to_k/to_v -> heads -> F.scaled_dot_product_attention -> to_out[0] -> dropout)
    GELU(dim_in, dim_out, approximate, bias)          proj + F.gelu(approximate=...)
    TimestepEmbedding(in_channels, time_embed_dim)    linear_1 -> SiLU -> linear_2
    SinusoidalPositionalEmbedding(embed_dim, max_seq_length)   pe[0,:,0::2] = sin, pe[0,:,1::2] = cos, x + pe[:, :T]
Parity of these four leaves is therefore UNPINNED (no diffusers here to check them against); everything above them - the
block wiring, adaLN-single modulation, ProjectLayer, timestep embedding, the Euler solver - is the reference's own code.
Classes the imported files merely name (GEGLU, AdaLayerNorm, ...) are empty placeholders that raise if constructed.
"""
import math
import sys
import types

import torch
import torch.nn.functional as F
from torch import nn


class SinusoidalPositionalEmbedding(nn.Module):
    """diffusers/models/embeddings.py (0.25): additive sin/cos table, interleaved even = sin, odd = cos."""

    def __init__(self, embed_dim: int, max_seq_length: int = 32):
        super().__init__()
        position = torch.arange(max_seq_length).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, embed_dim, 2) * (-math.log(10000.0) / embed_dim))
        pe = torch.zeros(1, max_seq_length, embed_dim)
        pe[0, :, 0::2] = torch.sin(position * div_term)
        pe[0, :, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe)

    def forward(self, x):
        _, seq_length, _ = x.shape
        return x + self.pe[:, :seq_length]


class TimestepEmbedding(nn.Module):
    """diffusers/models/embeddings.py (0.25) with its defaults: act_fn='silu', no cond_proj, no post_act."""

    def __init__(self, in_channels: int, time_embed_dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, sample, condition=None):
        return self.linear_2(self.act(self.linear_1(sample)))


class GELU(nn.Module):
    """diffusers/models/activations.py (0.25): Linear + GELU (tanh approximation when approximate='tanh')."""

    def __init__(self, dim_in: int, dim_out: int, approximate: str = "none", bias: bool = True):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out, bias=bias)
        self.approximate = approximate

    def forward(self, hidden_states):
        return F.gelu(self.proj(hidden_states), approximate=self.approximate)


class Attention(nn.Module):
    """diffusers/models/attention_processor.py (0.25), self-attention configuration used by BasicTransformerBlock.attn1
    (no cross attention, no group / qk norm, scale = dim_head ** -0.5, AttnProcessor2_0)."""

    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, dropout=0.0, bias=False, upcast_attention=False,
                 out_bias=True, **unused):
        super().__init__()
        assert cross_attention_dim is None, "stub: self-attention only"
        self.inner_dim = dim_head * heads
        self.heads = heads
        self.to_q = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_k = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_v = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(self.inner_dim, query_dim, bias=out_bias), nn.Dropout(dropout)])

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **unused):
        assert encoder_hidden_states is None and attention_mask is None, "stub: unmasked self-attention only"
        B, T, _ = hidden_states.shape
        q, k, v = self.to_q(hidden_states), self.to_k(hidden_states), self.to_v(hidden_states)
        hd = self.inner_dim // self.heads
        q = q.view(B, -1, self.heads, hd).transpose(1, 2)
        k = k.view(B, -1, self.heads, hd).transpose(1, 2)
        v = v.view(B, -1, self.heads, hd).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(B, -1, self.inner_dim).to(q.dtype)
        return self.to_out[1](self.to_out[0](o))


def _placeholder(name):
    class _Missing(nn.Module):
        def __init__(self, *a, **k):
            raise NotImplementedError(f"diffusers stub: {name} is not restated (not on the ada_norm_single path)")

    _Missing.__name__ = name
    return _Missing


class BaseOutput:
    pass


def register_to_config(init):
    return init


def install_diffusers_stub():
    """Register the stand-in under sys.modules['diffusers...'] (generator scripts only)."""

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        if "." not in name or name.count(".") < 2:
            m.__path__ = []
        sys.modules[name] = m
        return m

    mod("diffusers", __version__="0.25.0-stub")
    mod("diffusers.configuration_utils", ConfigMixin=object, register_to_config=register_to_config)
    mod("diffusers.utils", USE_PEFT_BACKEND=True, BaseOutput=BaseOutput, deprecate=lambda *a, **k: None,
        is_torch_version=lambda *a, **k: True)
    mod("diffusers.utils.torch_utils", maybe_allow_in_graph=lambda cls: cls)
    models = mod("diffusers.models")
    models.__path__ = []
    mod("diffusers.models.embeddings", ImagePositionalEmbeddings=_placeholder("ImagePositionalEmbeddings"),
        PatchEmbed=_placeholder("PatchEmbed"), PixArtAlphaTextProjection=_placeholder("PixArtAlphaTextProjection"),
        TimestepEmbedding=TimestepEmbedding, SinusoidalPositionalEmbedding=SinusoidalPositionalEmbedding)
    mod("diffusers.models.lora", LoRACompatibleConv=nn.Conv2d, LoRACompatibleLinear=nn.Linear)
    mod("diffusers.models.modeling_utils", ModelMixin=nn.Module)
    mod("diffusers.models.activations", GEGLU=_placeholder("GEGLU"), GELU=GELU, ApproximateGELU=_placeholder("ApproximateGELU"))
    mod("diffusers.models.attention_processor", Attention=Attention)
    mod("diffusers.models.normalization", AdaLayerNorm=_placeholder("AdaLayerNorm"),
        AdaLayerNormContinuous=_placeholder("AdaLayerNormContinuous"), AdaLayerNormZero=_placeholder("AdaLayerNormZero"),
        RMSNorm=_placeholder("RMSNorm"))
