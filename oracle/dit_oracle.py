"""CPU oracle for the flow-matching decoder of ReasoningCodec_film: the DiT estimator
`Transformer1DModel` and the Euler solver `BASECFM.solve_euler`.  TEST INFRASTRUCTURE ONLY - never imported by the product path.

Restated as pure torch-CPU fp32 functions over a flat state dict with the reference's own key names (paths relative to
/root/reference/tools/tokenizer/ReasoningCodec_film/models/):
    Transformer1DModel.forward                 transformer_1d_flow.py:284-386
    ProjectLayer.forward                       transformer_1d_flow.py:19-34      (Conv1d k3 'same' -> * k^-0.5 -> Linear)
    PixArtAlphaCombinedFlowEmbeddings          transformer_1d_flow.py:37-84      (sinusoidal(512) * 1000 -> TimestepEmbedding)
    AdaLayerNormSingleFlow.forward             transformer_1d_flow.py:87-117     (linear(silu(emb)) -> 6 * dim)
    BasicTransformerBlock.forward              attention.py:284-418              (ada_norm_single branch)
    FeedForward ('gelu-approximate')           attention.py:623-681
    BASECFM.solve_euler                        AudioDiffusion1D.py:89-129
and, from the un-vendored dependency diffusers (pyproject.toml:38 `diffusers>=0.25.0`), its published algorithms for
    Attention (AttnProcessor2_0), GELU(approximate='tanh'), TimestepEmbedding, SinusoidalPositionalEmbedding.

Parity status: the in-repo parts are PINNED - oracle/make_golden_dit.py imports the unmodified transformer_1d_flow.py and
attention.py (over oracle/diffusers_stub.py) and executes the unmodified source of class BASECFM, and asserts this
restatement is bit-identical to them on CPU; fixtures in tests/golden/dit_golden.pt.  The four diffusers leaves are
"parity UNPINNED": diffusers is not installed here, so they are checked only against their restatement in the stub.
The production checkpoint and sqcodec/DiT weights are not in the repository: all tests use seeded random weights.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict

import torch
import torch.nn.functional as F


@dataclass
class DitCfg:
    """model_config.json of the reference (defaults = the production values)."""

    num_attention_heads: int = 24
    attention_head_dim: int = 64
    in_channels: int = 1040
    out_channels: int = 136
    num_layers: int = 32
    norm_eps: float = 1e-6
    num_positional_embeddings: int = 3000  # Transformer1DModel default, transformer_1d_flow.py:195
    flow_t_size: int = 512                 # transformer_1d_flow.py:47

    @property
    def inner_dim(self):
        return self.num_attention_heads * self.attention_head_dim

    def ctor_kwargs(self):
        return dict(num_attention_heads=self.num_attention_heads, attention_head_dim=self.attention_head_dim,
                    in_channels=self.in_channels, out_channels=self.out_channels, num_layers=self.num_layers,
                    attention_bias=True, activation_fn="gelu-approximate", norm_type="ada_norm_single",
                    norm_elementwise_affine=False, norm_eps=self.norm_eps, num_embeds_ada_norm=1000,
                    num_positional_embeddings=self.num_positional_embeddings)


def state_dict_shapes(cfg: DitCfg) -> Dict[str, tuple]:
    D, I, O = cfg.inner_dim, cfg.in_channels, cfg.out_channels
    out = {"scale_shift_table": (2, D), "proj_in.ffn_1.weight": (D, I, 3), "proj_in.ffn_1.bias": (D,),
           "proj_in.ffn_2.weight": (D, D), "proj_in.ffn_2.bias": (D,), "pos_embed.pe": (1, cfg.num_positional_embeddings, D)}
    for i in range(cfg.num_layers):
        p = f"transformer_blocks.{i}."
        out[p + "scale_shift_table"] = (6, D)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            out[p + f"attn1.{n}.weight"] = (D, D)
            out[p + f"attn1.{n}.bias"] = (D,)
        out[p + "ff.net.0.proj.weight"] = (4 * D, D)
        out[p + "ff.net.0.proj.bias"] = (4 * D,)
        out[p + "ff.net.2.weight"] = (D, 4 * D)
        out[p + "ff.net.2.bias"] = (D,)
    out.update({"proj_out.ffn_1.weight": (O, D, 3), "proj_out.ffn_1.bias": (O,), "proj_out.ffn_2.weight": (O, O),
                "proj_out.ffn_2.bias": (O,),
                "adaln_single.emb.timestep_embedder.linear_1.weight": (D, cfg.flow_t_size),
                "adaln_single.emb.timestep_embedder.linear_1.bias": (D,),
                "adaln_single.emb.timestep_embedder.linear_2.weight": (D, D),
                "adaln_single.emb.timestep_embedder.linear_2.bias": (D,),
                "adaln_single.linear.weight": (6 * D, D), "adaln_single.linear.bias": (6 * D,)})
    return out


def sinusoidal_pe(embed_dim: int, max_seq_length: int) -> torch.Tensor:
    """diffusers SinusoidalPositionalEmbedding buffer `pe` (1, max_seq_length, embed_dim)."""
    position = torch.arange(max_seq_length).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, embed_dim, 2) * (-math.log(10000.0) / embed_dim))
    pe = torch.zeros(1, max_seq_length, embed_dim)
    pe[0, :, 0::2] = torch.sin(position * div_term)
    pe[0, :, 1::2] = torch.cos(position * div_term)
    return pe


def random_state_dict(cfg: DitCfg, seed: int) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in state_dict_shapes(cfg).items():
        if k == "pos_embed.pe":
            t = sinusoidal_pe(cfg.inner_dim, cfg.num_positional_embeddings)
        elif k.endswith("scale_shift_table"):
            t = torch.randn(shp, generator=g) / cfg.inner_dim ** 0.5  # the reference's init, transformer_1d_flow.py:235
        elif k.endswith(".bias"):
            t = 0.05 * torch.randn(shp, generator=g)
        else:
            fan_in = shp[1] * (shp[2] if len(shp) == 3 else 1)
            t = torch.randn(shp, generator=g) / math.sqrt(fan_in)
        sd[k] = t.float().contiguous()
    return sd


def project_layer(sd, prefix, x):
    """ProjectLayer.forward, transformer_1d_flow.py:28-34 (kernel_size = 3)."""
    x = F.conv1d(x.transpose(1, 2), sd[prefix + ".ffn_1.weight"], sd[prefix + ".ffn_1.bias"], padding=1).transpose(1, 2)
    x = x * 3 ** -0.5
    return F.linear(x, sd[prefix + ".ffn_2.weight"], sd[prefix + ".ffn_2.bias"])


def timestep_embedding(timesteps, flow_t_size=512, max_period=10000, scale=1000):
    """PixArtAlphaCombinedFlowEmbeddings.timestep_embedding, transformer_1d_flow.py:58-71."""
    half = flow_t_size // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half) / half).type(timesteps.type())
    args = timesteps[:, None] * freqs[None] * scale
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


class DitOracle:
    def __init__(self, cfg: DitCfg, sd: Dict[str, torch.Tensor]):
        self.cfg, self.sd = cfg, sd

    def adaln_single(self, timestep):
        """AdaLayerNormSingleFlow.forward -> (6*dim modulation, embedded_timestep), transformer_1d_flow.py:106-117."""
        sd, p = self.sd, "adaln_single.emb.timestep_embedder."
        proj = timestep_embedding(timestep, self.cfg.flow_t_size)
        emb = F.linear(F.silu(F.linear(proj, sd[p + "linear_1.weight"], sd[p + "linear_1.bias"])), sd[p + "linear_2.weight"],
                       sd[p + "linear_2.bias"])
        return F.linear(F.silu(emb), sd["adaln_single.linear.weight"], sd["adaln_single.linear.bias"]), emb

    def block(self, i, h, t6):
        """BasicTransformerBlock.forward, ada_norm_single branch (attention.py:308-415)."""
        c, sd, p = self.cfg, self.sd, f"transformer_blocks.{i}."
        B, T, D = h.shape
        shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = (sd[p + "scale_shift_table"][None] + t6.reshape(B, 6, -1)).chunk(6, dim=1)
        n = F.layer_norm(h, (D,), None, None, c.norm_eps)
        n = n * (1 + scale_msa) + shift_msa
        n = n.squeeze(1)
        q = F.linear(n, sd[p + "attn1.to_q.weight"], sd[p + "attn1.to_q.bias"])
        k = F.linear(n, sd[p + "attn1.to_k.weight"], sd[p + "attn1.to_k.bias"])
        v = F.linear(n, sd[p + "attn1.to_v.weight"], sd[p + "attn1.to_v.bias"])
        H, hd = c.num_attention_heads, c.attention_head_dim
        q, k, v = (t.view(B, -1, H, hd).transpose(1, 2) for t in (q, k, v))
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(B, -1, D)
        o = F.linear(o, sd[p + "attn1.to_out.0.weight"], sd[p + "attn1.to_out.0.bias"])
        h = gate_msa * o + h
        n = F.layer_norm(h, (D,), None, None, c.norm_eps)
        n = n * (1 + scale_mlp) + shift_mlp
        f = F.gelu(F.linear(n, sd[p + "ff.net.0.proj.weight"], sd[p + "ff.net.0.proj.bias"]), approximate="tanh")
        f = F.linear(f, sd[p + "ff.net.2.weight"], sd[p + "ff.net.2.bias"])
        return gate_mlp * f + h

    def forward(self, hidden_states, timestep):
        """Transformer1DModel.forward(...).sample: (B, T, in_channels), (B,) -> (B, T, out_channels)."""
        c, sd = self.cfg, self.sd
        h = project_layer(sd, "proj_in", hidden_states)
        h = h + sd["pos_embed.pe"][:, : h.shape[1]]
        t6, emb = self.adaln_single(timestep)
        for i in range(c.num_layers):
            h = self.block(i, h, t6)
        shift, scale = (sd["scale_shift_table"][None] + emb[:, None]).chunk(2, dim=1)
        h = F.layer_norm(h, (c.inner_dim,), None, None, 1e-6)
        h = h * (1 + scale) + shift
        return project_layer(sd, "proj_out", h)

    def solve_euler(self, x, incontext_x, incontext_length, t_span, mu, guidance_scale, sigma_min=1e-4):
        """BASECFM.solve_euler, AudioDiffusion1D.py:89-129 (x is updated in place on the in-context rows, like the reference)."""
        t, dt = t_span[0], t_span[1] - t_span[0]
        noise = x.clone()
        for step in range(1, len(t_span)):
            x[:, 0:incontext_length, :] = (1 - (1 - sigma_min) * t) * noise[:, 0:incontext_length, :] + t * incontext_x[:, 0:incontext_length, :]
            if guidance_scale > 1.0:
                inp = torch.cat([torch.cat([x, x], 0), torch.cat([incontext_x, incontext_x], 0),
                                 torch.cat([torch.zeros_like(mu), mu], 0)], 2)
                d = self.forward(inp, t.unsqueeze(-1).repeat(2))
                d_uncond, d_cond = d.chunk(2, 0)
                d = d_uncond + guidance_scale * (d_cond - d_uncond)
            else:
                # the reference's branch concatenates along TIME here (AudioDiffusion1D.py:119, dim 1), which cannot match
                # in_channels; every caller passes guidance_scale = 1.5 (reason_tokenizer.py:273-283)
                raise ValueError("solve_euler is served with classifier-free guidance (guidance_scale > 1) only")
            x = x + dt * d
            t = t + dt
            if step < len(t_span) - 1:
                dt = t_span[step + 1] - t
        return x
