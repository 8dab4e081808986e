"""CPU oracle for the Moshi-family streaming transformer and sampler that BASELINE.json's north_star names
(llm_modules/transformer.py + gating.py + rope.py, driven by llm_utils/sampling.py).
TEST INFRASTRUCTURE ONLY - never imported by the product path.

Restated as pure torch-CPU fp32 functions over a flat state dict with the reference's own key names
(paths relative to /root/reference):
    StreamingTransformer.forward            llm_modules/transformer.py:671-692   (sin embedding :126-152)
    StreamingTransformerLayer.forward       llm_modules/transformer.py:545-588   (_sa_block, _ff_block)
    StreamingMultiheadAttention.forward     llm_modules/transformer.py:375-419
    RingKVCache.complete                    llm_modules/transformer.py:242-278
    multi_linear                            llm_modules/transformer.py:155-179
    _rms_norm / create_norm_fn              llm_modules/transformer.py:34-46, :101-123
    ActivationGating / gating_forward_kernel  llm_modules/gating.py:12-51
    apply_rope (interleaved pairs)          llm_modules/rope.py:11-68
    sample_token / sample_top_k / sample_top_p / multinomial   llm_utils/sampling.py:15-105

Parity status: PINNED.  oracle/make_golden_moshi.py imports the unmodified reference modules in the build container
(alias import `modules` -> llm_modules, `utils.compile` -> llm_utils.compile) and asserts this
restatement is bit-identical to them on CPU, streaming and non-streaming; fixtures live in tests/golden/moshi_golden.pt.
The reference's own known-answer test for the sampler (llm_utils/sampling.py:156-174, multinomial frequencies) is
re-stated in tests/test_moshi_oracle.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Union

import torch
import torch.nn.functional as F


@dataclass
class StxCfg:
    """Constructor arguments of StreamingTransformer + StreamingTransformerLayer (transformer.py:616-669, :449-470)."""

    d_model: int = 64
    num_heads: int = 2
    num_layers: int = 2
    dim_feedforward: Union[int, List[int]] = 256
    causal: bool = True
    context: Optional[int] = None
    positional_embedding: str = "rope"  # sin | rope | sin_rope | none
    max_period: float = 10000.0
    positional_scale: float = 1.0
    norm: str = "layer_norm"  # layer_norm | layer_norm_f32 | rms_norm | rms_norm_f32
    layer_scale: Optional[float] = None
    gating: str = "none"  # none | silu (any name of gating.py:54-62 in the oracle)
    weights_per_step: int = 0

    def ff_dims(self) -> List[int]:
        if isinstance(self.dim_feedforward, list):
            return list(self.dim_feedforward)
        return [self.dim_feedforward] * max(1, self.weights_per_step)

    def gating_hidden(self, dim_ff: int) -> int:
        """gating.py:40-43."""
        return (21 * self.d_model) // 8 if dim_ff == 4 * self.d_model else (2 * dim_ff) // 3

    def norm_eps(self) -> float:
        """create_norm_fn, transformer.py:111-121."""
        return 1e-5 if self.norm in ("layer_norm", "rms_norm") else 1e-8

    def capacity(self) -> int:
        """_init_streaming_state, transformer.py:337-346."""
        if self.context is None:
            if self.weights_per_step:
                return self.weights_per_step
            raise RuntimeError("Cannot create a streaming KVCache without a context to estimate capacity.")
        return self.context


def state_dict_shapes(cfg: StxCfg) -> Dict[str, tuple]:
    """Key -> shape of StreamingTransformer(**cfg).state_dict() in the reference."""
    D = cfg.d_model
    mult = cfg.weights_per_step if cfg.weights_per_step else 1
    out: Dict[str, tuple] = {}
    for i in range(cfg.num_layers):
        p = f"layers.{i}."
        out[p + "self_attn.in_proj_weight"] = (mult * 3 * D, D)
        out[p + "self_attn.out_proj.weight"] = (mult * D, D)
        for n in ("norm1", "norm2"):
            if cfg.norm.startswith("layer_norm"):
                out[p + n + ".weight"] = (D,)
                out[p + n + ".bias"] = (D,)
            else:
                out[p + n + ".alpha"] = (1, 1, D)
        if cfg.gating == "none":
            ff = cfg.dim_feedforward
            assert isinstance(ff, int) and not cfg.weights_per_step
            out[p + "linear1.weight"] = (ff, D)
            out[p + "linear2.weight"] = (D, ff)
        elif cfg.weights_per_step:
            for s, ff in enumerate(cfg.ff_dims()):
                h = cfg.gating_hidden(ff)
                out[p + f"gating.{s}.linear_in.weight"] = (2 * h, D)
                out[p + f"gating.{s}.linear_out.weight"] = (D, h)
        else:
            h = cfg.gating_hidden(cfg.ff_dims()[0])
            out[p + "gating.linear_in.weight"] = (2 * h, D)
            out[p + "gating.linear_out.weight"] = (D, h)
        if cfg.layer_scale is not None:
            out[p + "layer_scale_1.scale"] = (D,)
            out[p + "layer_scale_2.scale"] = (D,)
    return out


def random_state_dict(cfg: StxCfg, seed: int) -> Dict[str, torch.Tensor]:
    """Seeded fp32 weights with every key of the reference module (norm affine parameters perturbed so that they matter)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in state_dict_shapes(cfg).items():
        if k.endswith("in_proj_weight") or k.endswith("linear_in.weight") or k.endswith("linear1.weight"):
            t = torch.randn(shp, generator=g) / math.sqrt(shp[1])
        elif k.endswith("out_proj.weight") or k.endswith("linear_out.weight") or k.endswith("linear2.weight"):
            t = torch.randn(shp, generator=g) / math.sqrt(shp[1])
        elif k.endswith(".bias"):
            t = 0.1 * torch.randn(shp, generator=g)
        elif k.endswith(".scale"):
            t = (cfg.layer_scale or 1.0) * (1.0 + 0.5 * torch.randn(shp, generator=g))
        else:  # norm weight / alpha
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        sd[k] = t.float().contiguous()
    return sd


# --------------------------------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------------------------------
def rms_norm(x, alpha, eps, dtype=None):
    """_rms_norm, transformer.py:34-46 (dtype=torch.float for rms_norm_f32; a no-op on fp32 inputs)."""
    x_dtype = x.dtype
    if dtype is not None:
        x = x.to(dtype)
    var = eps + torch.mean(x**2, dim=2, keepdim=True)
    return (x * (alpha.to(var) * torch.rsqrt(var))).to(x_dtype)


def create_sin_embedding(positions, dim, max_period=10000.0, dtype=torch.float32):
    """transformer.py:126-152."""
    half_dim = dim // 2
    positions = positions.to(dtype)
    adim = torch.arange(half_dim, dtype=dtype).view(1, 1, -1)
    max_period_tensor = torch.full([], max_period, dtype=dtype)
    phase = positions / (max_period_tensor ** (adim / (half_dim - 1)))
    return torch.cat([torch.cos(phase), torch.sin(phase)], dim=-1)


def sin_embedding_denominators(dim: int, max_period: float) -> torch.Tensor:
    """The (half_dim,) divisor table of create_sin_embedding, exactly as the reference evaluates it."""
    half_dim = dim // 2
    adim = torch.arange(half_dim, dtype=torch.float32)
    return torch.full([], max_period, dtype=torch.float32) ** (adim / (half_dim - 1))


def rope_freqs(head_dim: int, max_period: float) -> torch.Tensor:
    """freqs of apply_rope, rope.py:37-38."""
    ds = torch.arange(head_dim // 2, dtype=torch.float32)
    return torch.exp(ds * (-math.log(max_period) * 2 / head_dim))


def apply_rope(q, k, offset, max_period):
    """rope.py:11-68 with time_before_heads=False: q, k (B, H, T, D); interleaved (real, imaginary) pairs."""
    B, H, T, D = q.shape
    freqs = rope_freqs(D, max_period)
    ts = offset.float() + torch.arange(T, dtype=torch.float32)
    ts = ts.view(1, -1, 1)
    dims = q.shape[:-1]
    q = q.view(*dims, D // 2, 2)
    k = k.view(*dims, D // 2, 2)
    qr, qi = q[..., 0].float(), q[..., 1].float()
    kr, ki = k[..., 0].float(), k[..., 1].float()
    rotr = torch.cos(freqs * ts)
    roti = torch.sin(freqs * ts)
    qor = qr * rotr - qi * roti
    qoi = qr * roti + qi * rotr
    kor = kr * rotr - ki * roti
    koi = kr * roti + ki * rotr
    qo = torch.stack([qor, qoi], dim=-1)
    ko = torch.stack([kor, koi], dim=-1)
    return qo.view(*dims, D), ko.view(*dims, D)


def multi_linear(num_linear, weight, x, offset):
    """transformer.py:155-179: a different weight slab per time step."""
    B, T, C = x.shape
    chout, chin = weight.shape
    weight = weight.view(num_linear, -1, chin)
    ys = [F.linear(x[:, t], weight[t + offset]) for t in range(T)]
    return torch.stack(ys, 1)


def ring_positions(capacity: int, end_offset: int) -> torch.Tensor:
    """Position held by every slot of the ring AFTER `end_offset` keys have been written (transformer.py:254-276);
    -1 = never written.  Note the reference's `delta <= 0` branch: once the ring has wrapped, the slot at
    end_offset % capacity reports position `end_offset` (a future position), so the oldest key is never visible."""
    indexes = torch.arange(capacity, dtype=torch.long)
    invalid = indexes >= end_offset
    end_index = end_offset % capacity
    delta = indexes - end_index
    positions = torch.where(delta <= 0, end_offset + delta, end_offset + delta - capacity)
    return torch.where(invalid, torch.full_like(positions, -1), positions)


class RingKV:
    """RingKVCache, transformer.py:212-278 (fp32 cache here: the reference takes the dtype of in_proj_weight, :343)."""

    def __init__(self, B, H, hd, capacity):
        self.capacity = capacity
        self.cache = torch.zeros(2, B, H, capacity, hd)
        self.end_offset = 0

    def reset(self):
        self.end_offset = 0

    def complete(self, k, v):
        T = k.shape[2]
        idx = (torch.arange(T) + self.end_offset) % self.capacity
        self.cache[0].index_copy_(2, idx, k)
        self.cache[1].index_copy_(2, idx, v)
        self.end_offset += T
        return self.cache[0], self.cache[1], ring_positions(self.capacity, self.end_offset)


_ACT = {"silu": F.silu, "gelu": F.gelu, "relu": torch.relu, "sigmoid": torch.sigmoid, "tanh": torch.tanh,
        "elu": F.elu, "leaky_relu": F.leaky_relu, "mish": F.mish, "softsign": F.softsign}


def gating_forward(w_in, w_out, act, x):
    """gating_forward_kernel, gating.py:12-21."""
    x = F.linear(x, w_in)
    B, T, _ = x.shape
    x = x.view(B, T, 2, -1)
    x = act(x[..., 0, :]) * x[..., 1, :]
    return F.linear(x, w_out)


class StxOracle:
    """StreamingTransformer over a flat reference-keyed state dict; streaming state as in the reference's
    `with model.streaming(batch_size):` (llm_modules/streaming.py:103-111)."""

    def __init__(self, cfg: StxCfg, sd: Dict[str, torch.Tensor]):
        self.cfg = cfg
        self.sd = sd
        self.state = None

    # -- streaming API (streaming.py:94-126)
    def start_streaming(self, batch_size: int):
        c = self.cfg
        cap = c.capacity()
        hd = c.d_model // c.num_heads
        self.state = dict(offset=0, kv=[RingKV(batch_size, c.num_heads, hd, cap) for _ in range(c.num_layers)])

    def stop_streaming(self):
        self.state = None

    def reset_streaming(self):
        if self.state is None:
            raise ValueError("Trying to reset streaming, but the transformer wasn't streaming.")
        self.state["offset"] = 0
        for kv in self.state["kv"]:
            kv.reset()

    def _norm(self, x, prefix):
        c = self.cfg
        if c.norm.startswith("layer_norm"):
            return F.layer_norm(x, (c.d_model,), self.sd[prefix + ".weight"], self.sd[prefix + ".bias"], c.norm_eps())
        return rms_norm(x, self.sd[prefix + ".alpha"], c.norm_eps(), torch.float if c.norm == "rms_norm_f32" else None)

    def _attn(self, i, x, offset):
        """StreamingMultiheadAttention.forward, transformer.py:375-419."""
        c = self.cfg
        p = f"layers.{i}.self_attn."
        B, T, D = x.shape
        H = c.num_heads
        if c.weights_per_step:
            projected = multi_linear(c.weights_per_step, self.sd[p + "in_proj_weight"], x, offset)
        else:
            projected = F.linear(x, self.sd[p + "in_proj_weight"])
        q, k, v = projected.view(B, T, 3, H, D // H).permute(2, 0, 3, 1, 4)
        off_t = torch.full((1,), offset, dtype=torch.long)
        if c.positional_embedding in ("rope", "sin_rope"):
            q, k = apply_rope(q, k, off_t, c.max_period)
        if self.state is None:
            pos_k = torch.arange(T, dtype=torch.long)
        else:
            k, v, pos_k = self.state["kv"][i].complete(k, v)
        if c.causal:
            pos_k = pos_k.view(1, -1)
            pos_q = off_t + torch.arange(T, dtype=torch.long).view(-1, 1)
            delta = pos_q - pos_k
            attn_bias = (pos_k >= 0) & (delta >= 0)
            if c.context is not None:
                attn_bias = attn_bias & (delta < c.context)
        else:
            attn_bias = None
        y = F.scaled_dot_product_attention(q, k, v, attn_bias, dropout_p=0.0)
        y = y.permute(0, 2, 1, 3).reshape(B, T, D)
        if c.weights_per_step:
            return multi_linear(c.weights_per_step, self.sd[p + "out_proj.weight"], y, offset)
        return F.linear(y, self.sd[p + "out_proj.weight"])

    def _ff(self, i, x, offset):
        """_ff_block without the residual, transformer.py:545-569."""
        c = self.cfg
        p = f"layers.{i}."
        if c.gating == "none":
            return F.linear(F.gelu(F.linear(x, self.sd[p + "linear1.weight"])), self.sd[p + "linear2.weight"])
        act = _ACT[c.gating]
        if c.weights_per_step:
            ys = []
            for t in range(x.shape[1]):
                g = p + f"gating.{offset + t}."
                # .contiguous(): with nn.Parameter weights (the reference) ATen folds the strided (B, 1, D) slice like a
                # contiguous one; with plain tensors it takes another matmul path that differs in the last bit
                ys.append(gating_forward(self.sd[g + "linear_in.weight"], self.sd[g + "linear_out.weight"], act,
                                         x[:, t:t + 1].contiguous()))
            return torch.cat(ys, dim=1)
        return gating_forward(self.sd[p + "gating.linear_in.weight"], self.sd[p + "gating.linear_out.weight"], act, x)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """StreamingTransformer.forward, transformer.py:671-692: x (B, T, C) -> (B, T, C)."""
        c = self.cfg
        B, T, C = x.shape
        offset = 0 if self.state is None else self.state["offset"]
        if self.state is not None and not c.causal:
            raise AssertionError("Streaming only available for causal")
        if c.positional_embedding in ("sin", "sin_rope"):
            positions = torch.arange(T).view(1, -1, 1) + offset
            x = x + c.positional_scale * create_sin_embedding(positions, C, max_period=c.max_period, dtype=x.dtype)
        for i in range(c.num_layers):
            p = f"layers.{i}."
            upd = self._attn(i, self._norm(x, p + "norm1"), offset)
            x = x + (self.sd[p + "layer_scale_1.scale"] * upd if c.layer_scale is not None else upd)
            upd = self._ff(i, self._norm(x, p + "norm2"), offset)
            x = x + (self.sd[p + "layer_scale_2.scale"] * upd if c.layer_scale is not None else upd)
        if self.state is not None:
            self.state["offset"] += T
        return x


# --------------------------------------------------------------------------------------------------------------
# sampler (llm_utils/sampling.py).  `q` = the Exp(1) draws the reference takes inside multinomial()
# (torch.empty_like(input_).exponential_(1), :41); None -> drawn here from the global generator like the reference.
# --------------------------------------------------------------------------------------------------------------
def multinomial(input: torch.Tensor, q: Optional[torch.Tensor] = None) -> torch.Tensor:
    """sampling.py:15-46, num_samples=1 without replacement: argmax(p / Exp(1))."""
    input_ = input.reshape(-1, input.shape[-1])
    if q is None:
        q = torch.empty_like(input_).exponential_(1)
    q = input_ / q.reshape(input_.shape)
    out = q.argmax(dim=-1, keepdim=True)
    return out.reshape(*list(input.shape[:-1]), -1)


def sample_top_k(probs, k, q=None):
    """sampling.py:49-61; the noise has one entry per RANK of the sorted top-k, not per vocabulary id."""
    probs, indices = torch.topk(probs, k, dim=-1)
    nt = multinomial(probs, q)
    return indices.gather(-1, nt)


def sample_top_p(probs, p, q=None):
    """sampling.py:64-81."""
    probs_sort, probs_idx = torch.sort(probs, dim=-1, descending=True)
    probs_sum = torch.cumsum(probs_sort, dim=-1)
    mask = probs_sum - probs_sort > p
    probs_sort = probs_sort * (~mask).float()
    probs_sort = probs_sort / probs_sort.sum(dim=-1, keepdim=True)
    nt = multinomial(probs_sort, q)
    return torch.gather(probs_idx, -1, nt)


def sample_token(logits, use_sampling=False, temp=1.0, top_k=0, top_p=0.0, q=None, end_token: Optional[int] = None):
    """sample_token (sampling.py:84-105); end_token != None gives sample_token_audio (:107-130: probabilities of ids
    >= end_token are overwritten with -inf AFTER the softmax; sample_token_audio_2048 is end_token = 2048)."""
    if use_sampling and temp > 0.0:
        probs = torch.softmax(logits / temp, dim=-1)
        if end_token is not None:
            probs = probs.clone()
            probs[..., end_token:] = float("-inf")
        if top_p > 0.0:
            nt = sample_top_p(probs, top_p, q)
        elif top_k > 0:
            nt = sample_top_k(probs, top_k, q)
        else:
            nt = multinomial(probs, q)
    else:
        nt = torch.argmax(logits, dim=-1, keepdim=True)
    return nt[..., 0]
