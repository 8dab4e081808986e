"""CPU oracle for the UniAudio2 AR-decode hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.  The product path (uniaudio2_b200) never imports it and has no CPU fallback.

This is a restatement (not a copy) of the reference's algorithm in plain torch-CPU fp32 ops, written
as pure functions over a flat state dict that uses the reference's own parameter names.  Every
function cites the reference file:line it follows (paths relative to /root/reference).

Parity status: PINNED.  oracle/make_golden.py imports the unmodified reference (via oracle/ref_shims)
in the build container and asserts this restatement is bit-identical to it on CPU for prefill and
N frames of generate_frame (greedy and top-k with shared RNG stream); the resulting vectors are
committed under tests/golden/ and re-checked by tests/test_oracle_golden.py (no GPU needed).
The reference itself has no golden vectors or tests for this path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# configs (llm_models/config.py:785-899: the five Llama-3.2 entries on the path)
# --------------------------------------------------------------------------------------
@dataclass
class GPTCfg:
    n_layer: int
    n_embd: int
    n_head: int
    n_query_groups: int
    intermediate_size: int
    head_size: Optional[int] = None
    padded_vocab_size: int = 128256
    norm_eps: float = 1e-5  # config.py:38 default (the Llama-3.2 entries do not override it)
    rope_base: int = 500000
    rope_adjustments: Optional[dict] = field(
        default_factory=lambda: dict(factor=32.0, low_freq_factor=1.0, high_freq_factor=4.0, original_max_seq_len=8192)
    )
    block_size: int = 131072

    def __post_init__(self):
        if self.head_size is None:
            self.head_size = self.n_embd // self.n_head  # config.py:105-107


LLAMA_CFGS = {
    "Llama-3.2-3B": dict(n_layer=28, n_embd=3072, n_head=24, n_query_groups=8, intermediate_size=8192),
    "Llama-3.2-1B": dict(n_layer=16, n_embd=2048, n_head=32, n_query_groups=8, intermediate_size=8192),
    "Llama-3.2-300M": dict(n_layer=4, n_embd=2048, n_head=32, n_query_groups=8, intermediate_size=8192),
    "Llama-3.2-4Layer": dict(n_layer=4, n_embd=2048, n_head=32, n_query_groups=8, intermediate_size=8192),
    "Llama-3.2-Understanding": dict(n_layer=3, n_embd=3072, n_head=24, n_query_groups=8, intermediate_size=8192),
    "Llama-3.2-Generation": dict(n_layer=2, n_embd=3072, n_head=24, n_query_groups=8, intermediate_size=8192),
}


@dataclass
class Stage3Cfg:
    """Everything Model_stage3.__init__ derives from ModelArgs (model_new.py:340-355)."""

    backbone: GPTCfg
    decoder: GPTCfg
    understanding: GPTCfg
    generation: GPTCfg
    audio_vocab: int  # audio_semantic_vocab_size + audio_reason_vocab_size
    num_codebooks: int = 8
    max_seq_length: int = 2048  # model_new.py:560-565


def full_size_cfg(audio_reason_card: int = 4100, audio_semantic_card: int = 8200) -> Stage3Cfg:
    return Stage3Cfg(
        backbone=GPTCfg(**LLAMA_CFGS["Llama-3.2-3B"]),
        decoder=GPTCfg(**LLAMA_CFGS["Llama-3.2-300M"]),
        understanding=GPTCfg(**LLAMA_CFGS["Llama-3.2-Understanding"]),
        generation=GPTCfg(**LLAMA_CFGS["Llama-3.2-Generation"]),
        audio_vocab=audio_reason_card + audio_semantic_card,
    )


# --------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------
def build_rope_cache(seq_len: int, n_elem: int, base: int, extra: Optional[dict]):
    """llm_models/lit_model.py:634-706 (Llama-3 smooth scaling branch :662-676)."""
    theta = 1.0 / (base ** (torch.arange(0, n_elem, 2).float() / n_elem))
    if extra is not None:
        factor = extra["factor"]
        if "original_max_seq_len" in extra:
            wavelen = 2 * torch.pi / theta
            ratio = extra["original_max_seq_len"] / wavelen
            smooth = (ratio - extra["low_freq_factor"]) / (extra["high_freq_factor"] - extra["low_freq_factor"])
            smooth = torch.clamp(smooth, min=0.0, max=1.0)
            theta = (1 - smooth) * (theta / factor) + smooth * theta
        else:
            theta = theta / factor
    seq_idx = torch.arange(seq_len) / 1
    idx_theta = torch.outer(seq_idx, theta).repeat(1, 2)
    if idx_theta.shape[-1] > n_elem > 1:
        idx_theta = idx_theta[..., :n_elem]
    return torch.cos(idx_theta), torch.sin(idx_theta)


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """lit_model.py:883-890."""
    dtype = x.dtype
    x = x.float()
    norm_x = torch.mean(x * x, dim=-1, keepdim=True)
    x_normed = x * torch.rsqrt(norm_x + eps)
    return (x_normed * w.float()).to(dtype=dtype)


def apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """lit_model.py:778-807 (half-split rotation).  x (B,nh,T,hs); cos/sin (B|1,T,hs)."""
    half = x.size(-1) // 2
    x1 = x[..., :half]
    x2 = x[..., half:]
    rotated = torch.cat((-x2, x1), dim=-1)
    cos = cos.unsqueeze(1)
    sin = sin.unsqueeze(1)
    return ((x * cos) + (rotated * sin)).to(dtype=x.dtype)


class KV:
    """lit_model.py:814-860 - preallocated (B, n_query_groups, max_seq, hs) fp32 buffers."""

    def __init__(self, B, G, S, hs):
        self.k = torch.zeros(B, G, S, hs)
        self.v = torch.zeros(B, G, S, hs)

    def write(self, input_pos: torch.Tensor, k: torch.Tensor, v: torch.Tensor):
        bs = k.size(0)
        if input_pos.dim() == 1:  # lit_model.py:760-761
            self.k[:bs].index_copy_(-2, input_pos, k)
            self.v[:bs].index_copy_(-2, input_pos, v)
        else:  # lit_model.py:772-774 per-batch loop
            for i in range(bs):
                self.k[i].index_copy_(-2, input_pos[i], k[i])
                self.v[i].index_copy_(-2, input_pos[i], v[i])
        return self.k[:bs], self.v[:bs]


class GPTOracle:
    """llm_models/lit_model.py::GPT with wte/lm_head handled by the caller (model_new.py:111-120)."""

    def __init__(self, cfg: GPTCfg, sd: Dict[str, torch.Tensor], prefix: str):
        self.cfg, self.sd, self.prefix = cfg, sd, prefix
        self.kv: Optional[List[KV]] = None
        self.mask_cache = None
        self.max_seq = None
        self.cos = self.sin = None

    def w(self, name):
        return self.sd[self.prefix + name]

    def set_kv_cache(self, B: int, max_seq: int):
        """lit_model.py:224-254 + build_mask_cache :863-866.  The rope cache is only ever indexed at
        input_pos < max_seq, so building it for max_seq rows equals the reference's block_size rows."""
        c = self.cfg
        self.kv = [KV(B, c.n_query_groups, max_seq, c.head_size) for _ in range(c.n_layer)]
        self.mask_cache = torch.tril(torch.ones(max_seq, max_seq, dtype=torch.bool)).unsqueeze(0).unsqueeze(0)
        self.max_seq = max_seq
        self.cos, self.sin = build_rope_cache(max_seq, c.head_size, c.rope_base, c.rope_adjustments)

    def reset_kv_cache(self):
        """lit_model.py:256-269 (zero-fill)."""
        for kv in self.kv:
            kv.k.zero_()
            kv.v.zero_()

    def forward(self, x: torch.Tensor, input_pos: Optional[torch.Tensor], input_pos_maxp1: Optional[int] = None):
        """lit_model.py:83-180 (kv-cache branch :123-145; no-cache branch :146-152)."""
        c = self.cfg
        B, T, _ = x.shape
        if input_pos is not None:
            cos = self.cos[input_pos]
            sin = self.sin[input_pos]
            if input_pos.dim() == 1:
                cos, sin = cos.unsqueeze(0), sin.unsqueeze(0)
                mask = self.mask_cache[:, :, input_pos, :]  # (1,1,T,S)
            else:
                mask = self.mask_cache[0, :, input_pos, :].permute(1, 0, 2, 3)  # (B,1,T,S)
            if input_pos_maxp1 is not None:
                mask = mask[..., :input_pos_maxp1]
        else:
            cos = build_rope_cache(T, c.head_size, c.rope_base, c.rope_adjustments)[0].unsqueeze(0) if self.cos is None else self.cos[:T].unsqueeze(0)
            sin = build_rope_cache(T, c.head_size, c.rope_base, c.rope_adjustments)[1].unsqueeze(0) if self.sin is None else self.sin[:T].unsqueeze(0)
            mask = None
            input_pos_maxp1 = None
        for l in range(c.n_layer):
            x = self.block(l, x, cos, sin, mask, input_pos, input_pos_maxp1)
        return rms_norm(x, self.w("transformer.ln_f.weight"), c.norm_eps)

    def block(self, l, x, cos, sin, mask, input_pos, maxp1):
        """lit_model.py:307-349 (non-parallel residual), attention :382-511, LLaMAMLP :591-595."""
        c = self.cfg
        p = f"transformer.h.{l}."
        B, T, _ = x.shape
        xn = rms_norm(x, self.w(p + "norm_1.weight"), c.norm_eps)
        qkv = F.linear(xn, self.w(p + "attn.qkv.weight"))
        qs, ks = c.n_head * c.head_size, c.n_query_groups * c.head_size
        q, k, v = qkv.split((qs, ks, ks), dim=-1)
        q = q.view(B, T, c.n_head, c.head_size).transpose(1, 2)
        k = k.view(B, T, c.n_query_groups, c.head_size).transpose(1, 2)
        v = v.view(B, T, c.n_query_groups, c.head_size).transpose(1, 2)
        q = apply_rope(q, cos, sin)
        k = apply_rope(k, cos, sin)
        if input_pos is not None:
            k, v = self.kv[l].write(input_pos, k, v)
            if maxp1 is not None:
                k = k[..., :maxp1, :]
                v = v[..., :maxp1, :]
        if c.n_query_groups != c.n_head:
            rep = c.n_head // c.n_query_groups
            k = k.repeat_interleave(rep, dim=1)
            v = v.repeat_interleave(rep, dim=1)
        scale = 1.0 / math.sqrt(c.head_size)
        y = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=0.0, scale=scale, is_causal=mask is None)
        y = y.transpose(1, 2).reshape(B, T, c.head_size * c.n_head)
        x = F.linear(y, self.w(p + "attn.proj.weight")) + x
        xn = rms_norm(x, self.w(p + "norm_2.weight"), c.norm_eps)
        h = F.silu(F.linear(xn, self.w(p + "mlp.fc_1.weight"))) * F.linear(xn, self.w(p + "mlp.fc_2.weight"))
        return F.linear(h, self.w(p + "mlp.proj.weight")) + x


def sample_topk(logits, topk, temperature, noise=None):
    """model_new.py:146-156 + :141-143.  `noise` (same shape as logits, Exp(1) draws) replaces the
    in-place exponential_ so the draw can be shared with the device under test."""
    logits = logits / temperature
    indices_to_remove = logits < torch.topk(logits, topk)[0][..., -1, None]
    scores = logits.masked_fill(indices_to_remove, -float("Inf"))
    scores = F.log_softmax(scores, dim=-1)
    probs = F.softmax(scores, dim=-1)
    q = torch.empty_like(probs).exponential_(1) if noise is None else noise
    return torch.argmax(probs / q, dim=-1, keepdim=True).to(dtype=torch.int)


def audio_sample_topk(logits, topk, temperature, forbid_prefix=0, noise=None):
    """model_new.py:158-187 (error behaviour included)."""
    if temperature <= 0:
        raise ValueError("temperature must be > 0")
    if forbid_prefix < 0:
        raise ValueError("forbid_prefix must be >= 0")
    logits = logits.clone() / temperature
    vocab = logits.size(-1)
    if forbid_prefix >= vocab:
        raise ValueError("forbid_prefix must be smaller than vocab size")
    if forbid_prefix > 0:
        logits[..., :forbid_prefix] = float("-inf")
    eff = vocab - forbid_prefix
    if topk <= 0 or topk > eff:
        raise ValueError(f"topk must be in 1..{eff} given forbid_prefix={forbid_prefix}")
    indices_to_remove = logits < torch.topk(logits, topk)[0][..., -1, None]
    scores = logits.masked_fill(indices_to_remove, -float("Inf"))
    scores = F.log_softmax(scores, dim=-1)
    probs = F.softmax(scores, dim=-1)
    q = torch.empty_like(probs).exponential_(1) if noise is None else noise
    return torch.argmax(probs / q, dim=-1, keepdim=True).to(dtype=torch.int)


# --------------------------------------------------------------------------------------
# Model_stage3 restatement
# --------------------------------------------------------------------------------------
class Stage3Oracle:
    """llm_models/model_new.py::Model_stage3 (:334-687) over a flat fp32 CPU state dict with the
    reference's key names (backbone.*, decoder.*, audio_understanding_expert.*,
    audio_generation_expert.*, audio_embeddings.weight, projection.weight, audio_head)."""

    def __init__(self, cfg: Stage3Cfg, sd: Dict[str, torch.Tensor]):
        self.cfg, self.sd = cfg, sd
        self.backbone = GPTOracle(cfg.backbone, sd, "backbone.")
        self.decoder = GPTOracle(cfg.decoder, sd, "decoder.")
        self.und = GPTOracle(cfg.understanding, sd, "audio_understanding_expert.")
        self.gen = GPTOracle(cfg.generation, sd, "audio_generation_expert.")

    # model_new.py:554-565
    def setup_caches(self, max_batch_size: int):
        S = self.cfg.max_seq_length
        self.backbone.set_kv_cache(max_batch_size, S)
        self.decoder.set_kv_cache(max_batch_size, self.cfg.num_codebooks)
        self.und.set_kv_cache(max_batch_size, S)
        self.gen.set_kv_cache(max_batch_size, S)

    # model_new.py:647-651
    def reset_caches(self):
        for g in (self.backbone, self.decoder, self.gen, self.und):
            g.reset_kv_cache()

    # model_new.py:665-673
    def _embed_audio_tokens(self, tokens):
        V, nq = self.cfg.audio_vocab, self.cfg.num_codebooks
        at = tokens[:, :, :-1] + V * torch.arange(nq)
        return F.embedding(at.reshape(-1), self.sd["audio_embeddings.weight"]).reshape(tokens.size(0), tokens.size(1), nq, -1)

    # model_new.py:662-663
    def _embed_audio(self, cb, tok):
        return F.embedding(tok + cb * self.cfg.audio_vocab, self.sd["audio_embeddings.weight"])

    def _global(self, tokens, tokens_mask, input_pos, maxp1):
        """Shared body of forward_prefix (:474-497) and generate_frame (:593-613): embedding merge,
        understanding expert, backbone, generation expert.  tokens_mask already sliced to S rows."""
        dtype = torch.float32
        audio_step = tokens_mask[:, :, 0].unsqueeze(-1).to(dtype)
        text_step = tokens_mask[:, :, -1].unsqueeze(-1).to(dtype)
        emb = self._embed_audio_tokens(tokens)
        stream_mask = tokens_mask[:, :, :-1].unsqueeze(-1).to(dtype)
        audio_input = (emb * stream_mask).sum(dim=2)
        h_audio = self.und.forward(audio_input, input_pos, maxp1)
        text_emb = F.embedding(tokens[:, :, -1], self.sd["backbone.transformer.wte.weight"])
        backbone_input = h_audio * audio_step + text_emb * text_step
        h = self.backbone.forward(backbone_input, input_pos, maxp1)
        gen_in = h * audio_step
        h_audio = self.gen.forward(gen_in, input_pos, maxp1)
        return h_audio * audio_step + h * text_step

    def forward_prefix(self, tokens, tokens_mask, input_pos, compute_heads: bool = False):
        """model_new.py:456-507.  tokens (B,S-1,9), tokens_mask (B,S,9) (the reference slices [:, :-1]),
        input_pos (B,S-1).  Every caller discards the returned logits (tts_task.py:244); only the
        KV-cache side effect matters, so the head / cache-less local-decoder pass is optional here."""
        h_final = self._global(tokens, tokens_mask[:, :-1], input_pos, None)
        if not compute_heads:
            return h_final
        text_logits = F.linear(h_final, self.sd["backbone.lm_head.weight"])
        return h_final, text_logits

    def generate_frame(self, tokens, tokens_mask, input_pos, input_pos_maxp1, temperature, topk,
                       forbid_prefix=0, cfg_scale=1.0, noise: Optional[List[torch.Tensor]] = None, debug: Optional[dict] = None):
        """model_new.py:568-645.  Returns (B, 1+num_codebooks) int32."""
        B = tokens.size(0)
        nq = self.cfg.num_codebooks
        h_final = self._global(tokens, tokens_mask, input_pos, input_pos_maxp1)
        last_h = h_final[:, -1, :]
        text_logits = F.linear(last_h, self.sd["backbone.lm_head.weight"])
        use_cfg = cfg_scale > 1.0 and B > 1
        if use_cfg:
            lc = text_logits[1:, :] + (text_logits[0:1, :] - text_logits[1:, :]) * cfg_scale
            text_sample = sample_topk(lc, topk, temperature, None if noise is None else noise[0]).repeat(2, 1)
        else:
            text_sample = sample_topk(text_logits, topk, temperature, None if noise is None else noise[0])
        if debug is not None:
            debug["h_final"] = h_final.clone()
            debug["text_logits"] = text_logits.clone()
            debug["ci_logits"] = []
        curr_sample = text_sample
        curr_h = last_h.unsqueeze(1)
        curr_pos = torch.zeros(B, 1, dtype=torch.long)
        self.decoder.reset_kv_cache()  # model_new.py:629
        for i in range(nq):
            dec_in = F.linear(curr_h, self.sd["projection.weight"])
            dh = self.decoder.forward(dec_in, curr_pos)
            ci_logits = torch.mm(dh[:, -1, :], self.sd["audio_head"][i])
            if debug is not None:
                debug["ci_logits"].append(ci_logits.clone())
            nz = None if noise is None else noise[1 + i]
            if use_cfg:
                lc = ci_logits[1:, :] + (ci_logits[0:1, :] - ci_logits[1:, :]) * cfg_scale
                ci_sample = audio_sample_topk(lc, topk, temperature, forbid_prefix, nz).repeat(2, 1)
            else:
                ci_sample = audio_sample_topk(ci_logits, topk, temperature, forbid_prefix, nz)
            curr_h = self._embed_audio(i, ci_sample)
            curr_sample = torch.cat([curr_sample, ci_sample], dim=1)
            curr_pos = curr_pos[:, -1:] + 1
        return curr_sample


# --------------------------------------------------------------------------------------
# seeded weights (shared by tests, bench and the golden generator)
# --------------------------------------------------------------------------------------
def gpt_param_shapes(cfg: GPTCfg, prefix: str, with_embed: bool):
    c = cfg
    qkv_out = (c.n_head + 2 * c.n_query_groups) * c.head_size
    shapes = {}
    if with_embed:
        shapes[prefix + "lm_head.weight"] = (c.padded_vocab_size, c.n_embd)
        shapes[prefix + "transformer.wte.weight"] = (c.padded_vocab_size, c.n_embd)
    for l in range(c.n_layer):
        p = f"{prefix}transformer.h.{l}."
        shapes[p + "norm_1.weight"] = (c.n_embd,)
        shapes[p + "attn.qkv.weight"] = (qkv_out, c.n_embd)
        shapes[p + "attn.proj.weight"] = (c.n_embd, c.n_head * c.head_size)
        shapes[p + "norm_2.weight"] = (c.n_embd,)
        shapes[p + "mlp.fc_1.weight"] = (c.intermediate_size, c.n_embd)
        shapes[p + "mlp.fc_2.weight"] = (c.intermediate_size, c.n_embd)
        shapes[p + "mlp.proj.weight"] = (c.n_embd, c.intermediate_size)
    shapes[prefix + "transformer.ln_f.weight"] = (c.n_embd,)
    return shapes


def stage3_param_shapes(cfg: Stage3Cfg):
    """Key names / shapes of Model_stage3.state_dict() (model_new.py:340-355)."""
    s = {}
    s.update(gpt_param_shapes(cfg.backbone, "backbone.", True))
    s.update(gpt_param_shapes(cfg.decoder, "decoder.", False))
    s["audio_embeddings.weight"] = (cfg.audio_vocab * cfg.num_codebooks, cfg.backbone.n_embd)
    s["projection.weight"] = (cfg.decoder.n_embd, cfg.backbone.n_embd)
    s["audio_head"] = (cfg.num_codebooks, cfg.decoder.n_embd, cfg.audio_vocab)
    s.update(gpt_param_shapes(cfg.understanding, "audio_understanding_expert.", False))
    s.update(gpt_param_shapes(cfg.generation, "audio_generation_expert.", False))
    return s


def random_state_dict(cfg: Stage3Cfg, seed: int = 0, device="cpu", scale_override: Optional[float] = None):
    """Seeded synthetic weights (no checkpoints exist offline, SURVEY.md section 7 'hard parts').
    Linear ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)) like nn.Linear's default, embeddings ~ N(0,1)*0.02-ish,
    norm weights ~ 1 + 0.1*N(0,1), audio_head ~ N(0, 0.02^2)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    sd = {}
    for name, shape in stage3_param_shapes(cfg).items():
        if name.endswith("norm_1.weight") or name.endswith("norm_2.weight") or name.endswith("ln_f.weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g, device=device)
        elif name == "audio_head":
            t = 0.02 * torch.randn(shape, generator=g, device=device)
        elif name.endswith("wte.weight") or name == "audio_embeddings.weight":
            t = torch.randn(shape, generator=g, device=device)
        else:
            bound = 1.0 / math.sqrt(shape[-1]) if scale_override is None else scale_override
            t = (torch.rand(shape, generator=g, device=device) * 2 - 1) * bound
        sd[name] = t.float()
    return sd
