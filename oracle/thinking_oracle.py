"""TEST INFRASTRUCTURE ONLY (imported by tests/ and oracle/make_golden_thinking.py; never by uniaudio2_b200/).

CPU restatement of the reasoning encoder of ReasoningCodec_film's tokenize direction (`AudioThinking`), up to the
query tokens that go into `reasoning_vq`:

    tools/tokenizer/ReasoningCodec_film/models/AudioDiffusion1D.py
        :169-188   AudioThinking.__init__: cls_token, 5 x TransformerBlock(dim, dim_heads=128, power_normalized, layer_scale, add_rope,
                   attn qk_norm, ff mult 4 with biases), semantic_merge_proj Linear(whisper_dim + 1024, dim), down_sampling_layer_whisper
                   Conv1d(k 2, s 2), reasoning_vq
        :372-390   encode_reasoning_part: down-sample, concatenate with the BEST-RQ features, merge, set_masking, encoder, extract_mask_positions
        :458-486   set_masking / extract_mask_positions: one query token after every `interval` frames
    tools/tokenizer/ReasoningCodec_film/modules/transformer.py
        :645-783   TransformerBlock (power_normalized forces remove_norms: no pre-norms), :293-598 Attention (fused weight-normed to_qkv,
                   per-head LayerNorm of q and k, partial rotary embedding on the first 64 of 128 head dims, softmax attention, weight-normed
                   to_out), :206-291 GLU / FeedForward (sigmoid gate, weight-normed in / out), :197-202 LayerScale, :89-172 rotary helpers

Pinned bit-exactly: oracle/make_golden_thinking.py imports the UNMODIFIED modules/transformer.py (its one missing import,
soft_moe_pytorch, is unused by this configuration and stubbed) and executes the unmodified method source of encode_reasoning_part /
set_masking / extract_mask_positions on a stand-in self; tests/test_thinking_oracle.py checks this file against those fixtures.
On the CPU the reference's Attention takes its einsum path (flash-attn needs CUDA); on a GPU it would call flash_attn in fp16.
"""
import math

import torch
import torch.nn.functional as F

CFG = dict(dim=768, dim_heads=128, depth=5, interval=5, whisper_dim=1024, mu_dim=1024, ff_mult=4)


def wn(sd, pre):
    """Effective weight of nn.utils.parametrizations.weight_norm(module) (dim = 0): g * v / ||v|| per output row."""
    if pre + ".weight" in sd:
        return sd[pre + ".weight"]
    return torch._weight_norm(sd[pre + ".parametrizations.weight.original1"], sd[pre + ".parametrizations.weight.original0"], 0)


def rotary(t, inv_freq):
    """apply_rotary_pos_emb with freqs = cat(pos * inv_freq, pos * inv_freq): t (B, H, N, hd), first 2 * len(inv_freq) dims rotated."""
    n = t.shape[-2]
    freqs = torch.einsum("i , j -> i j", torch.arange(n).to(torch.float32), inv_freq)
    freqs = torch.cat((freqs, freqs), dim=-1)
    rot = freqs.shape[-1]
    a, rest = t[..., :rot], t[..., rot:]
    x1, x2 = a[..., :rot // 2], a[..., rot // 2:]
    a = (a * freqs.cos()) + (torch.cat((-x2, x1), dim=-1) * freqs.sin())
    return torch.cat((a, rest), dim=-1)


def block(sd, cfg, pre, x):
    B, N, D = x.shape
    hd = cfg["dim_heads"]
    H = D // hd
    q, k, v = F.linear(x, wn(sd, pre + "self_attn.to_qkv")).chunk(3, dim=-1)
    q, k, v = (t.view(B, N, H, hd).transpose(1, 2) for t in (q, k, v))
    q = F.layer_norm(q, (hd,), sd[pre + "self_attn.q_norm.weight"], sd[pre + "self_attn.q_norm.bias"], 1e-5)
    k = F.layer_norm(k, (hd,), sd[pre + "self_attn.k_norm.weight"], sd[pre + "self_attn.k_norm.bias"], 1e-5)
    q, k = rotary(q, sd[pre + "rope.inv_freq"]), rotary(k, sd[pre + "rope.inv_freq"])
    dots = torch.einsum("b h i d, b h j d -> b h i j", q, k) * (1.0 / (hd ** 0.5))
    att = torch.einsum("b h i j, b h j d -> b h i d", F.softmax(dots, dim=-1, dtype=torch.float32), v)
    att = att.transpose(1, 2).reshape(B, N, D)
    x = x + F.linear(att, wn(sd, pre + "self_attn.to_out")) * sd[pre + "self_attn_scale.scale"]
    a, gate = F.linear(x, wn(sd, pre + "ff.ff.0.proj"), sd[pre + "ff.ff.0.proj.bias"]).chunk(2, dim=-1)
    y = F.linear(a * torch.sigmoid(gate), wn(sd, pre + "ff.ff.2"), sd[pre + "ff.ff.2.bias"])
    return x + y * sd[pre + "ff_scale.scale"]


def set_masking(x, cls_token, interval):
    B, T, D = x.shape
    xr = x.reshape(B, T // interval, interval, D)
    tok = cls_token.repeat(1, T // interval, 1).repeat(B, 1, 1).unsqueeze(2)
    return torch.cat([xr, tok], dim=2).reshape(B, -1, D)


def encode(sd, cfg, whisper_embeds, mu_embeds, return_hidden=False):
    """whisper (B, whisper_dim, Tw), BEST-RQ (B, mu_dim, Tb) -> query tokens (B, T / interval, dim), T = min(Tw / 2, Tb)."""
    w = F.conv1d(whisper_embeds, sd["down_sampling_layer_whisper.weight"], sd["down_sampling_layer_whisper.bias"], stride=2).transpose(1, 2)
    mu = mu_embeds.transpose(1, 2)
    n = min(w.shape[1], mu.shape[1])
    x = F.linear(torch.cat((w[:, :n], mu[:, :n]), dim=-1), sd["semantic_merge_proj.weight"], sd["semantic_merge_proj.bias"])
    x = set_masking(x, sd["cls_token"], cfg["interval"])
    for i in range(cfg["depth"]):
        x = block(sd, cfg, f"encoder_transformers.{i}.", x)
    q = x[:, cfg["interval"]::cfg["interval"] + 1]
    return (q, x) if return_hidden else q


def random_state_dict(cfg, seed):
    """Seeded parameters under the reference's state-dict names (weight-normed linears as original0 = g (out, 1), original1 = v)."""
    g = torch.Generator().manual_seed(seed)
    D, hd, Fi = cfg["dim"], cfg["dim_heads"], cfg["dim"] * cfg["ff_mult"]

    def rn(*shape, scale=1.0):
        return torch.randn(*shape, generator=g) * scale

    def wnorm(pre, n_out, n_in, gain):
        v = rn(n_out, n_in, scale=1 / math.sqrt(n_in))
        sd[pre + ".parametrizations.weight.original0"] = v.norm(dim=1, keepdim=True) * (gain + rn(n_out, 1, scale=0.1 * gain))
        sd[pre + ".parametrizations.weight.original1"] = v

    sd = {"cls_token": rn(1, D),
          "down_sampling_layer_whisper.weight": rn(cfg["whisper_dim"], cfg["whisper_dim"], 2, scale=1 / math.sqrt(2 * cfg["whisper_dim"])),
          "down_sampling_layer_whisper.bias": rn(cfg["whisper_dim"], scale=0.1),
          "semantic_merge_proj.weight": rn(D, cfg["whisper_dim"] + cfg["mu_dim"], scale=1 / math.sqrt(cfg["whisper_dim"] + cfg["mu_dim"])),
          "semantic_merge_proj.bias": rn(D, scale=0.1)}
    rot = max(hd // 2, 32)
    for i in range(cfg["depth"]):
        pre = f"encoder_transformers.{i}."
        wnorm(pre + "self_attn.to_qkv", 3 * D, D, 1.0)
        wnorm(pre + "self_attn.to_out", D, D, 1.0)
        for n in ("q_norm", "k_norm"):
            sd[pre + f"self_attn.{n}.weight"] = 1 + rn(hd, scale=0.1)
            sd[pre + f"self_attn.{n}.bias"] = rn(hd, scale=0.1)
        sd[pre + "self_attn_scale.scale"] = 0.5 + rn(D, scale=0.1)   # (the reference initialises LayerScale at 1e-2: too small to test a branch)
        wnorm(pre + "ff.ff.0.proj", 2 * Fi, D, 1.0)
        sd[pre + "ff.ff.0.proj.bias"] = rn(2 * Fi, scale=0.1)
        wnorm(pre + "ff.ff.2", D, Fi, 1.0)
        sd[pre + "ff.ff.2.bias"] = rn(D, scale=0.1)
        sd[pre + "ff_scale.scale"] = 0.5 + rn(D, scale=0.1)
        sd[pre + "rope.inv_freq"] = 1.0 / (10000 ** (torch.arange(0, rot, 2).float() / rot))
    return sd
