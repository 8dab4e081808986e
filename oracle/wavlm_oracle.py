"""TEST INFRASTRUCTURE ONLY (imported by tests/ and oracle/make_golden_frontend.py; never by uniaudio2_b200/).

CPU restatement of the WavLM encoder as ReasoningCodec_film's tokenize direction calls it:

    tools/tokenizer/ReasoningCodec_film/models/AudioDiffusion1D.py
        :226      self.wavlm_encoder = AutoModel.from_pretrained(wav_lm_path)          -> transformers WavLMModel (768-wide: base / base-plus)
        :359-370  get_wavlm_feature(wav_24k, len_semantic): Resample(24000, 16000), 160 zeros appended,
                  wavlm_encoder(wav_16k, output_hidden_states=True).hidden_states, stack, [:, 6:10].mean(1), transpose, cut to 2 * len_semantic

The model code is third-party: transformers==4.57.0 (pyproject.toml:25), absent from /root/reference.  This image carries
transformers 5.5.0, so the restatement is pinned against the installed class: tests/test_wavlm_oracle.py builds
transformers.WavLMModel(WavLMConfig(...)) with random weights and requires every hidden state of this file to match it, and
oracle/make_golden_frontend.py stores such outputs as fixtures for the GPU tests.  Served: feat_extract_norm="group",
do_stable_layer_norm=False (wavlm-base, wavlm-base-plus), eval mode, no attention mask.

Functions follow transformers/models/wavlm/modeling_wavlm.py: WavLMFeatureEncoder (WavLMGroupNormConvLayer +
WavLMNoLayerNormConvLayer), WavLMFeatureProjection, WavLMPositionalConvEmbedding + WavLMSamePadLayer, WavLMEncoder,
WavLMEncoderLayer, WavLMAttention (compute_bias, _relative_positions_bucket, torch_multi_head_self_attention), WavLMFeedForward.
"""
import math

import torch
import torch.nn.functional as F

from oracle import frontend_oracle as FO

BASE_PLUS = dict(hidden_size=768, num_attention_heads=12, intermediate_size=3072, num_hidden_layers=12,
                 conv_dim=(512,) * 7, conv_kernel=(10, 3, 3, 3, 3, 2, 2), conv_stride=(5, 2, 2, 2, 2, 2, 2), conv_bias=False,
                 num_conv_pos_embeddings=128, num_conv_pos_embedding_groups=16, num_buckets=320, max_bucket_distance=800,
                 layer_norm_eps=1e-5)


def relative_positions_bucket(rel, num_buckets, max_distance):
    """WavLMAttention._relative_positions_bucket; rel = memory_position - context_position (long tensor)."""
    nb = num_buckets // 2
    buckets = (rel > 0).to(torch.long) * nb
    rel = torch.abs(rel)
    max_exact = nb // 2
    is_small = rel < max_exact
    large = torch.log(rel.float() / max_exact)
    large = large / math.log(max_distance / max_exact)
    large = large * (nb - max_exact)
    large = (max_exact + large).to(torch.long)
    large = torch.min(large, torch.full_like(large, nb - 1))
    return buckets + torch.where(is_small, rel, large)


def position_bias(sd, cfg, T):
    """(H, T, T) = rel_attn_embed(bucket(j - i)) of layer 0 (WavLMAttention.compute_bias)."""
    ctx = torch.arange(T, dtype=torch.long)[:, None]
    mem = torch.arange(T, dtype=torch.long)[None, :]
    bucket = relative_positions_bucket(mem - ctx, cfg["num_buckets"], cfg["max_bucket_distance"])
    return sd["encoder.layers.0.attention.rel_attn_embed.weight"][bucket].permute(2, 0, 1)


def pos_conv_weight(sd):
    """The weight-normalised kernel of encoder.pos_conv_embed.conv (nn.utils.parametrizations.weight_norm, dim=2)."""
    pre = "encoder.pos_conv_embed.conv."
    if pre + "weight" in sd:
        return sd[pre + "weight"]
    g, v = sd[pre + "parametrizations.weight.original0"], sd[pre + "parametrizations.weight.original1"]
    return torch._weight_norm(v, g, 2)


def feature_encoder(sd, cfg, wav):
    """wav (B, L) -> (B, C, T)"""
    x = wav[:, None]
    for i, (k, s) in enumerate(zip(cfg["conv_kernel"], cfg["conv_stride"])):
        pre = f"feature_extractor.conv_layers.{i}."
        x = F.conv1d(x, sd[pre + "conv.weight"], sd.get(pre + "conv.bias") if cfg["conv_bias"] else None, stride=s)
        if i == 0:
            C = x.shape[1]
            x = F.group_norm(x, C, sd[pre + "layer_norm.weight"], sd[pre + "layer_norm.bias"], 1e-5)
        x = F.gelu(x)
    return x


def gate(sd, cfg, pre, h):
    """(B, H, T) gate of the position bias from the layer input (WavLMAttention.forward steps 1-3)."""
    B, T, D = h.shape
    H = cfg["num_attention_heads"]
    g = h.view(B, T, H, D // H).permute(0, 2, 1, 3)
    proj = F.linear(g, sd[pre + "gru_rel_pos_linear.weight"], sd[pre + "gru_rel_pos_linear.bias"])
    proj = proj.view(B, H, T, 2, 4).sum(-1)
    ga, gb = torch.sigmoid(proj).chunk(2, dim=-1)
    return (ga * (gb * sd[pre + "gru_rel_pos_const"] - 1.0) + 2.0)[..., 0]


def attention(sd, cfg, pre, h, pos_bias):
    B, T, D = h.shape
    H = cfg["num_attention_heads"]
    hd = D // H
    bias = gate(sd, cfg, pre, h)[..., None] * pos_bias[None]              # (B, H, T, T)
    q = F.linear(h, sd[pre + "q_proj.weight"], sd[pre + "q_proj.bias"]).view(B, T, H, hd).transpose(1, 2)
    k = F.linear(h, sd[pre + "k_proj.weight"], sd[pre + "k_proj.bias"]).view(B, T, H, hd).transpose(1, 2)
    v = F.linear(h, sd[pre + "v_proj.weight"], sd[pre + "v_proj.bias"]).view(B, T, H, hd).transpose(1, 2)
    att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hd) + bias, dim=-1) @ v
    att = att.transpose(1, 2).reshape(B, T, D)
    return F.linear(att, sd[pre + "out_proj.weight"], sd[pre + "out_proj.bias"])


def hidden_states(sd, cfg, wav, n_layers=None):
    """wav (B, L) fp32 at 16 kHz -> list of hidden states [(B, T, D)] * (n_layers + 1), like output_hidden_states=True."""
    eps = cfg["layer_norm_eps"]
    D = cfg["hidden_size"]
    x = feature_encoder(sd, cfg, wav).transpose(1, 2)
    x = F.layer_norm(x, (x.shape[-1],), sd["feature_projection.layer_norm.weight"], sd["feature_projection.layer_norm.bias"], eps)
    h = F.linear(x, sd["feature_projection.projection.weight"], sd["feature_projection.projection.bias"])
    K = cfg["num_conv_pos_embeddings"]
    pc = F.conv1d(h.transpose(1, 2), pos_conv_weight(sd), sd["encoder.pos_conv_embed.conv.bias"], padding=K // 2,
                  groups=cfg["num_conv_pos_embedding_groups"])
    if K % 2 == 0:
        pc = pc[:, :, :-1]
    h = h + F.gelu(pc).transpose(1, 2)
    h = F.layer_norm(h, (D,), sd["encoder.layer_norm.weight"], sd["encoder.layer_norm.bias"], eps)
    out = [h]
    n_layers = cfg["num_hidden_layers"] if n_layers is None else n_layers
    pb = position_bias(sd, cfg, h.shape[1]) if n_layers else None
    for i in range(n_layers):
        pre = f"encoder.layers.{i}."
        h = h + attention(sd, cfg, pre + "attention.", h, pb)
        h = F.layer_norm(h, (D,), sd[pre + "layer_norm.weight"], sd[pre + "layer_norm.bias"], eps)
        ff = F.gelu(F.linear(h, sd[pre + "feed_forward.intermediate_dense.weight"], sd[pre + "feed_forward.intermediate_dense.bias"]))
        h = h + F.linear(ff, sd[pre + "feed_forward.output_dense.weight"], sd[pre + "feed_forward.output_dense.bias"])
        h = F.layer_norm(h, (D,), sd[pre + "final_layer_norm.weight"], sd[pre + "final_layer_norm.bias"], eps)
        out.append(h)
    return out


def get_wavlm_feature(sd, cfg, wav_24k, len_semantic, lo=6, hi=10):
    """AudioDiffusion1D.get_wavlm_feature: wav_24k (B, 1, T) -> (B, D, min(T', 2 * len_semantic))."""
    wav16 = FO.resample(wav_24k, 24000, 16000).squeeze(1)
    wav16 = torch.cat([wav16, torch.zeros(wav16.shape[0], 160)], dim=-1)
    hs = hidden_states(sd, cfg, wav16, n_layers=hi - 1)
    target = torch.stack(hs, dim=1)[:, lo:hi].mean(1).transpose(1, 2)
    return target[:, :, :min(target.shape[-1], len_semantic * 2)]


def random_state_dict(cfg, seed):
    """A seeded WavLMModel state dict (transformers' key names and shapes) with every term active: non-zero biases, gate constants
    off 1, a full-scale bucket embedding.  Fixtures store only the seed (tests/golden/frontend_golden.pt)."""
    g = torch.Generator().manual_seed(seed)
    D, Fi, H = cfg["hidden_size"], cfg["intermediate_size"], cfg["num_attention_heads"]

    def rn(*shape, scale=1.0):
        return torch.randn(*shape, generator=g) * scale

    def lin(pre, n_out, n_in):
        sd[pre + ".weight"] = rn(n_out, n_in, scale=1 / math.sqrt(n_in))
        sd[pre + ".bias"] = rn(n_out, scale=0.1)

    def norm(pre, n):
        sd[pre + ".weight"] = 1 + rn(n, scale=0.1)
        sd[pre + ".bias"] = rn(n, scale=0.1)

    sd = {"masked_spec_embed": torch.rand(D, generator=g)}
    for i, (co, k) in enumerate(zip(cfg["conv_dim"], cfg["conv_kernel"])):
        ci = 1 if i == 0 else cfg["conv_dim"][i - 1]
        pre = f"feature_extractor.conv_layers.{i}."
        sd[pre + "conv.weight"] = rn(co, ci, k, scale=math.sqrt(2 / (ci * k)))
        if cfg["conv_bias"]:
            sd[pre + "conv.bias"] = rn(co, scale=0.1)
        if i == 0:
            norm(pre + "layer_norm", co)
    norm("feature_projection.layer_norm", cfg["conv_dim"][-1])
    lin("feature_projection.projection", D, cfg["conv_dim"][-1])
    cg, K = D // cfg["num_conv_pos_embedding_groups"], cfg["num_conv_pos_embeddings"]
    v = rn(D, cg, K, scale=1 / math.sqrt(cg * K))
    sd["encoder.pos_conv_embed.conv.bias"] = rn(D, scale=0.1)
    sd["encoder.pos_conv_embed.conv.parametrizations.weight.original0"] = v.norm(dim=(0, 1), keepdim=True) * (1 + rn(1, 1, K, scale=0.2))
    sd["encoder.pos_conv_embed.conv.parametrizations.weight.original1"] = v
    norm("encoder.layer_norm", D)
    for i in range(cfg["num_hidden_layers"]):
        pre = f"encoder.layers.{i}."
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            lin(pre + "attention." + n, D, D)
        sd[pre + "attention.gru_rel_pos_const"] = 1 + rn(1, H, 1, 1, scale=0.3)
        lin(pre + "attention.gru_rel_pos_linear", 8, D // H)
        if i == 0:
            sd[pre + "attention.rel_attn_embed.weight"] = rn(cfg["num_buckets"], H)
        norm(pre + "layer_norm", D)
        lin(pre + "feed_forward.intermediate_dense", Fi, D)
        lin(pre + "feed_forward.output_dense", D, Fi)
        norm(pre + "final_layer_norm", D)
    return sd
