"""TEST INFRASTRUCTURE ONLY (CPU restatement, fp32) of the reference's Whisper ENCODER - the first SSL front-end of
ReasoningCodec_film's tokenize: tools/tokenizer/ReasoningCodec_film/models/modeling_whisper.py

    WhisperEncoder.forward          :766-867   conv1 k3 p1 + GELU, conv2 k3 s2 p1 + GELU, + embed_positions, layers, layer_norm
    WhisperEncoderLayer.forward     :394-443   pre-LN block: x + attn(LN(x)); x + fc2(gelu(fc1(LN(x))))
    WhisperAttention.forward        :255-374   q = (x Wq^T + bq) * hs^-0.5, k = x Wk^T (NO bias), v = x Wv^T + bv; softmax(q k^T) v; out_proj

as called from AudioDiffusion1D.get_whisper_feature (:334-343: `self.whisper_encoder(mels, return_dict=True).last_hidden_state`).
Pinned bit-exactly against the UNMODIFIED class source by oracle/make_golden_whisper.py (tests/golden/whisper_golden.pt).
Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may import this file."""
import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class WhisperCfg:
    d_model: int = 1024               # whisper-medium: 1024 / 16 heads / 4096 / 24 layers / 1500 positions / 80 mel bins
    encoder_attention_heads: int = 16
    encoder_ffn_dim: int = 4096
    encoder_layers: int = 24
    max_source_positions: int = 1500
    num_mel_bins: int = 80


def state_keys(cfg: WhisperCfg):
    keys = ["conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "embed_positions.weight"]
    for i in range(cfg.encoder_layers):
        p = f"layers.{i}."
        keys += [p + "self_attn.k_proj.weight", p + "self_attn.v_proj.weight", p + "self_attn.v_proj.bias", p + "self_attn.q_proj.weight",
                 p + "self_attn.q_proj.bias", p + "self_attn.out_proj.weight", p + "self_attn.out_proj.bias",
                 p + "self_attn_layer_norm.weight", p + "self_attn_layer_norm.bias", p + "fc1.weight", p + "fc1.bias", p + "fc2.weight",
                 p + "fc2.bias", p + "final_layer_norm.weight", p + "final_layer_norm.bias"]
    return keys + ["layer_norm.weight", "layer_norm.bias"]


def state_shapes(cfg: WhisperCfg):
    d, f = cfg.d_model, cfg.encoder_ffn_dim
    sh = {"conv1.weight": (d, cfg.num_mel_bins, 3), "conv1.bias": (d,), "conv2.weight": (d, d, 3), "conv2.bias": (d,),
          "embed_positions.weight": (cfg.max_source_positions, d), "layer_norm.weight": (d,), "layer_norm.bias": (d,)}
    for i in range(cfg.encoder_layers):
        p = f"layers.{i}."
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            sh[p + f"self_attn.{n}.weight"] = (d, d)
            if n != "k_proj":
                sh[p + f"self_attn.{n}.bias"] = (d,)
        for n in ("self_attn_layer_norm", "final_layer_norm"):
            sh[p + n + ".weight"] = (d,)
            sh[p + n + ".bias"] = (d,)
        sh[p + "fc1.weight"], sh[p + "fc1.bias"] = (f, d), (f,)
        sh[p + "fc2.weight"], sh[p + "fc2.bias"] = (d, f), (d,)
    return sh


def random_state_dict(cfg: WhisperCfg, seed: int):
    """Seeded stand-in weights (the checkpoint is not available offline): fan-in scaled matrices, LayerNorm gains around 1."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, s in state_shapes(cfg).items():
        if k.endswith("norm.weight"):
            sd[k] = 1.0 + 0.1 * torch.randn(s, generator=g)
        elif k.endswith(".bias"):
            sd[k] = 0.1 * torch.randn(s, generator=g)
        elif k == "embed_positions.weight":
            sd[k] = 0.5 * torch.randn(s, generator=g)
        else:
            fan_in = 1
            for n in s[1:]:
                fan_in *= n
            sd[k] = torch.randn(s, generator=g) / math.sqrt(fan_in)
    return sd


class WhisperEncoderOracle:
    def __init__(self, cfg: WhisperCfg, sd):
        self.cfg, self.sd = cfg, sd

    def _attn(self, p, x):  # :255-374 (self-attention branch, no masks, eval)
        sd, H = self.sd, self.cfg.encoder_attention_heads
        B, T, D = x.shape
        hs = D // H
        q = F.linear(x, sd[p + "q_proj.weight"], sd[p + "q_proj.bias"]) * hs ** -0.5
        k = F.linear(x, sd[p + "k_proj.weight"])
        v = F.linear(x, sd[p + "v_proj.weight"], sd[p + "v_proj.bias"])

        def shape(t):
            return t.view(B, T, H, hs).transpose(1, 2).contiguous().view(B * H, T, hs)

        w = torch.bmm(shape(q), shape(k).transpose(1, 2))
        w = F.softmax(w, dim=-1)
        o = torch.bmm(w, shape(v)).view(B, H, T, hs).transpose(1, 2).reshape(B, T, D)
        return F.linear(o, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])

    def forward(self, mel):  # (B, num_mel_bins, 2 * max_source_positions) -> (B, max_source_positions, d_model)
        sd, cfg = self.sd, self.cfg
        D = cfg.d_model
        x = F.gelu(F.conv1d(mel, sd["conv1.weight"], sd["conv1.bias"], padding=1))
        x = F.gelu(F.conv1d(x, sd["conv2.weight"], sd["conv2.bias"], stride=2, padding=1))
        h = x.permute(0, 2, 1) + sd["embed_positions.weight"]
        for i in range(cfg.encoder_layers):
            p = f"layers.{i}."
            r = h
            y = F.layer_norm(h, (D,), sd[p + "self_attn_layer_norm.weight"], sd[p + "self_attn_layer_norm.bias"], 1e-5)
            h = r + self._attn(p + "self_attn.", y)
            r = h
            y = F.layer_norm(h, (D,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], 1e-5)
            y = F.gelu(F.linear(y, sd[p + "fc1.weight"], sd[p + "fc1.bias"]))
            h = r + F.linear(y, sd[p + "fc2.weight"], sd[p + "fc2.bias"])
        return F.layer_norm(h, (D,), sd["layer_norm.weight"], sd["layer_norm.bias"], 1e-5)
