"""CPU oracle for the SEANet + transformer + residual-VQ codec path (the in-repo 'Mimi twin').
TEST INFRASTRUCTURE ONLY - never imported by the product path.

Restates, as pure torch-CPU fp32 functions over a flat state dict with the reference's own key names, the
algorithm of /root/reference/tools/tokenizer/MimiCodec (byte-identical twin of llm_modules/{seanet,conv,resample,
transformer,rope}.py):
    MimiCodec.encode  model/models/MimiCodec.py:93-101   SEANetEncoder -> ProjectedTransformer -> ConvDownsample1d -> SplitRVQ.encode
    MimiCodec.decode  model/models/MimiCodec.py:103-110  SplitRVQ.decode -> ConvTrUpsample1d -> ProjectedTransformer -> SEANetDecoder
Paths below are relative to tools/tokenizer/MimiCodec/model/.

Parity status: PINNED.  oracle/make_golden_codec.py imports the unmodified reference in the build container and
asserts this restatement is bit-identical to it on CPU (codes and waveform); fixtures live in tests/golden/.
The reference's own streaming-conv self-test (llm_modules/streaming.py:306-358) is re-stated in tests/.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List

import torch
import torch.nn.functional as F


@dataclass
class MimiCfg:
    """Constructor arguments of MimiCodec (models/MimiCodec.py:26-45); defaults = mimi_config.yaml + class defaults."""

    sample_rate: int = 24000
    n_filters: int = 64
    encoder_rates: List[int] = field(default_factory=lambda: [8, 6, 5, 4])
    compress: int = 2
    latent_dim: int = 512
    codebook_size: int = 2048
    codebook_dim: int = 256
    rvq_layers: int = 32
    num_heads: int = 8
    num_layers: int = 8
    layer_scale: float = 0.01
    context: int = 250
    dim_feedforward: int = 2048  # hard-wired in _transformer_kwargs (MimiCodec.py:57)
    target_frame_rate: float = 12.5
    kernel_size: int = 7
    residual_kernel_size: int = 3
    last_kernel_size: int = 3
    max_period: float = 10000.0

    @property
    def hop_length(self):
        return int(math.prod(self.encoder_rates))

    @property
    def resample_stride(self):
        return int((self.sample_rate / self.hop_length) / self.target_frame_rate)  # MimiCodec.py:66-67


# --------------------------------------------------------------------------------------------------------------
# convolutions with the causal / asymmetric padding logic of modules/conv.py
# --------------------------------------------------------------------------------------------------------------
def extra_padding_for_conv1d(length: int, kernel_size: int, stride: int, padding_total: int) -> int:
    """modules/conv.py:50-58 get_extra_padding_for_conv1d."""
    n_frames = (length - kernel_size + padding_total) / stride + 1
    ideal_length = (math.ceil(n_frames) - 1) * stride + (kernel_size - padding_total)
    return ideal_length - length


def conv1d_causal(x, w, b, stride=1, dilation=1, groups=1, pad_mode="constant"):
    """StreamingConv1d.forward, non-streaming causal branch (modules/conv.py:232-254)."""
    k_eff = (w.shape[-1] - 1) * dilation + 1
    padding_total = k_eff - stride
    extra = extra_padding_for_conv1d(x.shape[-1], k_eff, stride, padding_total)
    x = F.pad(x, (padding_total, extra), mode=pad_mode)
    return F.conv1d(x, w, b, stride=stride, dilation=dilation, groups=groups)


def convtr1d_causal(x, w, b, stride, groups=1):
    """StreamingConvTranspose1d.forward, causal, trim_right_ratio=1 (modules/conv.py:306-329)."""
    K = w.shape[-1]
    y = F.conv_transpose1d(x, w, b, stride=stride, groups=groups)
    padding_total = K - stride
    return y[..., : y.shape[-1] - padding_total]


def seanet_encoder(x, sd: Dict[str, torch.Tensor], cfg: MimiCfg, prefix="encoder."):
    """modules/seanet.py:97-241 with the MimiCodec kwargs (n_residual_layers=1, ELU, true_skip, norm none,
    pad_mode constant, causal).  ratios are used reversed (seanet.py:160)."""

    def conv(i, x, **kw):
        return conv1d_causal(x, sd[f"{prefix}model.{i}.conv.conv.weight"], sd[f"{prefix}model.{i}.conv.conv.bias"], **kw)

    x = conv(0, x)
    idx = 1
    for ratio in reversed(cfg.encoder_rates):
        # SEANetResnetBlock (seanet.py:21-94): x + conv_k1(elu(conv_k3(elu(x))))
        p = f"{prefix}model.{idx}.block."
        v = conv1d_causal(F.elu(x), sd[p + "1.conv.conv.weight"], sd[p + "1.conv.conv.bias"])
        v = conv1d_causal(F.elu(v), sd[p + "3.conv.conv.weight"], sd[p + "3.conv.conv.bias"])
        x = x + v
        x = conv(idx + 2, F.elu(x), stride=ratio)  # kernel 2*ratio, stride ratio (seanet.py:196-208)
        idx += 3
    return conv(idx + 1, F.elu(x))


def seanet_decoder(z, sd, cfg: MimiCfg, prefix="decoder."):
    """modules/seanet.py:244-395."""
    x = conv1d_causal(z, sd[f"{prefix}model.0.conv.conv.weight"], sd[f"{prefix}model.0.conv.conv.bias"])
    idx = 1
    for ratio in cfg.encoder_rates:
        x = convtr1d_causal(F.elu(x), sd[f"{prefix}model.{idx + 1}.convtr.convtr.weight"],
                            sd[f"{prefix}model.{idx + 1}.convtr.convtr.bias"], stride=ratio)
        p = f"{prefix}model.{idx + 2}.block."
        v = conv1d_causal(F.elu(x), sd[p + "1.conv.conv.weight"], sd[p + "1.conv.conv.bias"])
        v = conv1d_causal(F.elu(v), sd[p + "3.conv.conv.weight"], sd[p + "3.conv.conv.bias"])
        x = x + v
        idx += 3
    return conv1d_causal(F.elu(x), sd[f"{prefix}model.{idx + 1}.conv.conv.weight"], sd[f"{prefix}model.{idx + 1}.conv.conv.bias"])


# --------------------------------------------------------------------------------------------------------------
# Moshi-family transformer (modules/transformer.py, modules/rope.py), non-streaming forward
# --------------------------------------------------------------------------------------------------------------
def apply_rope_interleaved(q, k, offset: int, max_period: float):
    """modules/rope.py:11-68, layout [B, H, T, D] (time_before_heads=False)."""
    B, H, T, D = q.shape
    ds = torch.arange(D // 2, dtype=torch.float32)
    freqs = torch.exp(ds * (-math.log(max_period) * 2 / D))
    ts = (float(offset) + torch.arange(T, dtype=torch.float32)).view(1, -1, 1)
    q = q.view(B, H, T, D // 2, 2)
    k = k.view(B, H, T, D // 2, 2)
    qr, qi, kr, ki = q[..., 0].float(), q[..., 1].float(), k[..., 0].float(), k[..., 1].float()
    rotr, roti = torch.cos(freqs * ts), torch.sin(freqs * ts)
    qo = torch.stack([qr * rotr - qi * roti, qr * roti + qi * rotr], dim=-1)
    ko = torch.stack([kr * rotr - ki * roti, kr * roti + ki * rotr], dim=-1)
    return qo.view(B, H, T, D), ko.view(B, H, T, D)


def projected_transformer(x, sd, cfg: MimiCfg, prefix):
    """ProjectedTransformer(conv_layout=True, d_model == input_dimension == output_dimension) ->
    StreamingTransformer(positional_embedding='rope', causal, context, layer_scale, gating='none', norm='layer_norm')
    (modules/transformer.py:598-750; layer :430-588; attention :375-419)."""
    x = x.transpose(1, 2)  # [B, T, C]
    B, T, C = x.shape
    H = cfg.num_heads
    pos = torch.arange(T)
    delta = pos.view(-1, 1) - pos.view(1, -1)
    attn_bias = (delta >= 0) & (delta < cfg.context)  # :399-408 (pos_k >= 0 always in non-streaming mode)
    for l in range(cfg.num_layers):
        p = f"{prefix}transformer.layers.{l}."
        h = F.layer_norm(x, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
        proj = F.linear(h, sd[p + "self_attn.in_proj_weight"])
        q, k, v = proj.view(B, T, 3, H, C // H).permute(2, 0, 3, 1, 4)  # "b t (p h d) -> p b h t d"
        q, k = apply_rope_interleaved(q, k, 0, cfg.max_period)
        a = F.scaled_dot_product_attention(q, k, v, attn_bias, dropout_p=0.0)
        a = a.transpose(1, 2).reshape(B, T, C)
        x = x + sd[p + "layer_scale_1.scale"] * F.linear(a, sd[p + "self_attn.out_proj.weight"])
        h = F.layer_norm(x, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
        u = F.linear(F.gelu(F.linear(h, sd[p + "linear1.weight"])), sd[p + "linear2.weight"])
        x = x + sd[p + "layer_scale_2.scale"] * u
    return x.transpose(1, 2)


# --------------------------------------------------------------------------------------------------------------
# residual VQ (quantization/core_vq.py, quantization/vq.py)
# --------------------------------------------------------------------------------------------------------------
def codebook_embedding(sd, prefix, eps=1e-5):
    """EuclideanCodebook.embedding (core_vq.py:142-150): embedding_sum / clamp(cluster_usage, eps)."""
    return sd[prefix + "_codebook.embedding_sum"] / sd[prefix + "_codebook.cluster_usage"].clamp(min=eps)[:, None]


def rvq_encode(x, sd, prefix, n_q):
    """ResidualVectorQuantizer.encode (vq.py:134-145) -> ResidualVectorQuantization.encode (core_vq.py:365-376) ->
    EuclideanCodebook._quantize (core_vq.py:179-185: cdist p=2 + argmin).  x [B, C, T] -> codes [B, n_q, T]."""
    x = F.conv1d(x, sd[prefix + "input_proj.weight"])
    residual = x
    out = []
    for i in range(n_q):
        emb = codebook_embedding(sd, f"{prefix}vq.layers.{i}.")
        r = residual.transpose(1, 2)  # b n d
        flat = r.reshape(-1, r.shape[-1])
        dists = torch.cdist(flat[None], emb[None], p=2)[0]
        codes = dists.argmin(dim=-1).view(r.shape[:-1])
        quantized = F.embedding(codes, emb).transpose(1, 2)
        residual = residual - quantized
        out.append(codes)
    return torch.stack(out).transpose(0, 1)


def rvq_decode(codes, sd, prefix):
    """ResidualVectorQuantizer.decode (vq.py:147-155).  codes [B, n_q, T] -> [B, C_out, T]."""
    q = None
    for i in range(codes.shape[1]):
        emb = codebook_embedding(sd, f"{prefix}vq.layers.{i}.")
        e = F.embedding(codes[:, i], emb).transpose(1, 2)
        q = e if q is None else q + e
    return F.conv1d(q, sd[prefix + "output_proj.weight"])


class MimiOracle:
    def __init__(self, cfg: MimiCfg, sd: Dict[str, torch.Tensor]):
        self.cfg, self.sd = cfg, sd

    def encode_latent(self, wav):
        z = seanet_encoder(wav, self.sd, self.cfg)
        z = projected_transformer(z, self.sd, self.cfg, "encoder_transformer.")
        s = self.cfg.resample_stride
        # ConvDownsample1d(learnt=True, causal, pad_mode='replicate', bias=False) (modules/resample.py:14-65)
        return conv1d_causal(z, self.sd["downsample.conv.conv.conv.weight"], None, stride=s, pad_mode="replicate")

    def encode(self, wav):
        """MimiCodec.encode (MimiCodec.py:93-101) -> SplitResidualVectorQuantizer.encode (vq.py:305-315)."""
        z = self.encode_latent(wav)
        first = rvq_encode(z, self.sd, "quantizer.rvq_first.", 1)
        rest = rvq_encode(z, self.sd, "quantizer.rvq_rest.", self.cfg.rvq_layers - 1)
        return torch.cat([first, rest], dim=1)

    def decode_latent(self, codes):
        q = rvq_decode(codes[:, :1], self.sd, "quantizer.rvq_first.")
        if codes.shape[1] > 1:
            q = q + rvq_decode(codes[:, 1:], self.sd, "quantizer.rvq_rest.")  # vq.py:317-323
        return q

    def decode(self, codes):
        """MimiCodec.decode (MimiCodec.py:103-110)."""
        z = self.decode_latent(codes)
        s = self.cfg.resample_stride
        # ConvTrUpsample1d(learnt=True, channel_wise=True): depthwise transposed conv (modules/resample.py:68-119)
        z = convtr1d_causal(z, self.sd["upsample.convtr.convtr.convtr.weight"], None, stride=s, groups=z.shape[1])
        z = projected_transformer(z, self.sd, self.cfg, "decoder_transformer.")
        return seanet_decoder(z, self.sd, self.cfg)


# --------------------------------------------------------------------------------------------------------------
# seeded weights
# --------------------------------------------------------------------------------------------------------------
def mimi_param_shapes(cfg: MimiCfg):
    s = {}
    nf, D = cfg.n_filters, cfg.latent_dim

    def conv(name, cout, cin, k):
        s[name + ".weight"] = (cout, cin, k)
        s[name + ".bias"] = (cout,)

    # encoder
    conv("encoder.model.0.conv.conv", nf, 1, cfg.kernel_size)
    idx, mult = 1, 1
    for ratio in reversed(cfg.encoder_rates):
        ch = mult * nf
        conv(f"encoder.model.{idx}.block.1.conv.conv", ch // cfg.compress, ch, cfg.residual_kernel_size)
        conv(f"encoder.model.{idx}.block.3.conv.conv", ch, ch // cfg.compress, 1)
        conv(f"encoder.model.{idx + 2}.conv.conv", ch * 2, ch, ratio * 2)
        idx += 3
        mult *= 2
    conv(f"encoder.model.{idx + 1}.conv.conv", D, mult * nf, cfg.last_kernel_size)
    # decoder
    conv("decoder.model.0.conv.conv", mult * nf, D, cfg.kernel_size)
    idx = 1
    for ratio in cfg.encoder_rates:
        ch = mult * nf
        s[f"decoder.model.{idx + 1}.convtr.convtr.weight"] = (ch, ch // 2, ratio * 2)
        s[f"decoder.model.{idx + 1}.convtr.convtr.bias"] = (ch // 2,)
        conv(f"decoder.model.{idx + 2}.block.1.conv.conv", ch // 2 // cfg.compress, ch // 2, cfg.residual_kernel_size)
        conv(f"decoder.model.{idx + 2}.block.3.conv.conv", ch // 2, ch // 2 // cfg.compress, 1)
        idx += 3
        mult //= 2
    conv(f"decoder.model.{idx + 1}.conv.conv", 1, nf, cfg.last_kernel_size)
    st = cfg.resample_stride
    s["downsample.conv.conv.conv.weight"] = (D, D, 2 * st)
    s["upsample.convtr.convtr.convtr.weight"] = (D, 1, 2 * st)
    for t in ("encoder_transformer.", "decoder_transformer."):
        for l in range(cfg.num_layers):
            p = f"{t}transformer.layers.{l}."
            s[p + "self_attn.in_proj_weight"] = (3 * D, D)
            s[p + "self_attn.out_proj.weight"] = (D, D)
            for n in ("norm1", "norm2"):
                s[p + n + ".weight"] = (D,)
                s[p + n + ".bias"] = (D,)
            s[p + "linear1.weight"] = (cfg.dim_feedforward, D)
            s[p + "linear2.weight"] = (D, cfg.dim_feedforward)
            s[p + "layer_scale_1.scale"] = (D,)
            s[p + "layer_scale_2.scale"] = (D,)
    for name, nq in (("quantizer.rvq_first.", 1), ("quantizer.rvq_rest.", cfg.rvq_layers - 1)):
        s[name + "input_proj.weight"] = (cfg.codebook_dim, D, 1)
        s[name + "output_proj.weight"] = (D, cfg.codebook_dim, 1)
        for i in range(nq):
            s[f"{name}vq.layers.{i}._codebook.cluster_usage"] = (cfg.codebook_size,)
            s[f"{name}vq.layers.{i}._codebook.embedding_sum"] = (cfg.codebook_size, cfg.codebook_dim)
    return s


def random_mimi_state_dict(cfg: MimiCfg, seed=0):
    """Seeded synthetic codec weights: conv/linear ~ U(+-1/sqrt(fan_in)) (nn defaults), layer scales 0.01-ish but
    large enough (0.3) for the transformer to matter in parity tests, codebooks ~ N(0,1) with usage in [0.5, 2]."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in mimi_param_shapes(cfg).items():
        if name.endswith("cluster_usage"):
            t = 0.5 + 1.5 * torch.rand(shape, generator=g)
        elif name.endswith("embedding_sum"):
            t = torch.randn(shape, generator=g)
        elif name.endswith("layer_scale_1.scale") or name.endswith("layer_scale_2.scale"):
            t = 0.3 + 0.1 * torch.rand(shape, generator=g)
        elif ".norm" in name and name.endswith(".weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            t = 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = math.prod(shape[1:]) if len(shape) > 1 else shape[0]
            if "convtr" in name and len(shape) == 3:
                fan_in = shape[0] * shape[2] / max(1, 1)
            t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(max(1.0, fan_in))
        sd[name] = t.float()
    return sd
