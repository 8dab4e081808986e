"""CPU oracle for ScalarModel (the SQ-codec wave encoder/decoder inside ReasoningCodec_film).  TEST INFRASTRUCTURE ONLY.

Restates /root/reference/tools/tokenizer/ReasoningCodec_film/models/scalar24k.py::ScalarModel (:306-425) as pure torch-CPU
functions over a flat state dict with the reference's key names (weight_norm parameters `weight_g` / `weight_v` included).
Pinned by oracle/make_golden_scalar.py against the real class (imported with pytorch_lightning / omegaconf stubs).
The shipped configuration (sqcodec_config.yaml) is NOT in the repository: the config below is assumed
and the parity claim covers the algorithm, not the production hyper-parameters.
"""
from dataclasses import dataclass, field
from typing import Dict, List

import torch
import torch.nn.functional as F


@dataclass
class ScalarCfg:
    num_bands: int = 1
    sample_rate: int = 24000
    causal: bool = True
    num_samples: int = 1
    downsample_factors: List[int] = field(default_factory=lambda: [2, 4, 4, 5, 6])
    downsample_kernel_sizes: List[int] = field(default_factory=lambda: [4, 8, 8, 10, 12])
    upsample_factors: List[int] = field(default_factory=lambda: [6, 5, 4, 4, 2])
    upsample_kernel_sizes: List[int] = field(default_factory=lambda: [12, 10, 8, 8, 4])
    latent_hidden_dim: int = 136
    default_kernel_size: int = 7
    delay_kernel_size: int = 5
    init_channel: int = 48
    res_kernel_size: int = 7


def fold_wn(sd, prefix):
    """torch.nn.utils.weight_norm (dim=0): w = g * v / ||v|| with the norm over every dim but 0."""
    v, g = sd[prefix + "weight_v"], sd[prefix + "weight_g"]
    return torch._weight_norm(v, g, 0)


def conv(x, w, b, causal, stride=1, dilation=1):
    """scalar24k.py:30-69 Conv1d: causal -> left pad dilation*(k-1), else symmetric get_padding (:18-19)."""
    k = w.shape[-1]
    if causal:
        x = F.pad(x, (dilation * (k - 1), 0))
        return F.conv1d(x, w, b, stride=stride, dilation=dilation)
    return F.conv1d(x, w, b, stride=stride, dilation=dilation, padding=int((k * dilation - dilation) / 2))


def convtr(x, w, b, causal, stride):
    """scalar24k.py:71-106 ConvTranspose1d: causal -> no padding, drop the last `stride` samples; else padding (k-s)//2."""
    k = w.shape[-1]
    if causal:
        return F.conv_transpose1d(x, w, b, stride=stride)[:, :, :-stride]
    return F.conv_transpose1d(x, w, b, stride=stride, padding=(k - stride) // 2)


def res_unit(x, sd, p, dilation, causal):
    """ResidualUnit (:139-150): PReLU(conv1(x)) -> PReLU(conv2(.)) + x, both convs weight-normed."""
    o = F.prelu(conv(x, fold_wn(sd, p + "conv1."), sd[p + "conv1.bias"], causal, dilation=dilation), sd[p + "activation1.weight"])
    o = F.prelu(conv(o, fold_wn(sd, p + "conv2."), sd[p + "conv2.bias"], causal), sd[p + "activation2.weight"])
    return o + x


def scalar_decode(z, sd: Dict[str, torch.Tensor], cfg: ScalarCfg):
    """ScalarModel.decode (:403-407): round(9x)/9 -> delay conv (never causal, :351-355) -> ResDecoderBlocks -> last conv."""
    assert cfg.num_samples == 1
    x = torch.round(9 * z) / 9
    x = conv(x, fold_wn(sd, "decoder.0."), sd["decoder.0.bias"], False)
    n = len(cfg.upsample_factors)
    for i, s in enumerate(cfg.upsample_factors):
        p = f"decoder.{1 + i}."
        x = convtr(x, fold_wn(sd, p + "up_conv.layer."), sd[p + "up_conv.layer.bias"], cfg.causal, s)  # activation=None (:171)
        for j, d in enumerate((1, 3, 5, 7, 9)):
            x = res_unit(x, sd, f"{p}convs.{j}.", d, cfg.causal)
    p = f"decoder.{1 + n}."
    return conv(x, fold_wn(sd, p), sd[p + "bias"], cfg.causal)


def scalar_encode(wav, sd, cfg: ScalarCfg):
    """ScalarModel.encode (:395-402): conv -> ResEncoderBlocks (5 res units + strided down conv + PReLU) -> tanh(conv)."""
    assert cfg.num_samples == 1
    x = conv(wav, fold_wn(sd, "encoder.0."), sd["encoder.0.bias"], cfg.causal)
    n = len(cfg.downsample_factors)
    for i, s in enumerate(cfg.downsample_factors):
        p = f"encoder.{1 + i}."
        for j, d in enumerate((1, 3, 5, 7, 9)):
            x = res_unit(x, sd, f"{p}convs.{j}.", d, cfg.causal)
        x = conv(x, fold_wn(sd, p + "down_conv.layer."), sd[p + "down_conv.layer.bias"], cfg.causal, stride=s)
        x = F.prelu(x, sd[p + "down_conv.activation.weight"])
    p = f"encoder.{1 + n}."
    return torch.tanh(conv(x, fold_wn(sd, p), sd[p + "bias"], cfg.causal))
