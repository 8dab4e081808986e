"""Drop-in for ScalarModel, the SQ-codec wave encoder / decoder inside ReasoningCodec_film
(reference: tools/tokenizer/ReasoningCodec_film/models/scalar24k.py:306-425; called from reason_tokenizer.py:215,295,368 as
`self.SQCodec.decode(latent.transpose(1, 2))`).

Same constructor arguments and state-dict keys (weight_norm `weight_g` / `weight_v` pairs, PReLU slopes), same methods
`encode(x)` / `decode(x)`.  Weight norm is folded once when the handle is built; every convolution runs as an implicit /
phase GEMM on the register-tiled fp32 core of libua2_b200.so with the PReLU, bias and residual add fused in the epilogue.
`num_samples > 1` (PreProcessor / PostProcessor pooling) is not on the shipped path and is rejected.
"""
from typing import Dict, Tuple

import torch
import torch.nn as nn

from ..... import _lib


def _tree(shapes: Dict[str, Tuple[int, ...]], device) -> nn.Module:
    root = nn.Module()
    for key, shape in shapes.items():
        parts = key.split(".")
        m = root
        for p in parts[:-1]:
            if not hasattr(m, p):
                m.add_module(p, nn.Module())
            m = getattr(m, p)
        m.register_parameter(parts[-1], nn.Parameter(torch.empty(*shape, device=device), requires_grad=False))
    return root


class ScalarModel(nn.Module):
    def __init__(self, num_bands, sample_rate, causal, num_samples, downsample_factors, downsample_kernel_sizes, upsample_factors,
                 upsample_kernel_sizes, latent_hidden_dim, default_kernel_size, delay_kernel_size, init_channel, res_kernel_size,
                 device=None):
        super().__init__()
        if num_samples != 1:
            raise ValueError("num_samples > 1 (Pre/PostProcessor pooling) is not supported on this path")
        for s, k in zip(upsample_factors, upsample_kernel_sizes):
            if k != 2 * s:
                raise ValueError("transposed convs with kernel != 2 * stride are not supported")
        self.cfg = dict(num_bands=num_bands, causal=bool(causal), down=list(downsample_factors), down_k=list(downsample_kernel_sizes),
                        up=list(upsample_factors), up_k=list(upsample_kernel_sizes), latent=latent_hidden_dim, k=default_kernel_size,
                        delay_k=delay_kernel_size, c0=init_channel, res_k=res_kernel_size)
        sh: Dict[str, Tuple[int, ...]] = {}

        def wn_conv(prefix, cout, cin, k):
            sh[prefix + "bias"] = (cout,)
            sh[prefix + "weight_g"] = (cout, 1, 1)
            sh[prefix + "weight_v"] = (cout, cin, k)

        def res_units(prefix, ch):
            for j in range(5):
                wn_conv(f"{prefix}convs.{j}.conv1.", ch, ch, res_kernel_size)
                wn_conv(f"{prefix}convs.{j}.conv2.", ch, ch, 1)
                sh[f"{prefix}convs.{j}.activation1.weight"] = (1,)
                sh[f"{prefix}convs.{j}.activation2.weight"] = (1,)

        c0, nd, nu = init_channel, len(downsample_factors), len(upsample_factors)
        wn_conv("encoder.0.", c0, num_bands, default_kernel_size)
        for i, (s, k) in enumerate(zip(downsample_factors, downsample_kernel_sizes)):
            cin, cout = c0 * 2 ** i, c0 * 2 ** (i + 1)
            res_units(f"encoder.{1 + i}.", cin)
            wn_conv(f"encoder.{1 + i}.down_conv.layer.", cout, cin, k)
            sh[f"encoder.{1 + i}.down_conv.activation.weight"] = (1,)
        wn_conv(f"encoder.{1 + nd}.", latent_hidden_dim, c0 * 2 ** nd, default_kernel_size)
        wn_conv("decoder.0.", c0 * 2 ** nu, latent_hidden_dim, delay_kernel_size)
        for i, (s, k) in enumerate(zip(upsample_factors, upsample_kernel_sizes)):
            cin, cout = c0 * 2 ** (nu - i), c0 * 2 ** (nu - i - 1)
            sh[f"decoder.{1 + i}.up_conv.layer.bias"] = (cout,)
            sh[f"decoder.{1 + i}.up_conv.layer.weight_g"] = (cin, 1, 1)  # weight_norm dim=0 of the (Cin, Cout, K) weight
            sh[f"decoder.{1 + i}.up_conv.layer.weight_v"] = (cin, cout, k)
            res_units(f"decoder.{1 + i}.", cout)
        wn_conv(f"decoder.{1 + nu}.", num_bands, c0, default_kernel_size)
        tree = _tree(sh, device)
        self.encoder = tree.encoder
        self.decoder = tree.decoder
        self._folded = None

    # ------------------------------------------------------------------ plumbing
    def load_state_dict(self, sd, strict=True, **kw):
        self._folded = None
        return super().load_state_dict(sd, strict=strict, **kw)

    def _fold(self):
        """weight_norm (dim 0): w = g * v / ||v||; transposed convs additionally go to the per-phase GEMM layout."""
        if self._folded is not None:
            return self._folded
        sd = self.state_dict()
        dev = next(iter(sd.values())).device
        if dev.type != "cuda":
            raise _lib.Ua2Error("uniaudio2_b200 ScalarModel runs on a CUDA device only (no CPU fallback)")
        L = _lib.lib()
        f = {}
        with torch.cuda.device(dev):
            for k in sd:
                if k.endswith("weight_v"):
                    p = k[: -len("weight_v")]
                    w = torch._weight_norm(sd[k], sd[p + "weight_g"], 0).contiguous()
                    if ".up_conv." in k:
                        cin, cout, kk = w.shape
                        wp = torch.empty(kk // 2 * cout * cin * 2, device=dev)
                        _lib.check(L.ua2_convtr1d_repack_phase_f32(_lib.ptr(w), _lib.ptr(wp), cin, cout, kk // 2, _lib.current_stream()))
                        f[p + "weight"] = wp
                        f[p + "shape"] = (cin, cout, kk)
                    else:
                        f[p + "weight"] = w
                elif not k.endswith("weight_g"):
                    f[k] = sd[k].contiguous()
            torch.cuda.synchronize(dev)
        self._folded = f
        return f

    def _conv(self, x, prefix, causal, stride=1, dilation=1, prelu=None, residual=None):
        f = self._fold()
        w, b = f[prefix + "weight"], f[prefix + "bias"]
        cout, cin, k = w.shape
        B, _, T = x.shape
        k_eff = (k - 1) * dilation + 1
        if causal:  # scalar24k.py:47-50, :66-68
            pl, pr = dilation * (k - 1), 0
        else:       # get_padding, scalar24k.py:18-19
            pl = pr = int((k * dilation - dilation) / 2)
        T_out = (T + pl + pr - k_eff) // stride + 1
        y = torch.empty(B, cout, T_out, device=x.device)
        _lib.check(_lib.lib().ua2_conv1d_f32(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(prelu), _lib.ptr(residual), _lib.ptr(y), B,
                                             cin, cout, T, k, stride, dilation, pl, pr, _lib.current_stream()), "conv1d")
        return y

    def _convtr(self, x, prefix, stride):
        f = self._fold()
        cin, cout, k = f[prefix + "shape"]
        B, _, T = x.shape
        if self.cfg["causal"]:  # scalar24k.py:102-105: no padding, drop the last `stride` samples
            crop, T_out = 0, T * stride
        else:                   # padding (k - s) // 2 on both sides
            crop = (k - stride) // 2
            T_out = (T - 1) * stride + k - 2 * crop
        y = torch.empty(B, cout, T_out, device=x.device)
        _lib.check(_lib.lib().ua2_convtr1d_f32(_lib.ptr(x), _lib.ptr(f[prefix + "weight"]), _lib.ptr(f[prefix + "bias"]), None, _lib.ptr(y),
                                               B, cin, cout, T, stride, crop, T_out, _lib.current_stream()), "convtr1d")
        return y

    def _res_units(self, x, prefix):
        f = self._fold()
        causal = self.cfg["causal"]
        for j, d in enumerate((1, 3, 5, 7, 9)):  # ResidualUnit, scalar24k.py:139-150
            p = f"{prefix}convs.{j}."
            o = self._conv(x, p + "conv1.", causal, dilation=d, prelu=f[p + "activation1.weight"])
            x = self._conv(o, p + "conv2.", causal, prelu=f[p + "activation2.weight"], residual=x)
        return x

    def _ew(self, x, op, param=0.0):
        y = torch.empty_like(x)
        _lib.check(_lib.lib().ua2_elementwise_f32(_lib.ptr(x), _lib.ptr(y), x.numel(), op, float(param), _lib.current_stream()))
        return y

    # ------------------------------------------------------------------ API
    @torch.inference_mode()
    def decode(self, x: torch.Tensor) -> torch.Tensor:
        """scalar24k.py:403-407.  x (B, latent, T) -> (B, num_bands, T * prod(upsample_factors))."""
        f = self._fold()
        dev = f["decoder.0.bias"].device
        with torch.cuda.device(dev):
            x = self._ew(x.to(device=dev, dtype=torch.float32).contiguous(), 0, 9.0)  # vq: round(9x)/9
            x = self._conv(x, "decoder.0.", False)  # delay conv is never causal (:351-355)
            for i, s in enumerate(self.cfg["up"]):
                x = self._convtr(x, f"decoder.{1 + i}.up_conv.layer.", s)
                x = self._res_units(x, f"decoder.{1 + i}.")
            return self._conv(x, f"decoder.{1 + len(self.cfg['up'])}.", self.cfg["causal"])

    @torch.inference_mode()
    def encode(self, x: torch.Tensor) -> torch.Tensor:
        """scalar24k.py:395-402 (returns the un-quantised tanh embedding, like the reference)."""
        f = self._fold()
        dev = f["encoder.0.bias"].device
        causal = self.cfg["causal"]
        with torch.cuda.device(dev):
            x = self._conv(x.to(device=dev, dtype=torch.float32).contiguous(), "encoder.0.", causal)
            for i, s in enumerate(self.cfg["down"]):
                p = f"encoder.{1 + i}."
                x = self._res_units(x, p)
                x = self._conv(x, p + "down_conv.layer.", causal, stride=s, prelu=f[p + "down_conv.activation.weight"])
            x = self._conv(x, f"encoder.{1 + len(self.cfg['down'])}.", causal)
            return self._ew(x, 1)
