"""Drop-in for the reference's tools/tokenizer/MimiCodec/model/models/MimiCodec.py::MimiCodec (inference surface).

Same constructor arguments, same state-dict key names (encoder.model.*, decoder.model.*, downsample.*, upsample.*,
encoder_transformer.*, decoder_transformer.*, quantizer.rvq_first.*, quantizer.rvq_rest.*), same methods:
    encode(audio_data (B, 1, T) fp32) -> codes (B, rvq_layers, T') int64        MimiCodec.py:93-101
    decode(codes) -> waveform (B, 1, T' * stride * hop) fp32                    MimiCodec.py:103-110
All arithmetic runs in libua2_b200.so (csrc/ua2_codec*.cu + the skinny-linear / attention kernels); this module owns
the parameters (torch tensors on the GPU) and the native handle.  No torch/CPU fallback.
"""
import ctypes as C
import math

import torch
import torch.nn as nn

from .... import _lib


class _P(nn.Module):
    """Leaf parameter holder so that state-dict keys nest like the reference's modules."""

    def __init__(self, **tensors):
        super().__init__()
        for name, t in tensors.items():
            if name.startswith("buf_"):
                self.register_buffer(name[4:], t)
            else:
                setattr(self, name, nn.Parameter(t, requires_grad=False))


def _seq(mods: dict) -> nn.Module:
    m = nn.Module()
    for k, v in mods.items():
        m.add_module(str(k), v)
    return m


def _conv(cout, cin, k, device, bias=True):
    inner = _P(weight=torch.empty(cout, cin, k, device=device), **({"bias": torch.empty(cout, device=device)} if bias else {}))
    return _seq({"conv": _seq({"conv": inner})})  # StreamingConv1d.conv (NormConv1d) .conv (nn.Conv1d)


def _convtr(cin, cout, k, device):
    inner = _P(weight=torch.empty(cin, cout, k, device=device), bias=torch.empty(cout, device=device))
    return _seq({"convtr": _seq({"convtr": inner})})


def _resblock(ch, hidden, k, device):
    return _seq({"block": _seq({1: _conv(hidden, ch, k, device), 3: _conv(ch, hidden, 1, device)})})


class MimiCodec(nn.Module):
    def __init__(self, sample_rate=24000, n_filters=64, encoder_rates=[4, 5, 6, 8], compress=2, causal=True, latent_dim=512,
                 codebook_size=4096, codebook_dim=32, rvq_layers=8, num_heads=8, num_layers=8, layer_scale=0.01, context=250,
                 dim_feedforward=2048, semantic_feature_dim=1024, target_frame_rate=12.5, device=None):
        super().__init__()
        if not causal:
            raise ValueError("only the causal configuration used by MimiCodec is on this path")
        self.sample_rate = sample_rate
        self.encoder_rates = list(encoder_rates)
        self.hop_length = int(math.prod(self.encoder_rates))
        self.encoder_frame_rate = 24000 / self.hop_length  # MimiCodec.py:65 (hard-coded 24000 in the reference too)
        self.target_frame_rate = target_frame_rate
        self.resample_stride = int(self.encoder_frame_rate / self.target_frame_rate)
        self.cfg = dict(n_filters=n_filters, latent_dim=latent_dim, codebook_size=codebook_size, codebook_dim=codebook_dim,
                        rvq_layers=rvq_layers, num_heads=num_heads, num_layers=num_layers, context=context,
                        dim_feedforward=2048)  # _transformer_kwargs hard-wires 2048 (MimiCodec.py:57)
        nf, D, dev = n_filters, latent_dim, device
        # ---- SEANetEncoder / Decoder parameter trees (modules/seanet.py:97-241, :244-395; index layout of nn.Sequential)
        enc, idx, mult = {0: _conv(nf, 1, 7, dev)}, 1, 1
        for ratio in reversed(self.encoder_rates):
            ch = mult * nf
            enc[idx] = _resblock(ch, ch // compress, 3, dev)
            enc[idx + 2] = _conv(ch * 2, ch, ratio * 2, dev)
            idx += 3
            mult *= 2
        enc[idx + 1] = _conv(D, mult * nf, 3, dev)
        self.encoder = _seq({"model": _seq(enc)})
        dec, idx = {0: _conv(mult * nf, D, 7, dev)}, 1
        for ratio in self.encoder_rates:
            ch = mult * nf
            dec[idx + 1] = _convtr(ch, ch // 2, ratio * 2, dev)
            dec[idx + 2] = _resblock(ch // 2, ch // 2 // compress, 3, dev)
            idx += 3
            mult //= 2
        dec[idx + 1] = _conv(1, nf, 3, dev)
        self.decoder = _seq({"model": _seq(dec)})
        st = self.resample_stride
        self.downsample = _seq({"conv": _conv(D, D, 2 * st, dev, bias=False)})
        self.upsample = _seq({"convtr": _seq({"convtr": _seq({"convtr": _P(weight=torch.empty(D, 1, 2 * st, device=dev))})})})
        for name in ("encoder_transformer", "decoder_transformer"):
            layers = {}
            for l in range(num_layers):
                lay = nn.Module()
                lay.self_attn = _P(in_proj_weight=torch.empty(3 * D, D, device=dev))
                lay.self_attn.out_proj = _P(weight=torch.empty(D, D, device=dev))
                lay.norm1 = _P(weight=torch.empty(D, device=dev), bias=torch.empty(D, device=dev))
                lay.norm2 = _P(weight=torch.empty(D, device=dev), bias=torch.empty(D, device=dev))
                lay.linear1 = _P(weight=torch.empty(2048, D, device=dev))
                lay.linear2 = _P(weight=torch.empty(D, 2048, device=dev))
                lay.layer_scale_1 = _P(scale=torch.full((D,), float(layer_scale), device=dev))
                lay.layer_scale_2 = _P(scale=torch.full((D,), float(layer_scale), device=dev))
                layers[l] = lay
            setattr(self, name, _seq({"transformer": _seq({"layers": _seq(layers)})}))
        q = nn.Module()
        for name, nq in (("rvq_first", 1), ("rvq_rest", rvq_layers - 1)):
            r = nn.Module()
            r.input_proj = _P(weight=torch.empty(codebook_dim, D, 1, device=dev))
            r.output_proj = _P(weight=torch.empty(D, codebook_dim, 1, device=dev))
            r.vq = _seq({"layers": _seq({i: _seq({"_codebook": _P(
                buf__initialized=torch.ones(1, device=dev), buf_cluster_usage=torch.ones(codebook_size, device=dev),
                buf_embedding_sum=torch.zeros(codebook_size, codebook_dim, device=dev))}) for i in range(nq)})})
            setattr(q, name, r)
        self.quantizer = q
        self._h = None
        self._keep = []

    # ------------------------------------------------------------------ native handle
    def _destroy(self):
        if getattr(self, "_h", None) is not None:
            _lib.lib().ua2_codec_destroy(self._h)
            self._h = None
            self._keep = []

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    def load_state_dict(self, sd, strict=True, **kw):
        # the reference checkpoint also carries the training-time distillation head (semantic_mapping_layer.*)
        sd = {k: v for k, v in sd.items() if not k.startswith("semantic_mapping_layer.")}
        self._destroy()
        return super().load_state_dict(sd, strict=strict, **kw)

    def _apply(self, fn, *args, **kwargs):
        # .to() / .cuda() / .float() move the tensors the handle points at: rebuild it lazily on the next encode / decode
        self._destroy()
        return super()._apply(fn, *args, **kwargs)

    def _ensure(self):
        if self._h is not None:
            return
        L = _lib.lib()
        dev = self.upsample.convtr.convtr.convtr.weight.device
        if dev.type != "cuda":
            raise _lib.Ua2Error("uniaudio2_b200 codec runs on a CUDA device only (no CPU fallback): call .to('cuda') first")
        c = self.cfg
        ratios = (C.c_int32 * 8)(*(self.encoder_rates + [0] * (8 - len(self.encoder_rates))))
        cfg = _lib.CodecCfg(c["n_filters"], ratios, len(self.encoder_rates), c["latent_dim"], c["codebook_size"], c["codebook_dim"],
                            c["rvq_layers"], c["num_heads"], c["num_layers"], c["context"], c["dim_feedforward"],
                            self.resample_stride, 10000.0)
        h = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(L.ua2_codec_create(C.byref(cfg), C.byref(h)), "ua2_codec_create")
            keep = []
            for key, t in self.state_dict().items():
                t = t.detach()
                if t.dtype != torch.float32:
                    raise _lib.Ua2Error(f"{key} has dtype {t.dtype}; this path computes in fp32")
                t = t.contiguous()
                keep.append(t)
                shape = (C.c_int64 * t.dim())(*t.shape)
                _lib.check(L.ua2_codec_load_weight(h, key.encode(), _lib.ptr(t), shape, t.dim()), f"load_weight({key})")
            _lib.check(L.ua2_codec_finalize(h, _lib.current_stream()), "ua2_codec_finalize")
        self._h, self._keep = h, keep

    # ------------------------------------------------------------------ API
    @torch.inference_mode()
    def encode(self, audio_data: torch.Tensor) -> torch.Tensor:
        self._ensure()
        if audio_data.dim() != 3 or audio_data.shape[1] != 1:
            raise ValueError("expected audio of shape (B, 1, T)")
        dev = self._keep[0].device
        x = audio_data.to(device=dev, dtype=torch.float32).contiguous()
        B, _, T = x.shape
        L = _lib.lib()
        Tq = int(L.ua2_codec_frames(self._h, T))
        codes = torch.empty(B, self.cfg["rvq_layers"], Tq, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.ua2_codec_encode(self._h, _lib.ptr(x), B, T, _lib.ptr(codes), _lib.current_stream()), "encode")
        return codes

    @torch.inference_mode()
    def decode(self, codes: torch.Tensor) -> torch.Tensor:
        self._ensure()
        if codes.dim() != 3 or codes.shape[1] != self.cfg["rvq_layers"]:
            raise ValueError("expected codes of shape (B, rvq_layers, T)")
        dev = self._keep[0].device
        cds = codes.to(device=dev, dtype=torch.int64).contiguous()
        B, _, Tq = cds.shape
        wav = torch.empty(B, 1, Tq * self.resample_stride * self.hop_length, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().ua2_codec_decode(self._h, _lib.ptr(cds), B, Tq, _lib.ptr(wav), _lib.current_stream()), "decode")
        return wav

    @classmethod
    def from_config(cls, config_path):  # MimiCodec.py:112-117
        import json

        with open(config_path, "r") as f:
            return cls(**json.load(f))
