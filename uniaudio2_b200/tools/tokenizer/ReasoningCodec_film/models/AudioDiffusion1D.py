"""Drop-in for the solver of the reference's tools/tokenizer/ReasoningCodec_film/models/AudioDiffusion1D.py::BASECFM
(:62-129): `solve_euler` with classifier-free guidance, the whole loop (in-context blend, CFG batch assembly, estimator,
guidance mix, Euler update) on the device through ua2_dit_solve_euler - no host synchronisation between steps."""
import ctypes as C

import torch
import torch.nn as nn

from ..... import _lib
from .transformer_1d_flow import Transformer1DModel


class BASECFM(nn.Module):
    def __init__(self, estimator: Transformer1DModel):
        super().__init__()
        if not isinstance(estimator, Transformer1DModel):
            raise TypeError("estimator must be the uniaudio2_b200 Transformer1DModel")
        self.sigma_min = 1e-4
        self.estimator = estimator

    @torch.inference_mode()
    def solve_euler(self, x, incontext_x, incontext_length, t_span, mu, added_cond_kwargs=None, guidance_scale=1.5):
        """x (1, T, latent) noise, incontext_x (1, T, latent), t_span (n + 1,), mu (1, T, cond) -> (1, T, latent).
        The solution is returned as a new tensor; the caller's `x` is left untouched (the reference overwrites its in-context
        rows in place, AudioDiffusion1D.py:106 - no caller reads them afterwards)."""
        est = self.estimator
        h = est._ensure()
        dev = est.device
        if x.dim() != 3 or x.shape[0] != 1:
            raise RuntimeError("solve_euler runs batch 1 (the reference repeats the timestep twice, AudioDiffusion1D.py:114)")
        if not guidance_scale > 1.0:
            raise ValueError("solve_euler is served with classifier-free guidance (guidance_scale > 1) only")
        T, lat = x.shape[1], x.shape[2]
        cond = est.in_channels - 2 * est.out_channels
        if lat != est.out_channels or tuple(incontext_x.shape) != (1, T, lat) or tuple(mu.shape) != (1, T, cond):
            raise ValueError(f"expected x / incontext_x (1, T, {est.out_channels}) and mu (1, T, {cond})")
        xs = x.to(device=dev, dtype=torch.float32).contiguous().clone()
        ic = incontext_x.to(device=dev, dtype=torch.float32).contiguous()
        m = mu.to(device=dev, dtype=torch.float32).contiguous()
        ts = [float(v) for v in torch.as_tensor(t_span, dtype=torch.float32).cpu().tolist()]
        arr = (C.c_float * len(ts))(*ts)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().ua2_dit_solve_euler(h, _lib.ptr(xs), _lib.ptr(ic), int(incontext_length), arr, len(ts), _lib.ptr(m), T,
                                                      float(guidance_scale), float(self.sigma_min), _lib.current_stream()), "solve_euler")
        return xs
