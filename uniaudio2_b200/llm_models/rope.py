"""Host-side RoPE table (cos, sin) for the litgpt half-split rotation.

Restates llm_models/lit_model.py:634-706 build_rope_cache incl. the Llama-3 smooth frequency scaling
(:662-676).  Computed once on the CPU in fp32 (like the reference does at module construction) and uploaded;
the kernels index it by position (lit_model.py:129-130).
"""
import torch


def build_rope_cache(seq_len: int, n_elem: int, base: int = 10000, extra_config=None):
    theta = 1.0 / (base ** (torch.arange(0, n_elem, 2).float() / n_elem))
    if extra_config is not None:
        factor = extra_config["factor"]
        if "original_max_seq_len" in extra_config:
            wavelen = 2 * torch.pi / theta
            ratio = extra_config["original_max_seq_len"] / wavelen
            smooth = (ratio - extra_config["low_freq_factor"]) / (
                extra_config["high_freq_factor"] - extra_config["low_freq_factor"])
            smooth = torch.clamp(smooth, min=0.0, max=1.0)
            theta = (1 - smooth) * (theta / factor) + smooth * theta
        else:
            theta = theta / factor
    seq_idx = torch.arange(seq_len) / 1
    idx_theta = torch.outer(seq_idx, theta).repeat(1, 2)
    if idx_theta.shape[-1] > n_elem > 1:
        idx_theta = idx_theta[..., :n_elem]
    return torch.cos(idx_theta).contiguous(), torch.sin(idx_theta).contiguous()
