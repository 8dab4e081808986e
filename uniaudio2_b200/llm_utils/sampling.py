"""Drop-in for the reference's llm_utils/sampling.py (sample_token and friends) on CUDA tensors.

    sample_token(logits [*, Card], use_sampling=False, temp=1.0, top_k=0, top_p=0.0) -> LongTensor [*]      sampling.py:84-105
    sample_token_audio(logits, ..., end_token=-1)   ids >= end_token excluded after the softmax                   :107-130
    sample_token_audio_2048(logits, ...)            end_token = 2048                                               :132-154

The whole pipeline (softmax, top-k selection / top-p sort, argmax(p / Exp(1))) is one kernel launch per call
(csrc/ua2_stream.cu::sample_token_kernel).  The Exp(1) draws come from torch's generator on the logits' device with
exactly the shapes the reference draws them in (`torch.empty_like(input_).exponential_(1)`, sampling.py:41): (rows, k)
for top-k (one draw per RANK of the sorted top-k), (rows, Card) otherwise - so a seeded run consumes the generator like the
reference does.  Pass `noise=` to supply the draws explicitly (parity tests).  No torch / CPU fallback.
"""
import torch

from .. import _lib


def _launch(logits, use_sampling, temp, top_k, top_p, end_token, noise):
    if not logits.is_cuda:
        raise _lib.Ua2Error("uniaudio2_b200 sample_token runs on CUDA tensors only (no CPU fallback)")
    card = logits.shape[-1]
    lg = logits.to(torch.float32).contiguous().view(-1, card)
    rows = lg.shape[0]
    sampling = bool(use_sampling) and temp > 0.0
    if sampling:
        n = top_k if (top_k > 0 and not top_p > 0.0) else card
        if noise is None:
            noise = torch.empty(rows, n, device=lg.device, dtype=torch.float32).exponential_(1)
        else:
            noise = noise.to(device=lg.device, dtype=torch.float32).contiguous()
            if noise.numel() != rows * n:
                raise ValueError(f"noise must hold rows x {n} Exp(1) draws")
    else:
        noise = None
    out = torch.empty(rows, dtype=torch.int64, device=lg.device)
    with torch.cuda.device(lg.device):
        _lib.check(_lib.lib().ua2_sample_token_f32(_lib.ptr(lg), rows, card, int(bool(use_sampling)), float(temp), int(top_k),
                                                   float(top_p), int(end_token), _lib.ptr(noise), 0, 0, _lib.ptr(out),
                                                   _lib.current_stream()), "sample_token")
    return out.view(logits.shape[:-1])


def sample_token(logits: torch.Tensor, use_sampling: bool = False, temp: float = 1.0, top_k: int = 0, top_p: float = 0.0,
                 noise: torch.Tensor = None) -> torch.Tensor:
    """Given logits of shape [*, Card], returns a LongTensor of shape [*]."""
    return _launch(logits, use_sampling, temp, top_k, top_p, -1, noise)


def sample_token_audio(logits: torch.Tensor, use_sampling: bool = False, temp: float = 1.0, top_k: int = 0, top_p: float = 0.0,
                       end_token: int = -1, noise: torch.Tensor = None) -> torch.Tensor:
    """sampling.py:107-130: `probs[:, :, :, end_token:] = -inf` (python slice semantics: a negative end_token counts from the
    end, the default -1 removes the last id); needs 4-D logits like the reference's indexing."""
    if logits.dim() != 4:
        raise IndexError("too many indices for tensor of dimension %d" % logits.dim())
    card = logits.shape[-1]
    start = max(card + end_token, 0) if end_token < 0 else min(end_token, card)
    return _launch(logits, use_sampling, temp, top_k, top_p, start, noise)


def sample_token_audio_2048(logits: torch.Tensor, use_sampling: bool = False, temp: float = 1.0, top_k: int = 0,
                            top_p: float = 0.0, noise: torch.Tensor = None) -> torch.Tensor:
    return sample_token_audio(logits, use_sampling, temp, top_k, top_p, end_token=2048, noise=noise)
