"""Drop-in for the reference's llm_models/model_new.py::Model_stage3 on the B200-native path.

Same constructor argument (`ModelArgs`), same state-dict key names, same methods and argument meaning:
    setup_caches(max_batch_size)                         model_new.py:554-565
    reset_caches()                                       model_new.py:647-651
    forward_prefix(tokens, labels, tokens_mask, ...)     model_new.py:456-507   (KV-cache side effect)
    generate_frame(tokens, tokens_mask, input_pos, ...)  model_new.py:568-645   -> (B, 1+nq) int32
All arithmetic runs in libua2_b200.so (hand-written sm_100a kernels, see csrc/); this module only owns the
parameters (torch tensors on the GPU), the RoPE tables and the ctypes handle.  There is no torch/CPU fallback.
"""
import ctypes as C
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.nn as nn

from .. import _lib
from .config import Config as gpt_config
from .rope import build_rope_cache


@dataclass
class ModelArgs:  # model_new.py:190-199
    llm_name: str
    decoder_name: str
    llm_pretrained_model: str
    audio_embeddings_path: str
    audio_understanding_expert_path: str
    audio_semantic_vocab_size: int
    audio_reason_vocab_size: int
    audio_num_codebooks: int


class _Weight(nn.Module):
    def __init__(self, *shape, device=None):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(*shape, dtype=torch.float32, device=device), requires_grad=False)


class _Attn(nn.Module):
    def __init__(self, c: gpt_config, device):
        super().__init__()
        self.qkv = _Weight((c.n_head + 2 * c.n_query_groups) * c.head_size, c.n_embd, device=device)
        self.proj = _Weight(c.n_embd, c.head_size * c.n_head, device=device)


class _MLP(nn.Module):
    def __init__(self, c: gpt_config, device):
        super().__init__()
        self.fc_1 = _Weight(c.intermediate_size, c.n_embd, device=device)
        self.fc_2 = _Weight(c.intermediate_size, c.n_embd, device=device)
        self.proj = _Weight(c.n_embd, c.intermediate_size, device=device)


class _Block(nn.Module):
    def __init__(self, c: gpt_config, device):
        super().__init__()
        self.norm_1 = _Weight(c.n_embd, device=device)
        self.attn = _Attn(c, device)
        self.norm_2 = _Weight(c.n_embd, device=device)
        self.mlp = _MLP(c, device)


class GPT(nn.Module):
    """Parameter container with the key layout of lit_model.py::GPT (:22-37).  `with_embed=False` reproduces
    _prepare_transformer (model_new.py:111-115: wte / lm_head replaced by Identity -> no parameters)."""

    def __init__(self, config: gpt_config, with_embed: bool, device=None):
        super().__init__()
        self.config = config
        if with_embed:
            self.lm_head = _Weight(config.padded_vocab_size, config.n_embd, device=device)
        mods = dict(h=nn.ModuleList(_Block(config, device) for _ in range(config.n_layer)),
                    ln_f=_Weight(config.n_embd, device=device))
        if with_embed:
            mods["wte"] = _Weight(config.padded_vocab_size, config.n_embd, device=device)
        self.transformer = nn.ModuleDict(mods)


def _gpt_cfg_struct(c: gpt_config) -> _lib.GptCfg:
    return _lib.GptCfg(c.n_layer, c.n_embd, c.n_head, c.n_query_groups, c.head_size, c.intermediate_size, c.norm_eps)


class Model_stage3(nn.Module):
    """Stage 3 text-audio model: understanding expert (3 L) -> backbone (28 L) -> generation expert (2 L) ->
    text head + 8-step local decoder with per-codebook heads (model_new.py:334-355)."""

    MAX_SEQ_LENGTH = 2048  # model_new.py:560-565

    def __init__(self, config: ModelArgs, device=None, max_seq_length: Optional[int] = None):
        super().__init__()
        self.config = config
        llm_config = gpt_config.from_name(config.llm_name)
        dec_config = gpt_config.from_name(config.decoder_name)
        und_config = gpt_config.from_name("meta-llama/Llama-3.2-Understanding")
        gen_config = gpt_config.from_name("meta-llama/Llama-3.2-Generation")
        self.backbone = GPT(llm_config, True, device)
        self.decoder = GPT(dec_config, False, device)
        V = config.audio_semantic_vocab_size + config.audio_reason_vocab_size
        self.audio_embeddings = _Weight(V * config.audio_num_codebooks, llm_config.n_embd, device=device)
        self.projection = _Weight(dec_config.n_embd, llm_config.n_embd, device=device)
        self.audio_head = nn.Parameter(torch.empty(config.audio_num_codebooks, dec_config.n_embd, V, dtype=torch.float32,
                                                   device=device), requires_grad=False)
        self.audio_understanding_expert = GPT(und_config, False, device)
        self.audio_generation_expert = GPT(gen_config, False, device)
        self._max_seq_length = max_seq_length or self.MAX_SEQ_LENGTH
        self._h = None  # ua2_llm handle
        self._keep = []  # tensors whose device memory the handle references
        self._max_batch = 0
        self.rng_mode = "torch"  # "torch": draw Exp(1) like model_new.py:141-143 (9 calls/frame); "philox": in-kernel
        self.seed = 888  # multi_task_inference.py:596
        self._noise = None
        self._out = None

    # ------------------------------------------------------------------ handle management
    def _stacks(self):
        return (("backbone.", self.backbone), ("decoder.", self.decoder),
                ("audio_understanding_expert.", self.audio_understanding_expert),
                ("audio_generation_expert.", self.audio_generation_expert))

    def _destroy(self):
        if getattr(self, "_h", None) is not None:
            _lib.lib().ua2_llm_destroy(self._h)
            self._h = None
            self._keep = []

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    # The handle holds raw parameter pointers plus a transposed copy of audio_head made in setup_caches: parameters that change
    # (load_state_dict, resume_for_inference) or move (.to / .cuda / .float) after setup_caches would leave it stale.  Both drop the
    # handle; the next forward_prefix / generate_frame raises the reference's "You need to call `setup_caches()`" (lit_model.py:134-135).
    def load_state_dict(self, *args, **kwargs):
        self._destroy()
        self._max_batch = 0
        return super().load_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self._destroy()
        self._max_batch = 0
        return super()._apply(fn, *args, **kwargs)

    def setup_caches(self, max_batch_size: int) -> None:
        """model_new.py:554-565: KV caches for the three global stacks at 2048 slots and the local decoder at
        audio_num_codebooks slots.  Also binds the (GPU-resident, fp32) parameters to the native handle."""
        L = _lib.lib()
        dev = self.audio_head.device
        if dev.type != "cuda":
            raise _lib.Ua2Error("uniaudio2_b200 runs on a CUDA device only (no CPU fallback): call model.to('cuda') first")
        for n, p in self.named_parameters():
            if p.dtype != torch.float32:
                raise _lib.Ua2Error(f"parameter {n} has dtype {p.dtype}; this path computes in fp32 like the reference")
        self._destroy()
        cfgs = [m.config for _, m in self._stacks()]
        V = self.config.audio_semantic_vocab_size + self.config.audio_reason_vocab_size
        cfg = _lib.LlmCfg(_gpt_cfg_struct(cfgs[0]), _gpt_cfg_struct(cfgs[1]), _gpt_cfg_struct(cfgs[2]),
                          _gpt_cfg_struct(cfgs[3]), cfgs[0].padded_vocab_size, V, self.config.audio_num_codebooks,
                          self._max_seq_length)
        h = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(L.ua2_llm_create(C.byref(cfg), C.byref(h)), "ua2_llm_create")
            self._h = h
            keep = []

            def reg(key, t):
                t = t.detach()
                if not t.is_contiguous():
                    t = t.contiguous()
                keep.append(t)
                shape = (C.c_int64 * t.dim())(*t.shape)
                _lib.check(L.ua2_llm_load_weight(h, key.encode(), _lib.ptr(t), shape, t.dim()), f"load_weight({key})")

            for name, p in self.named_parameters():
                reg(name, p)
            for prefix, m in self._stacks():
                c = m.config
                n_pos = self.config.audio_num_codebooks if prefix == "decoder." else self._max_seq_length
                cos, sin = build_rope_cache(n_pos, c.rope_n_elem, c.rope_base, c.rope_adjustments)
                reg(prefix + "rope_cos", cos.to(dev))
                reg(prefix + "rope_sin", sin.to(dev))
            _lib.check(L.ua2_llm_setup_caches(h, int(max_batch_size), _lib.current_stream()), "setup_caches")
        self._keep = keep
        self._max_batch = int(max_batch_size)
        nq = self.config.audio_num_codebooks
        self._out = torch.zeros(max_batch_size, nq + 1, dtype=torch.int32, device=dev)
        self._noise = torch.empty(max_batch_size * (cfgs[0].padded_vocab_size + nq * V), dtype=torch.float32, device=dev)

    def _require_handle(self):
        if self._h is None:
            raise TypeError("You need to call `setup_caches()`")  # lit_model.py:134-135

    def reset_caches(self):
        """model_new.py:647-651."""
        self._require_handle()
        with torch.cuda.device(self.audio_head.device):
            _lib.check(_lib.lib().ua2_llm_reset_caches(self._h, _lib.current_stream()), "reset_caches")

    def set_option(self, name: str, value: int):
        self._require_handle()
        _lib.check(_lib.lib().ua2_llm_set_option(self._h, name.encode(), int(value)), "set_option")

    def last_launch_count(self) -> int:
        return _lib.lib().ua2_llm_last_launch_count(self._h)

    # ------------------------------------------------------------------ forward paths
    @torch.inference_mode()
    def forward_prefix(self, tokens: torch.Tensor, labels: torch.Tensor = None, tokens_mask: torch.Tensor = None,
                       loss_mask: torch.Tensor = None, input_pos=None, input_pos_maxp1=None):
        """model_new.py:456-507.  tokens (B,S-1,nq+1), tokens_mask (B,S,nq+1) (the reference slices [:, :-1]),
        input_pos (B,S-1).  Only the KV-cache side effect is produced: every caller discards the returned
        logits (e.g. evaluation/tts_task.py:244), so the lm_head / cache-less local-decoder pass is skipped
        and (None, None, None, None) is returned in place of (text_logits, ci_logits, label, mask)."""
        self._require_handle()
        B, T, C1 = tokens.shape
        if C1 != self.config.audio_num_codebooks + 1:
            raise ValueError("last stream must be text")
        if input_pos is None:
            raise ValueError("forward_prefix needs input_pos (the KV-cache path, lit_model.py:123)")
        dev = self.audio_head.device
        tok = tokens.to(device=dev, dtype=torch.int64).contiguous()
        msk = tokens_mask[:, :T].to(device=dev, dtype=torch.uint8).contiguous()
        pos = input_pos.to(device=dev, dtype=torch.int64)
        if pos.dim() == 1:
            pos = pos.unsqueeze(0).expand(B, T)
        if pos.shape[-1] != T:
            raise ValueError(f"input_pos.shape[-1] = {pos.shape[-1]} != {T} = idx.shape[1], must be the same")  # lit_model.py:127-128
        pos = pos.contiguous()
        max_pos = int(pos.max().item()) if input_pos_maxp1 is None else int(input_pos_maxp1) - 1
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().ua2_llm_prefill(self._h, _lib.ptr(tok), _lib.ptr(msk), _lib.ptr(pos), B, T, max_pos,
                                                  _lib.current_stream()), "forward_prefix")
        return None, None, None, None

    def draw_noise(self, rows: int) -> torch.Tensor:
        """The Exp(1) draws of model_new.py:141-143, taken from torch's generator on the model device in the
        same order and shapes as the reference would (text head first, then the 8 codebook heads)."""
        nq = self.config.audio_num_codebooks
        Vt = self.backbone.config.padded_vocab_size
        Va = self.config.audio_semantic_vocab_size + self.config.audio_reason_vocab_size
        buf = self._noise
        buf[: rows * Vt].view(rows, Vt).exponential_(1)
        off = rows * Vt
        for _ in range(nq):
            buf[off: off + rows * Va].view(rows, Va).exponential_(1)
            off += rows * Va
        return buf

    @torch.inference_mode()
    def tts_frames(self, tokens, tokens_mask, input_pos: int, n_frames: int, state: torch.Tensor, frames_out: torch.Tensor,
                   temperature: float, topk: int, reason_eos: int, end_tok: int, reason_card: int, fixed_switch: int = -1):
        """`n_frames` consecutive frames of the TTS hot loop (evaluation/tts_task.py:253-279, B = 1) with the feedback of the sample and
        the phase / EOS state machine on the device (ua2_llm_tts_frames): no host round trip between frames.  tokens / tokens_mask
        (1, 1, nq+1): the prompt's last row for the first frame of an utterance, None to continue from the previous sample.  state: int32[4]
        on the device (zero it per utterance); frames_out: int32 (cap, 1+nq) on the device."""
        self._require_handle()
        dev = self.audio_head.device
        nq = self.config.audio_num_codebooks
        tok = msk = None
        if tokens is not None:
            tok = tokens.to(device=dev, dtype=torch.int64).contiguous()
            msk = tokens_mask.to(device=dev, dtype=torch.uint8).contiguous()
            assert tok.numel() == nq + 1 and msk.numel() == nq + 1, "one row of nq + 1 streams (B = 1)"
        noise, stride = None, 0
        with torch.cuda.device(dev):
            if self.rng_mode == "torch":  # the same draws, in the same order, as n_frames calls of generate_frame would take
                Vt = self.backbone.config.padded_vocab_size
                Va = self.config.audio_semantic_vocab_size + self.config.audio_reason_vocab_size
                stride = Vt + nq * Va
                if getattr(self, "_noise_frames", None) is None or self._noise_frames.numel() < n_frames * stride:
                    self._noise_frames = torch.empty(n_frames * stride, dtype=torch.float32, device=dev)
                noise = self._noise_frames
                for f in range(n_frames):
                    base = f * stride
                    noise[base: base + Vt].view(1, Vt).exponential_(1)
                    for i in range(nq):
                        noise[base + Vt + i * Va: base + Vt + (i + 1) * Va].view(1, Va).exponential_(1)
            _lib.check(_lib.lib().ua2_llm_tts_frames(self._h, _lib.ptr(tok), _lib.ptr(msk), int(input_pos), int(n_frames), float(temperature),
                                                     int(topk), _lib.ptr(noise), int(stride), int(self.seed), int(reason_eos), int(end_tok),
                                                     int(reason_card), int(fixed_switch), _lib.ptr(state), _lib.ptr(frames_out),
                                                     int(frames_out.shape[0]), _lib.current_stream()), "tts_frames")

    @torch.inference_mode()
    def generate_frame(self, tokens: torch.Tensor, tokens_mask: torch.Tensor, input_pos: torch.Tensor,
                       input_pos_maxp1=None, temperature: float = 1.0, topk: int = 1, forbid_prefix: int = 0,
                       cfg_scale: float = 1.0, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """model_new.py:568-645.  tokens / tokens_mask (B,1,nq+1); input_pos 1-element tensor (or int) shared by
        all rows (tts_task.py:245).  Returns (B, 1+nq) int32 on the model's device (col 0 = text token).
        `noise` optionally injects the Exp(1) draws (layout of ua2_llm_generate_frame)."""
        self._require_handle()
        B, S, C1 = tokens.shape
        nq = self.config.audio_num_codebooks
        assert C1 == nq + 1, "last stream must be text"
        if S != 1:
            raise ValueError("generate_frame processes one frame per call (S == 1), as every reference task loop does")
        dev = self.audio_head.device
        tok = tokens.to(device=dev, dtype=torch.int64).contiguous()
        msk = tokens_mask.to(device=dev, dtype=torch.uint8).contiguous()
        if isinstance(input_pos, torch.Tensor):
            if input_pos.numel() != 1:
                raise ValueError("generate_frame expects a single shared position")
            pos = int(input_pos.item()) if input_pos.device.type == "cpu" else int(input_pos.reshape(-1)[0].item())
        else:
            pos = int(input_pos)
        if B > self._max_batch:
            raise ValueError(f"batch size {B} exceeds setup_caches({self._max_batch})")
        use_cfg = cfg_scale > 1.0 and B > 1
        rows = 1 if use_cfg else B
        with torch.cuda.device(dev):
            if noise is None and self.rng_mode == "torch":
                noise = self.draw_noise(rows)
            out = self._out[:B]
            _lib.check(_lib.lib().ua2_llm_generate_frame(self._h, _lib.ptr(tok), _lib.ptr(msk), B, pos, float(temperature),
                                                         int(topk), int(forbid_prefix), float(cfg_scale), _lib.ptr(noise),
                                                         int(self.seed), _lib.ptr(out), _lib.current_stream()),
                       "generate_frame")
        return out.clone()

    # ------------------------------------------------------------------ introspection for parity tests
    def kv_cache(self, which: int, layer: int):
        """(k, v) views (B, G, S_max, hs) of a stack's cache. which: 0 backbone 1 decoder 2 understanding 3 generation."""
        self._require_handle()
        k, v = C.c_void_p(), C.c_void_p()
        _lib.check(_lib.lib().ua2_llm_get_kv(self._h, which, layer, C.byref(k), C.byref(v)))
        m = [self.backbone, self.decoder, self.audio_understanding_expert, self.audio_generation_expert][which]
        c = m.config
        S = self.config.audio_num_codebooks if which == 1 else self._max_seq_length
        shape = (self._max_batch, c.n_query_groups, S, c.head_size)
        return _from_ptr(k.value, shape, self.audio_head.device), _from_ptr(v.value, shape, self.audio_head.device)

    def debug_buffer(self, name: str, B: int):
        self._require_handle()
        p, n = C.c_void_p(), C.c_int64()
        _lib.check(_lib.lib().ua2_llm_get_buffer(self._h, name.encode(), C.byref(p), C.byref(n)))
        nq = self.config.audio_num_codebooks
        if name == "h_final":
            shape = (self._max_batch, self.backbone.config.n_embd)
        elif name == "text_logits":
            shape = (self._max_batch, self.backbone.config.padded_vocab_size)
        else:
            shape = (n.value,)
        t = _from_ptr(p.value, shape, self.audio_head.device)
        if name == "audio_logits":
            Va = self.config.audio_semantic_vocab_size + self.config.audio_reason_vocab_size
            return t[: nq * B * Va].view(nq, B, Va)
        if name in ("h_final", "text_logits"):
            return t[:B]
        return t  # raw word buffer (e.g. "chain_prof")

    def get_fsdp_wrap_module_list(self) -> List[nn.Module]:  # model_new.py:686-687 (API surface only)
        return (list(self.backbone.transformer.h) + list(self.audio_understanding_expert.transformer.h)
                + list(self.audio_generation_expert.transformer.h))


def _from_ptr(addr: int, shape, device) -> torch.Tensor:
    """Zero-copy fp32 torch view of library-owned device memory (via __cuda_array_interface__)."""
    n = 1
    for s in shape:
        n *= s

    class _Holder:
        pass

    hld = _Holder()
    hld.__cuda_array_interface__ = dict(shape=(n,), typestr="<f4", data=(addr, False), version=3, strides=None)
    return torch.as_tensor(hld, device=device).view(*shape)
