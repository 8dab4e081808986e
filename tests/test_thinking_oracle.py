"""CPU: oracle/thinking_oracle.py (the reasoning encoder `AudioThinking` up to its query tokens) against fixtures written by
oracle/make_golden_thinking.py from the UNMODIFIED reference code (modules/transformer.py imported as it stands; the method source of
encode_reasoning_part / set_masking / extract_mask_positions executed on a stand-in self), and the whole ua2_thinking_* handle of
csrc/ua2_thinking.cu as shipped on the CPU shim (tensor-core GEMM swapped for a CPU GEMM) against the oracle."""
import ctypes as C
import os

import pytest
import torch

from oracle import thinking_oracle as TO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(ROOT, "tests", "golden", "thinking_golden.pt"), weights_only=False)


@pytest.mark.parametrize("name", ["small", "ragged"])
def test_thinking_oracle_matches_reference_fixture(gold, name):
    """The generator asserts bit-equality with the reference source inside its own process; across processes ATen's matmul blocking (and
    so the last bits of a 2-block chain of 768-wide sums) depends on the thread count, hence a 2e-6 bar here."""
    c = gold["cases"][name]
    sd = TO.random_state_dict(c["cfg"], c["seed"])
    with torch.no_grad():
        q = TO.encode(sd, c["cfg"], c["whisper"], c["mu"])
    ref = c["query_tokens"]
    assert q.shape == ref.shape
    assert float((q - ref).abs().max()) < 2e-6 * max(1.0, float(ref.abs().max()))


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def fold_weight_norm(sd):
    """The parameters as ua2_thinking_load_weight takes them: weight-normed linears folded to their effective weights."""
    out = {}
    for k, v in sd.items():
        if k.endswith(".parametrizations.weight.original0"):
            pre = k[:-len(".parametrizations.weight.original0")]
            out[pre + ".weight"] = TO.wn(sd, pre).contiguous()
        elif not k.endswith(".parametrizations.weight.original1"):
            out[k] = v.contiguous()
    return out


def test_whole_thinking_handle_on_cpu_against_reference_golden(encoder_handles_shim, gold, name="small"):
    """Parameter loading, the down-sampling convolution as a GEMM, concatenation, merge projection into the set_masking layout, query
    token rows, both blocks (per-head LayerNorm + rotary embedding, attention at head size 128, LayerScale residuals, sigmoid GLU) as
    shipped, against the query tokens the UNMODIFIED reference code produced.  (The `ragged` fixture - 20 rows, below the tensor-core
    path's 32-row threshold - runs on the GPU only: the skinny linear kernels take a minute of OS-thread emulation.)"""
    from uniaudio2_b200 import _lib

    lib = encoder_handles_shim
    c = gold["cases"][name]
    cfg = c["cfg"]
    sd = fold_weight_norm(TO.random_state_dict(cfg, c["seed"]))
    ccfg = _lib.ThinkingCfg(cfg["dim"], cfg["dim_heads"], cfg["depth"], cfg["interval"], cfg["whisper_dim"], cfg["mu_dim"], cfg["ff_mult"])
    h = C.c_void_p()
    assert lib.ua2_thinking_create(C.byref(ccfg), C.byref(h)) == 0, lib.ua2_last_error()
    for key, t in sd.items():
        shape = (C.c_int64 * t.dim())(*t.shape)
        assert lib.ua2_thinking_load_weight(h, key.encode(), _p(t), shape, t.dim()) == 0, (key, lib.ua2_last_error())
    assert lib.ua2_thinking_finalize(h, None) == 0, lib.ua2_last_error()
    B, _, Tw = c["whisper"].shape
    Tb = c["mu"].shape[-1]
    rows = int(lib.ua2_thinking_rows(h, Tw, Tb))
    assert rows == min(Tw // 2, Tb) * (cfg["interval"] + 1) // cfg["interval"]
    out = torch.full((B, rows, cfg["dim"]), float("nan"))
    assert lib.ua2_thinking_encode(h, _p(c["whisper"].contiguous()), _p(c["mu"].contiguous()), B, Tw, Tb, _p(out), None) == 0, lib.ua2_last_error()
    q = out[:, cfg["interval"]::cfg["interval"] + 1]
    ref = c["query_tokens"]
    assert q.shape == ref.shape
    assert float((q - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max())), float((q - ref).abs().max())
    assert lib.ua2_thinking_encode(h, _p(c["whisper"].contiguous()), _p(c["mu"].contiguous()), B, Tw - 2, Tb, _p(out), None) != 0 or (min((Tw - 2) // 2, Tb) % cfg["interval"] == 0)
    assert lib.ua2_thinking_destroy(h) == 0
