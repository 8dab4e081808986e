"""Generate tests/golden/whisper_golden.pt by executing the UNMODIFIED source of WhisperPositionalEmbedding / WhisperAttention /
WhisperEncoderLayer / WhisperEncoder (tools/tokenizer/ReasoningCodec_film/models/modeling_whisper.py:212-443, :723-867) and assert that
oracle/whisper_oracle.py is bit-identical.  The module itself cannot be imported here (it is a fork of a transformers 4.2x file and
imports names this image's transformers 5.5 no longer has), so the four class bodies are compiled as they stand into a namespace that
supplies torch, `ACT2FN["gelu"] = F.gelu` (transformers.activations.GELUActivation is F.gelu), a WhisperPreTrainedModel that is a
plain nn.Module with a no-op post_init, and a BaseModelOutput record.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_whisper
"""
import ast
import math
import os
import random
import sys
import types
from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import whisper_oracle as WO  # noqa: E402
from oracle.ref_shims import REF_ROOT  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "whisper_golden.pt")
THREADS = 4
CASES = [  # (name, cfg, batch, param seed, input seed)
    ("tiny_h2", WO.WhisperCfg(d_model=128, encoder_attention_heads=2, encoder_ffn_dim=256, encoder_layers=2, max_source_positions=150,
                              num_mel_bins=80), 2, 11, 12),
    ("tiny_h4_ragged", WO.WhisperCfg(d_model=256, encoder_attention_heads=4, encoder_ffn_dim=512, encoder_layers=3, max_source_positions=67,
                                     num_mel_bins=80), 1, 21, 22),
]


def load_classes():
    path = os.path.join(REF_ROOT, "tools", "tokenizer", "ReasoningCodec_film", "models", "modeling_whisper.py")
    src = open(path).read()

    class WhisperPreTrainedModel(nn.Module):
        def __init__(self, config):
            super().__init__()
            self.config = config

        def post_init(self):
            pass

    class BaseModelOutput:
        def __init__(self, last_hidden_state=None, hidden_states=None, attentions=None):
            self.last_hidden_state, self.hidden_states, self.attentions = last_hidden_state, hidden_states, attentions

    ns = {"torch": torch, "nn": nn, "math": math, "random": random, "np": np, "Optional": Optional, "Tuple": Tuple,
          "ACT2FN": {"gelu": F.gelu}, "WhisperPreTrainedModel": WhisperPreTrainedModel, "BaseModelOutput": BaseModelOutput,
          "WhisperConfig": object}
    want = ["WhisperPositionalEmbedding", "WhisperAttention", "WhisperEncoderLayer", "WhisperEncoder"]
    for node in ast.parse(src).body:
        if isinstance(node, ast.ClassDef) and node.name in want:
            text = "\n".join(src.splitlines()[node.lineno - 1:node.end_lineno])
            exec(compile(text, path, "exec"), ns)
    assert all(w in ns for w in want)
    return ns


def hf_config(cfg: WO.WhisperCfg):
    return types.SimpleNamespace(d_model=cfg.d_model, encoder_attention_heads=cfg.encoder_attention_heads, encoder_ffn_dim=cfg.encoder_ffn_dim,
                                 encoder_layers=cfg.encoder_layers, max_source_positions=cfg.max_source_positions,
                                 num_mel_bins=cfg.num_mel_bins, dropout=0.0, attention_dropout=0.0, activation_dropout=0.0,
                                 activation_function="gelu", encoder_layerdrop=0.0, pad_token_id=0, scale_embedding=False,
                                 output_attentions=False, output_hidden_states=False, use_return_dict=True)


def main():
    torch.set_num_threads(THREADS)
    ns = load_classes()
    out = {"threads": THREADS, "cases": {}}
    for name, cfg, B, pseed, iseed in CASES:
        enc = ns["WhisperEncoder"](hf_config(cfg)).eval()
        sd = WO.random_state_dict(cfg, pseed)
        assert sorted(enc.state_dict().keys()) == sorted(WO.state_keys(cfg)), "state-dict keys of the restatement differ from the reference module"
        enc.load_state_dict(sd, strict=True)
        g = torch.Generator().manual_seed(iseed)
        mel = torch.randn(B, cfg.num_mel_bins, 2 * cfg.max_source_positions, generator=g)
        with torch.no_grad():
            ref = enc(mel, return_dict=True).last_hidden_state
            got = WO.WhisperEncoderOracle(cfg, sd).forward(mel)
        assert torch.equal(ref, got), f"{name}: restatement differs from the reference source (max {float((ref - got).abs().max())})"
        out["cases"][name] = dict(cfg=cfg.__dict__.copy(), batch=B, param_seed=pseed, input_seed=iseed, out=ref)
        print(f"[ok] {name}: {tuple(ref.shape)} bit-exact, |out| max {float(ref.abs().max()):.3f}")
    torch.save(out, GOLDEN)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN) / 1e3, "KB")


if __name__ == "__main__":
    main()
