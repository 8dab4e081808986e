"""TEST INFRASTRUCTURE ONLY.  Writes tests/golden/thinking_golden.pt by running the UNMODIFIED reference code of the reasoning encoder:
modules/transformer.py is imported as it stands (soft_moe_pytorch, which this configuration never touches, is stubbed), the five
TransformerBlocks are built exactly like AudioThinking.__init__ builds them (AudioDiffusion1D.py:175-181), and the method source of
encode_reasoning_part / set_masking / extract_mask_positions (:372-390, :458-486) is executed on a stand-in self.  Asserts that
oracle/thinking_oracle.py reproduces the query tokens bit for bit.

    python -m oracle.make_golden_thinking
"""
import importlib.util
import os
import sys
import types
import warnings

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import thinking_oracle as TO  # noqa: E402
from oracle.make_golden_film import load_methods  # noqa: E402
from oracle.ref_shims import REF_ROOT  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "thinking_golden.pt")
SMALL = dict(dim=256, dim_heads=128, depth=2, interval=5, whisper_dim=64, mu_dim=48, ff_mult=4)


def reference_transformer_module():
    sys.modules.setdefault("soft_moe_pytorch", types.SimpleNamespace(SoftMoE=object))
    path = os.path.join(REF_ROOT, "tools", "tokenizer", "ReasoningCodec_film", "modules", "transformer.py")
    spec = importlib.util.spec_from_file_location("ref_thinking_transformer", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def build_reference(cfg, sd, mod):
    """The members of AudioThinking that encode_reasoning_part touches, built with the reference's own constructor arguments."""
    at = types.SimpleNamespace(interval=cfg["interval"], cls_token=nn.Parameter(sd["cls_token"].clone()))
    blocks = [mod.TransformerBlock(cfg["dim"], dim_heads=cfg["dim_heads"], causal=False, zero_init_branch_outputs=False, remove_norms=False,
                                   power_normalized=True, conformer=False, layer_scale=True, add_rope=True, attn_kwargs={"qk_norm": True},
                                   ff_kwargs={"mult": cfg["ff_mult"], "no_bias": False}, norm_kwargs={"eps": 1e-2}) for _ in range(cfg["depth"])]
    at.encoder_transformers = nn.Sequential(*blocks)
    own = {k[len("encoder_transformers."):]: v for k, v in sd.items() if k.startswith("encoder_transformers.")}
    missing, unexpected = at.encoder_transformers.load_state_dict(own, strict=True), None
    at.semantic_merge_proj = nn.Linear(cfg["whisper_dim"] + cfg["mu_dim"], cfg["dim"])
    at.semantic_merge_proj.load_state_dict({"weight": sd["semantic_merge_proj.weight"], "bias": sd["semantic_merge_proj.bias"]})
    at.down_sampling_layer_whisper = nn.Conv1d(cfg["whisper_dim"], cfg["whisper_dim"], 2, stride=2)
    at.down_sampling_layer_whisper.load_state_dict({"weight": sd["down_sampling_layer_whisper.weight"], "bias": sd["down_sampling_layer_whisper.bias"]})
    at.reasoning_vq = lambda q: (q, None, None)  # the third-party ResidualVQ is a18's: here the query tokens themselves are the fixture
    at.encoder_transformers.eval()
    return at


def main():
    warnings.filterwarnings("ignore")
    torch.set_num_threads(4)
    mod = reference_transformer_module()
    m = load_methods(["encode_reasoning_part", "set_masking", "extract_mask_positions"])
    out = {"cases": {}}
    for name, cfg, seed, (B, Tw, Tb) in (("small", SMALL, 21, (2, 60, 30)), ("ragged", SMALL, 22, (1, 44, 20))):
        sd = TO.random_state_dict(cfg, seed)
        self_ = types.SimpleNamespace(audio_thinking=build_reference(cfg, sd, mod))
        self_.set_masking = lambda x: m["set_masking"](self_, x)
        self_.extract_mask_positions = lambda x: m["extract_mask_positions"](self_, x)
        g = torch.Generator().manual_seed(seed + 100)
        whisper, mu = torch.randn(B, cfg["whisper_dim"], Tw, generator=g), torch.randn(B, cfg["mu_dim"], Tb, generator=g)
        with torch.no_grad():
            ref, _, _ = m["encode_reasoning_part"](self_, whisper, mu)
            mine = TO.encode(sd, cfg, whisper, mu)
        assert ref.shape == (B, min(Tw // 2, Tb) // cfg["interval"], cfg["dim"]), ref.shape
        assert torch.equal(ref, mine), (name, float((ref - mine).abs().max()))
        out["cases"][name] = {"cfg": cfg, "seed": seed, "whisper": whisper, "mu": mu, "query_tokens": ref}
        print(f"[ok] {name}: encode_reasoning_part source == restatement bit-exact, query tokens {tuple(ref.shape)}, scale {float(ref.abs().max()):.3f}")
    torch.save(out, GOLDEN)
    print(GOLDEN, os.path.getsize(GOLDEN) // 1024, "KiB")


if __name__ == "__main__":
    main()
