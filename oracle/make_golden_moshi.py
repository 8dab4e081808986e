"""Generate tests/golden/moshi_golden.pt from the UNMODIFIED reference Moshi-family modules (llm_modules/transformer.py,
gating.py, rope.py via the alias import of oracle/ref_shims.py, and llm_utils/sampling.py) and assert that
oracle/moshi_oracle.py is bit-identical to them on CPU.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_moshi
"""
import dataclasses
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import moshi_oracle as MO  # noqa: E402
from oracle.ref_shims import install_moshi_shims  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "moshi_golden.pt")


def stx_cfgs():
    return {
        # the codec's transformer flavour (MimiCodec.py:54-58) in streaming mode: LayerNorm, GELU FF, LayerScale, RoPE, hs 32
        "mimi_like": MO.StxCfg(d_model=128, num_heads=4, num_layers=2, dim_feedforward=256, context=12, positional_embedding="rope",
                               norm="layer_norm", layer_scale=0.01, gating="none"),
        # the Moshi LM flavour: RMSNorm (fp32), SiLU gating with hidden = 21 d / 8, RoPE, hs 64
        "lm_like": MO.StxCfg(d_model=256, num_heads=4, num_layers=2, dim_feedforward=1024, context=16, positional_embedding="rope",
                             norm="rms_norm_f32", layer_scale=None, gating="silu"),
        # the depformer flavour: one weight slab per step, capacity = weights_per_step, no positional embedding
        "dep_like": MO.StxCfg(d_model=128, num_heads=4, num_layers=2, dim_feedforward=[192, 384, 192, 384], context=None,
                              positional_embedding="none", norm="rms_norm", layer_scale=None, gating="silu", weights_per_step=4),
        # sinusoidal + rotary, LayerNormF32, hs 128, window shorter than the ring traffic
        "sin_like": MO.StxCfg(d_model=256, num_heads=2, num_layers=1, dim_feedforward=512, context=8, positional_embedding="sin_rope",
                              norm="layer_norm_f32", layer_scale=None, gating="none", positional_scale=0.5),
    }


# streaming schedules: chunk lengths T fed one forward at a time (they cross the ring wrap of every config)
SCHEDULES = {
    "mimi_like": [1, 1, 2, 3, 1, 2, 1, 1, 4, 1, 1, 2],  # 20 steps through a 12-slot ring
    "lm_like": [1] * 20,
    "dep_like": [1, 1, 1, 1],
    "sin_like": [2, 2, 1, 3, 2, 1, 1],
}
BATCH = {"mimi_like": 2, "lm_like": 2, "dep_like": 3, "sin_like": 1}
NONSTREAM_T = {"mimi_like": 17, "lm_like": 9, "dep_like": 4, "sin_like": 11}


def build_reference(tr, cfg: MO.StxCfg, sd):
    act = {"none": F.gelu}.get(cfg.gating, F.gelu)
    m = tr.StreamingTransformer(d_model=cfg.d_model, num_heads=cfg.num_heads, num_layers=cfg.num_layers,
                                dim_feedforward=cfg.dim_feedforward, causal=cfg.causal, context=cfg.context,
                                positional_embedding=cfg.positional_embedding, max_period=cfg.max_period,
                                positional_scale=cfg.positional_scale, norm=cfg.norm, layer_scale=cfg.layer_scale,
                                gating=cfg.gating, weights_per_step=cfg.weights_per_step, activation=act)
    full = m.state_dict()
    assert set(full.keys()) == set(sd.keys()), (sorted(set(full) ^ set(sd)))
    for k, v in sd.items():
        assert full[k].shape == v.shape, (k, full[k].shape, v.shape)
    m.load_state_dict(sd, strict=True)
    return m.float().eval()


def sampler_cases():
    # (name, logits shape, kwargs)
    return [
        ("greedy", (2, 3, 64), dict(use_sampling=False)),
        ("plain", (2, 3, 64), dict(use_sampling=True, temp=0.8)),
        ("topk5", (2, 3, 64), dict(use_sampling=True, temp=0.7, top_k=5)),
        ("topk25_big", (1, 1, 3000), dict(use_sampling=True, temp=1.0, top_k=25)),
        ("topk250", (3, 8, 2048), dict(use_sampling=True, temp=0.8, top_k=250)),
        ("topp", (2, 2, 500), dict(use_sampling=True, temp=0.9, top_p=0.8)),
        ("temp0_is_greedy", (2, 3, 64), dict(use_sampling=True, temp=0.0, top_k=5)),
    ]


def main():
    torch.set_num_threads(4)
    tr, samp = install_moshi_shims()
    out = {}
    with torch.no_grad():
        for name, cfg in stx_cfgs().items():
            sd = MO.random_state_dict(cfg, seed=2025)
            ref = build_reference(tr, cfg, sd)
            orc = MO.StxOracle(cfg, sd)
            g = torch.Generator().manual_seed(5)
            B = BATCH[name]
            # ---- non-streaming forward
            xn = torch.randn(B, NONSTREAM_T[name], cfg.d_model, generator=g)
            yn_ref = ref(xn)
            yn = orc.forward(xn)
            assert torch.equal(yn_ref, yn), f"{name}: non-streaming oracle != reference ({(yn_ref - yn).abs().max()})"
            # ---- streaming, two passes separated by reset_streaming()
            xs, ys = [], []
            with ref.streaming(B):
                orc.start_streaming(B)
                for rep in range(2):
                    for T in SCHEDULES[name]:
                        x = torch.randn(B, T, cfg.d_model, generator=g)
                        y_ref = ref(x)
                        y = orc.forward(x)
                        assert torch.equal(y_ref, y), f"{name}: streaming oracle != reference ({(y_ref - y).abs().max()})"
                        xs.append(x)
                        ys.append(y_ref)
                    kv_ref = ref.layers[-1].self_attn._streaming_state.kv_cache
                    kv_o = orc.state["kv"][-1]
                    assert torch.equal(kv_ref.cache, kv_o.cache) and int(kv_ref.end_offset) == kv_o.end_offset
                    last_cache = kv_ref.cache.clone()
                    last_end = int(kv_ref.end_offset)
                    if rep == 0:
                        ref.reset_streaming()
                        orc.reset_streaming()
                orc.stop_streaming()
            # ring position recovery, stated on its own
            for E in (0, 1, 5, cfg.capacity(), cfg.capacity() + 1, 3 * cfg.capacity() + 2):
                rk = tr.RingKVCache(1, 1, 2, cfg.capacity(), device=torch.device("cpu"), dtype=torch.float32)
                rk.end_offset += E
                pos = rk.complete(torch.zeros(1, 1, 0, 2), torch.zeros(1, 1, 0, 2)).positions
                assert torch.equal(pos, MO.ring_positions(cfg.capacity(), E)), (name, E)
            out[name] = dict(cfg=dataclasses.asdict(cfg), x_nonstream=xn, y_nonstream=yn_ref, xs=xs, ys=ys,
                             schedule=SCHEDULES[name], batch=B, last_cache=last_cache, last_end=last_end)
            out[f"__checksum_{name}"] = {k: float(v.double().sum()) for k, v in sd.items()}
            print(f"[ok] {name}: non-streaming T={NONSTREAM_T[name]} and {2 * len(SCHEDULES[name])} streaming calls bit-exact")

        # ---- sampler
        for cname, shape, kw in sampler_cases():
            g = torch.Generator().manual_seed(len(cname) * 7 + shape[-1])
            logits = torch.randn(*shape, generator=g) * 3.0
            rows = logits[..., 0].numel()
            n_noise = kw["top_k"] if (kw.get("top_k", 0) > 0 and not kw.get("top_p", 0.0) > 0.0) else shape[-1]
            torch.manual_seed(99)
            tok_ref = samp.sample_token(logits, **kw)
            torch.manual_seed(99)
            q = torch.empty(rows, n_noise).exponential_(1)
            tok = MO.sample_token(logits, q=q, **kw)
            assert torch.equal(tok_ref, tok), f"sampler {cname}: oracle != reference"
            out[f"sampler_{cname}"] = dict(logits=logits, kwargs=kw, q=q, tokens=tok_ref)
            print(f"[ok] sampler {cname}: {tuple(tok_ref.shape)} ids equal")
        # sample_token_audio (4-D logits, ids >= end_token excluded after the softmax)
        g = torch.Generator().manual_seed(3)
        logits = torch.randn(2, 1, 3, 64, generator=g) * 3.0
        for cname, kw in (("audio_topk", dict(use_sampling=True, temp=0.8, top_k=5)), ("audio_plain", dict(use_sampling=True, temp=1.1))):
            torch.manual_seed(7)
            tok_ref = samp.sample_token_audio(logits, end_token=40, **kw)
            torch.manual_seed(7)
            q = torch.empty(6, kw.get("top_k", 0) or 64).exponential_(1)
            tok = MO.sample_token(logits, q=q, end_token=40, **kw)
            assert torch.equal(tok_ref, tok), f"sampler {cname}: oracle != reference"
            assert int(tok_ref.max()) < 40
            out[f"sampler_{cname}"] = dict(logits=logits, kwargs=dict(kw, end_token=40), q=q, tokens=tok_ref)
            print(f"[ok] sampler {cname}: ids equal, all < end_token")
    torch.save(out, GOLDEN)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN) / 1e6, "MB")


if __name__ == "__main__":
    main()
