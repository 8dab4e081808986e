"""Generate tests/golden/llm_*.pt by running the UNMODIFIED reference (imported from /root/reference
through oracle/ref_shims) on seeded tiny configs, and assert that oracle/llm_oracle.py is bit-identical
to it on CPU.  TEST INFRASTRUCTURE ONLY.  Run in the build container:

    python -m oracle.make_golden

The fixtures hold inputs, the reference's outputs (token ids, logits, hidden states) and a checksum
of the seeded weights (weights are regenerated from the seed by oracle.llm_oracle.random_state_dict).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import llm_oracle as O  # noqa: E402
from oracle.ref_shims import install_llm_shims  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


from oracle.cases import CASES, REASON_CARD, run_case, sd_checksum, tiny_cfgs  # noqa: E402


def _ref_cfg_dict(c: O.GPTCfg, name):
    return dict(
        name=name, hf_config=dict(org="meta-llama", name=name), block_size=c.block_size, vocab_size=c.padded_vocab_size,
        padded_vocab_size=c.padded_vocab_size, n_layer=c.n_layer, n_embd=c.n_embd, n_head=c.n_head,
        n_query_groups=c.n_query_groups, rotary_percentage=1.0, parallel_residual=False, bias=False,
        norm_class_name="RMSNorm", mlp_class_name="LLaMAMLP", intermediate_size=c.intermediate_size,
        rope_base=c.rope_base, rope_adjustments=dict(c.rope_adjustments),
    )


def build_reference(model_new, cfg: O.Stage3Cfg, sd):
    import llm_models.config as rc

    rc.name_to_config["ua2-backbone"] = _ref_cfg_dict(cfg.backbone, "ua2-backbone")
    rc.name_to_config["ua2-decoder"] = _ref_cfg_dict(cfg.decoder, "ua2-decoder")
    rc.name_to_config["meta-llama/Llama-3.2-Understanding"] = _ref_cfg_dict(cfg.understanding, "Llama-3.2-Understanding")
    rc.name_to_config["meta-llama/Llama-3.2-Generation"] = _ref_cfg_dict(cfg.generation, "Llama-3.2-Generation")
    # the reference hard-codes max_seq_length=2048 in setup_caches; shrink block_size so rope tables stay small
    for k in ("ua2-backbone", "ua2-decoder", "meta-llama/Llama-3.2-Understanding", "meta-llama/Llama-3.2-Generation"):
        rc.name_to_config[k]["block_size"] = 2048
    args = model_new.ModelArgs(
        llm_name="ua2-backbone", decoder_name="ua2-decoder", llm_pretrained_model="", audio_embeddings_path="",
        audio_understanding_expert_path="", audio_semantic_vocab_size=cfg.audio_vocab - 7, audio_reason_vocab_size=7,
        audio_num_codebooks=cfg.num_codebooks,
    )
    m = model_new.Model_stage3(args)
    missing = m.load_state_dict(sd, strict=True)
    m.eval()
    return m


def main():
    torch.set_num_threads(8)
    model_new = install_llm_shims()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    cfgs = tiny_cfgs()
    fixtures = {}
    for cname, cfg in cfgs.items():
        sd = O.random_state_dict(cfg, seed=1234)
        ref = build_reference(model_new, cfg, sd)
        # the reference hard-codes 2048 cache slots (model_new.py:560-565); the oracle/product may use fewer
        ref.setup_caches(3)
        orc = O.Stage3Oracle(cfg, sd)
        orc.setup_caches(3)
        for (name, cn, kind, B, S, nf, topk, temp, cfgs_) in CASES:
            if cn != cname:
                continue
            r = run_case(ref, kind, cfg, B, S, nf, topk, temp, cfgs_, REASON_CARD[cname], 42, True)
            o = run_case(orc, kind, cfg, B, S, nf, topk, temp, cfgs_, REASON_CARD[cname], 42, False)
            # the explicit-noise route (used to share draws with the device under test) must consume the same stream
            o2 = run_case(orc, kind, cfg, B, S, nf, topk, temp, cfgs_, REASON_CARD[cname], 42, False, explicit_noise=True)
            assert torch.equal(r["frames"], o2["frames"]), f"{name}: explicit-noise oracle run != reference"
            assert torch.equal(r["frames"], o["frames"]), f"{name}: oracle tokens != reference tokens"
            # bit-exact hidden state check through the KV caches of the last backbone layer
            rk = ref.backbone.transformer.h[-1].attn.kv_cache.k[:B, :, : S + nf]
            ok = orc.backbone.kv[-1].k[:B, :, : S + nf]
            bb_k = ok.clone()
            assert torch.equal(rk, ok), f"{name}: oracle KV cache != reference KV cache (max diff {(rk-ok).abs().max()})"
            rk = ref.audio_generation_expert.transformer.h[-1].attn.kv_cache.v[:B, :, : S + nf]
            ok = orc.gen.kv[-1].v[:B, :, : S + nf]
            assert torch.equal(rk, ok), f"{name}: gen-expert V cache mismatch"
            print(f"[ok] {name}: oracle == reference bit-exact; frames=\n{r['frames'][:, 0].tolist()}")
            # margins (top1 - top2 of every sampled head) so GPU tests know how much slack greedy ids have
            tl = o["text_logits"]
            top2 = tl.topk(2, dim=-1)[0]
            margin_text = float((top2[..., 0] - top2[..., 1]).min())
            cl = o["ci_logits"]
            top2 = cl.topk(2, dim=-1)[0]
            margin_audio = float((top2[..., 0] - top2[..., 1]).min())
            fixtures[name] = dict(
                cfg_name=cname, kind=kind, B=B, S=S, n_frames=nf, topk=topk, temperature=temp, cfg_scale=cfgs_,
                reason_card=REASON_CARD[cname], prompt_tokens=r["prompt_tokens"], prompt_mask=r["prompt_mask"],
                ref_frames=r["frames"], text_logits=o["text_logits"], ci_logits=o["ci_logits"], h_final=o["h_final"],
                last_backbone_k=bb_k, last_gen_v=ok.clone(), margin_text=margin_text, margin_audio=margin_audio,
            )
            print(f"     margins: text {margin_text:.3e} audio {margin_audio:.3e}")
        fixtures[f"__checksum_{cname}"] = sd_checksum(sd)
    torch.save(fixtures, os.path.join(GOLDEN_DIR, "llm_golden.pt"))
    print("wrote", os.path.join(GOLDEN_DIR, "llm_golden.pt"), os.path.getsize(os.path.join(GOLDEN_DIR, "llm_golden.pt")) / 1e6, "MB")


if __name__ == "__main__":
    main()
