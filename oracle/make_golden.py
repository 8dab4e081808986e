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


def tiny_cfgs():
    """Two shrunken Stage3 configs: 'tiny' (hs 64/64, q_per_kv 3/4) and 'mid' (hs 128/64 like full size)."""
    out = {}
    out["tiny"] = O.Stage3Cfg(
        backbone=O.GPTCfg(n_layer=2, n_embd=384, n_head=6, n_query_groups=2, intermediate_size=512, padded_vocab_size=1024),
        decoder=O.GPTCfg(n_layer=2, n_embd=256, n_head=4, n_query_groups=1, intermediate_size=384, padded_vocab_size=1024),
        understanding=O.GPTCfg(n_layer=2, n_embd=384, n_head=6, n_query_groups=2, intermediate_size=512, padded_vocab_size=1024),
        generation=O.GPTCfg(n_layer=1, n_embd=384, n_head=6, n_query_groups=2, intermediate_size=512, padded_vocab_size=1024),
        audio_vocab=40 + 90, num_codebooks=8, max_seq_length=64,
    )
    out["mid"] = O.Stage3Cfg(
        backbone=O.GPTCfg(n_layer=3, n_embd=768, n_head=6, n_query_groups=2, intermediate_size=1280, padded_vocab_size=2048),
        decoder=O.GPTCfg(n_layer=2, n_embd=512, n_head=8, n_query_groups=2, intermediate_size=768, padded_vocab_size=2048),
        understanding=O.GPTCfg(n_layer=1, n_embd=768, n_head=6, n_query_groups=2, intermediate_size=1280, padded_vocab_size=2048),
        generation=O.GPTCfg(n_layer=1, n_embd=768, n_head=6, n_query_groups=2, intermediate_size=1280, padded_vocab_size=2048),
        audio_vocab=300 + 500, num_codebooks=8, max_seq_length=96,
    )
    return out


REASON_CARD = {"tiny": 40, "mid": 300}


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


def make_prompt(kind, cfg: O.Stage3Cfg, B, S, gen: torch.Generator, reason_card):
    """(B,S,9) tokens / masks.  'text': text-only prompt (TTS, tts_task.py:192-206).
    'mixed': text prompt then audio frames (ASR/caption, asr_task.py:299-326)."""
    nq = cfg.num_codebooks
    V_t = cfg.backbone.padded_vocab_size
    tokens = torch.zeros(B, S, nq + 1, dtype=torch.long)
    mask = torch.zeros(B, S, nq + 1, dtype=torch.bool)
    n_text = S if kind == "text" else S // 3
    tokens[:, :n_text, -1] = torch.randint(0, V_t, (B, n_text), generator=gen)
    mask[:, :n_text, -1] = True
    if n_text < S:
        tokens[:, n_text:, :-1] = torch.randint(0, cfg.audio_vocab, (B, S - n_text, nq), generator=gen)
        mask[:, n_text:, :-1] = True
    return tokens, mask


def run_case(model, kind, cfg, B, S, n_frames, topk, temperature, cfg_scale, reason_card, seed, is_ref):
    """Drive prefill + n_frames of generate_frame the way tts_task.py:208-285 / asr_task.py:630-688 do."""
    gen = torch.Generator().manual_seed(seed)
    tokens, mask = make_prompt(kind, cfg, B, S, gen, reason_card)
    if B == 2 and cfg_scale > 1.0:
        tokens[1] = tokens[0]
        tokens[1, :, -1] = 7  # text_pad-like negative prompt (tts_task.py:171-190)
        mask[1] = mask[0]
    model.reset_caches()
    pos = torch.arange(0, S).unsqueeze(0).repeat(B, 1)
    with torch.inference_mode():
        if is_ref:
            model.forward_prefix(tokens[:, :-1], labels=tokens[:, 1:, :-1], tokens_mask=mask, loss_mask=mask, input_pos=pos[:, :-1])
        else:
            model.forward_prefix(tokens[:, :-1], mask, pos[:, :-1])
        curr_tokens, curr_mask = tokens[:, -1:], mask[:, -1:]
        curr_pos = torch.tensor([S - 1], dtype=torch.long)
        maxp1 = S
        torch.manual_seed(888)  # multi_task_inference.py:596 sampler seed
        frames, text_logits, ci_logits, h_final = [], [], [], []
        for f in range(n_frames):
            forbid = 0 if f < n_frames // 2 else reason_card
            if is_ref:
                # capture logits with a hook-free trick: re-run heads is not possible without
                # touching caches, so the reference run only records samples.
                s = model.generate_frame(curr_tokens, curr_mask, input_pos=curr_pos, input_pos_maxp1=maxp1,
                                         temperature=temperature, topk=topk, forbid_prefix=forbid, cfg_scale=cfg_scale)
            else:
                dbg = {}
                s = model.generate_frame(curr_tokens, curr_mask, curr_pos, maxp1, temperature, topk, forbid, cfg_scale, debug=dbg)
                text_logits.append(dbg["text_logits"])
                ci_logits.append(torch.stack(dbg["ci_logits"]))
                h_final.append(dbg["h_final"])
            frames.append(s.clone())
            if kind == "text" or True:
                # feed back like tts_task.py:276-279: audio tokens in cols 0-7, text token in col 8, audio-step mask
                audio = s[:, 1:].long()
                text = s[:, 0:1].long()
                if kind == "asr_decode":
                    curr_tokens = torch.cat([torch.zeros_like(audio), text], dim=-1).unsqueeze(1)
                    curr_mask = torch.cat([torch.zeros_like(audio).bool(), torch.ones(B, 1).bool()], dim=1).unsqueeze(1)
                else:
                    curr_tokens = torch.cat([audio, text], dim=-1).unsqueeze(1)
                    curr_mask = torch.cat([torch.ones_like(audio).bool(), torch.zeros(B, 1).bool()], dim=1).unsqueeze(1)
            curr_pos = curr_pos + 1
            maxp1 += 1
    out = dict(prompt_tokens=tokens, prompt_mask=mask, frames=torch.stack(frames))
    if not is_ref:
        out.update(text_logits=torch.stack(text_logits), ci_logits=torch.stack(ci_logits), h_final=torch.stack(h_final))
    return out


CASES = [
    # name, cfg, kind, B, S, frames, topk, temp, cfg_scale
    ("tiny_tts_greedy", "tiny", "text", 1, 12, 8, 1, 1.0, 1.0),
    ("tiny_tts_topk", "tiny", "text", 1, 9, 6, 5, 0.9, 1.0),
    ("tiny_mixed_greedy_b2", "tiny", "mixed", 2, 15, 6, 1, 1.0, 1.0),
    ("tiny_asr_decode", "tiny", "asr_decode", 1, 14, 5, 1, 1.0, 1.0),
    ("tiny_cfg", "tiny", "text", 2, 10, 4, 3, 0.8, 1.5),
    ("mid_tts_greedy", "mid", "text", 1, 20, 10, 1, 1.0, 1.0),
    ("mid_mixed_topk_b3", "mid", "mixed", 3, 33, 6, 20, 0.9, 1.0),
]


def sd_checksum(sd):
    return {k: float(v.double().sum()) for k, v in sd.items()}


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
            kind_prompt = "mixed" if kind == "asr_decode" else kind
            r = run_case(ref, kind if kind == "asr_decode" else kind_prompt, cfg, B, S, nf, topk, temp, cfgs_, REASON_CARD[cname], 42, True)
            o = run_case(orc, kind if kind == "asr_decode" else kind_prompt, cfg, B, S, nf, topk, temp, cfgs_, REASON_CARD[cname], 42, False)
            assert torch.equal(r["frames"], o["frames"]), f"{name}: oracle tokens != reference tokens"
            # bit-exact hidden state check through the KV caches of the last backbone layer
            rk = ref.backbone.transformer.h[-1].attn.kv_cache.k[:B, :, : S + nf]
            ok = orc.backbone.kv[-1].k[:B, :, : S + nf]
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
                last_backbone_k=ok.clone(), margin_text=margin_text, margin_audio=margin_audio,
            )
            print(f"     margins: text {margin_text:.3e} audio {margin_audio:.3e}")
        fixtures[f"__checksum_{cname}"] = sd_checksum(sd)
    torch.save(fixtures, os.path.join(GOLDEN_DIR, "llm_golden.pt"))
    print("wrote", os.path.join(GOLDEN_DIR, "llm_golden.pt"), os.path.getsize(os.path.join(GOLDEN_DIR, "llm_golden.pt")) / 1e6, "MB")


if __name__ == "__main__":
    main()
