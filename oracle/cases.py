"""Shared seeded test cases for the AR-decode path (configs, prompts, the prefill + frame loop driver).
TEST INFRASTRUCTURE ONLY - imported by oracle/make_golden.py, tests/ and bench.py's CPU-baseline leg.

`run_case` drives any object exposing the reference's Model_stage3 API (the unmodified reference, the product's
uniaudio2_b200 Model_stage3) or the oracle (is_ref=False) the way evaluation/tts_task.py:208-285 and
evaluation/asr_task.py:630-688 do.
"""
import torch

from oracle import llm_oracle as O


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


def cpu_noise_for_frame(rows, V_t, V_a, nq):
    """The Exp(1) draws one generate_frame call consumes from torch's CPU generator, in the reference's order
    (model_new.py:141-143 via :623 then :639 x nq), flattened to the layout ua2_llm_generate_frame expects."""
    parts = [torch.empty(rows, V_t).exponential_(1)]
    for _ in range(nq):
        parts.append(torch.empty(rows, V_a).exponential_(1))
    return parts


def run_case(model, kind, cfg, B, S, n_frames, topk, temperature, cfg_scale, reason_card, seed, is_ref,
             device="cpu", explicit_noise=False):
    """Drive prefill + n_frames of generate_frame the way tts_task.py:208-285 / asr_task.py:630-688 do.
    is_ref=True: `model` has the reference's Model_stage3 API (reference itself or the product);
    is_ref=False: `model` is the Stage3Oracle.  explicit_noise: draw the CPU noise stream here and pass it
    through the `noise=` argument (product under test / oracle) instead of letting the model draw."""
    gen = torch.Generator().manual_seed(seed)
    prompt_kind = "mixed" if kind == "asr_decode" else kind
    tokens, mask = make_prompt(prompt_kind, cfg, B, S, gen, reason_card)
    if B == 2 and cfg_scale > 1.0:
        tokens[1] = tokens[0]
        tokens[1, :, -1] = 7  # text_pad-like negative prompt (tts_task.py:171-190)
        mask[1] = mask[0]
    tokens, mask = tokens.to(device), mask.to(device)
    model.reset_caches()
    pos = torch.arange(0, S, device=device).unsqueeze(0).repeat(B, 1)
    nq = cfg.num_codebooks
    rows = 1 if (cfg_scale > 1.0 and B > 1) else B
    with torch.inference_mode():
        if is_ref:
            model.forward_prefix(tokens[:, :-1], labels=tokens[:, 1:, :-1], tokens_mask=mask, loss_mask=mask, input_pos=pos[:, :-1])
        else:
            model.forward_prefix(tokens[:, :-1], mask, pos[:, :-1])
        curr_tokens, curr_mask = tokens[:, -1:], mask[:, -1:]
        curr_pos = torch.tensor([S - 1], dtype=torch.long, device=device)
        maxp1 = S
        torch.manual_seed(888)  # multi_task_inference.py:596 sampler seed
        frames, text_logits, ci_logits, h_final = [], [], [], []
        for f in range(n_frames):
            forbid = 0 if f < n_frames // 2 else reason_card
            noise = cpu_noise_for_frame(rows, cfg.backbone.padded_vocab_size, cfg.audio_vocab, nq) if explicit_noise else None
            if is_ref:
                kw = {}
                if noise is not None:
                    kw["noise"] = torch.cat([n.reshape(-1) for n in noise]).to(device)
                s = model.generate_frame(curr_tokens, curr_mask, input_pos=curr_pos, input_pos_maxp1=maxp1,
                                         temperature=temperature, topk=topk, forbid_prefix=forbid, cfg_scale=cfg_scale, **kw)
            else:
                dbg = {}
                s = model.generate_frame(curr_tokens, curr_mask, curr_pos, maxp1, temperature, topk, forbid, cfg_scale,
                                         noise=noise, debug=dbg)
                text_logits.append(dbg["text_logits"])
                ci_logits.append(torch.stack(dbg["ci_logits"]))
                h_final.append(dbg["h_final"])
            frames.append(s.clone())
            # feed back like tts_task.py:276-279 (audio tokens in cols 0..nq-1, text token in col nq, audio-step mask)
            # or like asr_task.py:671-679 (text token only, audio slots zero, text-step mask)
            audio = s[:, 1:].long()
            text = s[:, 0:1].long()
            if kind == "asr_decode":
                curr_tokens = torch.cat([torch.zeros_like(audio), text], dim=-1).unsqueeze(1)
                curr_mask = torch.cat([torch.zeros_like(audio).bool(), torch.ones(B, 1, device=device).bool()], dim=1).unsqueeze(1)
            else:
                curr_tokens = torch.cat([audio, text], dim=-1).unsqueeze(1)
                curr_mask = torch.cat([torch.ones_like(audio).bool(), torch.zeros(B, 1, device=device).bool()], dim=1).unsqueeze(1)
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
