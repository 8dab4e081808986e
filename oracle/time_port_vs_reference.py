"""TEST INFRASTRUCTURE ONLY (build container, CPU): time the UNMODIFIED reference Model_stage3 next to the oracle port
(oracle/llm_oracle.py) at FULL size on the same weights, same inputs, same thread count - the evidence that the CPU arm of
bench.py (`cpu_baseline.kind: "port"`) runs at the reference's own speed.  Result recorded in BASELINE.md.

    python -m oracle.time_port_vs_reference [--frames 8] [--threads 8]
"""
import argparse
import json
import time

import torch

from oracle import llm_oracle as O
from oracle.ref_shims import install_llm_shims

NQ, REASON_CARD, SEMANTIC_CARD = 8, 4100, 8200


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--threads", type=int, default=8)
    a = ap.parse_args()
    torch.set_num_threads(a.threads)
    model_new = install_llm_shims()
    t0 = time.perf_counter()
    torch.manual_seed(0)
    args = model_new.ModelArgs(llm_name="Llama-3.2-3B", decoder_name="Llama-3.2-300M", llm_pretrained_model="", audio_embeddings_path="",
                               audio_understanding_expert_path="", audio_semantic_vocab_size=SEMANTIC_CARD, audio_reason_vocab_size=REASON_CARD,
                               audio_num_codebooks=NQ)
    with torch.inference_mode():
        ref = model_new.Model_stage3(args)
        ref.eval()
        torch.nn.init.normal_(ref.audio_head, 0.0, 0.02)
        print(f"reference built in {time.perf_counter() - t0:.0f} s", flush=True)
        ref.setup_caches(1)
        sd = ref.state_dict()  # shared storage: the port runs on the very same tensors
        orc = O.Stage3Oracle(O.full_size_cfg(REASON_CARD, SEMANTIC_CARD), sd)
        orc.setup_caches(1)
        g = torch.Generator().manual_seed(888)
        S = 40
        tokens = torch.zeros(1, S, NQ + 1, dtype=torch.long)
        tokens[0, :, -1] = torch.randint(0, 128000, (S,), generator=g)
        mask = torch.zeros(1, S, NQ + 1, dtype=torch.bool)
        mask[..., -1] = True
        pos = torch.arange(S).unsqueeze(0)
        audio_mask = torch.cat([torch.ones(1, 1, NQ, dtype=torch.bool), torch.zeros(1, 1, 1, dtype=torch.bool)], -1)
        out = {}
        ids = {}
        for name in ("reference", "port", "reference", "port"):  # two rounds: the second one is reported (warm caches / allocator)
            model = ref if name == "reference" else orc
            model.reset_caches()
            t0 = time.perf_counter()
            if name == "reference":
                model.forward_prefix(tokens[:, :-1], labels=tokens[:, 1:, :-1], tokens_mask=mask, loss_mask=mask, input_pos=pos[:, :-1])
            else:
                model.forward_prefix(tokens[:, :-1], mask, pos[:, :-1])
            t_prefill = time.perf_counter() - t0
            ct, cm = tokens[:, -1:], mask[:, -1:]
            times, frames = [], []
            for f in range(a.frames):
                t0 = time.perf_counter()
                if name == "reference":
                    s = model.generate_frame(ct, cm, input_pos=torch.tensor([S - 1 + f]), input_pos_maxp1=S + f, temperature=1.0, topk=1, forbid_prefix=0)
                else:
                    s = model.generate_frame(ct, cm, torch.tensor([S - 1 + f]), S + f, 1.0, 1, 0)
                times.append(time.perf_counter() - t0)
                frames.append(s[0].clone())
                sl = s.long()
                ct = torch.cat([sl[:, 1:], sl[:, 0:1]], dim=-1).unsqueeze(1)
                cm = audio_mask
            out[name] = dict(prefill_s=round(t_prefill, 3), frame_ms=round(1e3 * sum(times[1:]) / len(times[1:]), 1))
            ids[name] = torch.stack(frames)
        out["ids_equal"] = bool(torch.equal(ids["reference"], ids["port"]))
        out["port_over_reference_frame_time"] = round(out["port"]["frame_ms"] / out["reference"]["frame_ms"], 3)
        out["threads"] = a.threads
        out["frames"] = a.frames
        print(json.dumps(out))


if __name__ == "__main__":
    main()
