"""Profiling driver (run under ncu): full-size model, one 39-position prefill + N eager generate_frame calls
(graph replay off so every kernel is a separate launch), TTS-10 s config.  Not a benchmark."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--pdl", type=int, default=0)
    ap.add_argument("--graph", type=int, default=0)
    ap.add_argument("--start-pos", type=int, default=0, help="extra context to emulate later frames")
    ap.add_argument("--batch", type=int, default=1, help="rows per frame (32 = config 3's batched frames on the tensor-core path)")
    a = ap.parse_args()
    from uniaudio2_b200.evaluation.tts_task import Generator, default_train_args
    from uniaudio2_b200.llm_models.model_new import Model_stage3

    dev = torch.device("cuda", 0)
    with torch.inference_mode():
        model = Model_stage3(bench.model_args(), device=dev)
        bench.init_weights_(model, 0)
        gen = Generator(model, default_train_args(bench.REASON_CARD, bench.SEMANTIC_CARD))
        model.set_option("graph", a.graph)
        model.set_option("pdl", a.pdl)
        tp, text = bench.synthetic_prompt(0)
        tokens, mask = gen.prepare_tts_task(tp, text)
        B = a.batch
        tokens, mask = tokens.unsqueeze(0).to(dev).repeat(B, 1, 1), mask.bool().unsqueeze(0).to(dev).repeat(B, 1, 1)
        S = tokens.size(1)
        pos = torch.arange(S, device=dev).unsqueeze(0).repeat(B, 1)
        if B > 1:
            model.setup_caches(B)
        model.reset_caches()
        model.forward_prefix(tokens[:, :-1], None, mask, None, input_pos=pos[:, :-1], input_pos_maxp1=S - 1)
        print("prefill launches", model.last_launch_count())
        ct, cm = tokens[:, -1:], mask[:, -1:]
        am = torch.cat([torch.ones(B, 1, 8, dtype=torch.bool), torch.zeros(B, 1, 1, dtype=torch.bool)], -1).to(dev)
        for f in range(a.frames):
            s = model.generate_frame(ct, cm, input_pos=S - 1 + f + a.start_pos, input_pos_maxp1=S + f + a.start_pos,
                                     temperature=0.9, topk=50, forbid_prefix=0)
            sl = s.long()
            ct = torch.cat([sl[:, 1:], sl[:, 0:1]], -1).unsqueeze(1)
            cm = am
        torch.cuda.synchronize()
        print("frame launches", model.last_launch_count(), s.tolist()[:1])


if __name__ == "__main__":
    main()
