"""Dev tool: per-op phase times of the persistent chain kernel (global stack launch), from the clock64 samples the kernel
records for CTA 0 and CTA G/2 when option chain_profile is set.  Phases per op: prologue (op start -> activations staged),
stream (-> last weight chunk consumed), tail (-> op done), barrier (-> grid barrier passed)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    from uniaudio2_b200.evaluation.tts_task import Generator, default_train_args
    from uniaudio2_b200.llm_models.model_new import Model_stage3

    dev = torch.device("cuda", 0)
    mhz = 1965.0
    with torch.inference_mode():
        model = Model_stage3(bench.model_args(), device=dev)
        bench.init_weights_(model, 0)
        gen = Generator(model, default_train_args(bench.REASON_CARD, bench.SEMANTIC_CARD))
        model.set_option("chain", 1)
        model.set_option("chain_profile", 1)
        tp, text = bench.synthetic_prompt(0)
        tokens, mask = gen.prepare_tts_task(tp, text)
        tokens, mask = tokens.unsqueeze(0).to(dev), mask.bool().unsqueeze(0).to(dev)
        S = tokens.size(1)
        pos = torch.arange(S, device=dev).unsqueeze(0)
        model.reset_caches()
        model.forward_prefix(tokens[:, :-1], None, mask, None, input_pos=pos[:, :-1], input_pos_maxp1=S - 1)
        ct, cm = tokens[:, -1:], mask[:, -1:]
        am = torch.cat([torch.ones(1, 1, 8, dtype=torch.bool), torch.zeros(1, 1, 1, dtype=torch.bool)], -1).to(dev)
        for f in range(4):
            s = model.generate_frame(ct, cm, input_pos=S - 1 + f + 60, input_pos_maxp1=S + f + 60, temperature=0.9, topk=50, forbid_prefix=0)
            sl = s.long()
            ct = torch.cat([sl[:, 1:], sl[:, 0:1]], -1).unsqueeze(1)
            cm = am
        torch.cuda.synchronize()
        raw = model.debug_buffer("chain_prof", 1).view(torch.int64).cpu().view(2, -1, 5)
    for c in range(2):
        t = raw[c]
        n = int((t[:, 0] > 0).sum())
        t = t[:n].double() / mhz  # us
        print(f"== CTA {'0' if c == 0 else 'G/2'}: {n} ops, chain wall {float(t[n - 1, 3] - t[0, 0]):.1f} us")
        gemv = t[:, 1] > 0
        pro = (t[:, 1] - t[:, 0])[gemv]
        stream = (t[:, 2] - t[:, 1])[gemv]
        tail = (t[:, 3] - t[:, 2])[gemv]
        other = (t[:, 3] - t[:, 0])[~gemv]
        bar = (t[:-1, 4] - t[:-1, 3])
        print(f"   gemv ops {int(gemv.sum())}: prologue {pro.mean():.2f} us, stream {stream.mean():.2f} us, tail {tail.mean():.2f} us (sums {pro.sum():.0f} / {stream.sum():.0f} / {tail.sum():.0f} us)")
        print(f"   non-gemv ops {int((~gemv).sum())}: {other.mean():.2f} us each (sum {other.sum():.0f} us)")
        print(f"   barrier wait: mean {bar.mean():.2f} us, sum {bar.sum():.0f} us; after gemv {bar[gemv[:-1]].mean():.2f}, after non-gemv {bar[~gemv[:-1]].mean():.2f}")
        print("   first ops (us): start-rel | prologue stream tail | barrier")
        for i in range(min(n, 14)):
            print(f"   op {i:3d} @{float(t[i, 0] - t[0, 0]):8.2f} | {float(t[i, 1] - t[i, 0]) if t[i, 1] > 0 else -1:6.2f} {float(t[i, 2] - t[i, 1]) if t[i, 1] > 0 else -1:7.2f} "
                  f"{float(t[i, 3] - (t[i, 2] if t[i, 1] > 0 else t[i, 0])):6.2f} | {float(t[i, 4] - t[i, 3]):6.2f}")


if __name__ == "__main__":
    main()
