"""Measure the SURVEY section 8(d) configurations that `bench.py` does not quote on its headline line (1 GPU, full-size
random-weight model, synthetic inputs).  Prints one JSON object per measurement; not a benchmark contract.

  prefill      forward_prefix at (B=32, S=188) [config 3] and (B=1, S=540) [30 s ASR prompt], SIMT tiled GEMM vs tcgen05 3xTF32
  caption32    config 3: batched prefill + 64 greedy text frames at B = 32
  ttm500       config 4, one rank's share: one prompt, 500 frames (the reference's cap)
  codec_sweep  config 5: clip length x batch sweep of the Mimi-twin codec

    python tools/measure_configs.py [--only prefill,caption32,ttm500,codec_sweep]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

P_GLOBAL_LAYERS = 33 * 100.66e6  # parameters of the 33 global layers


def emit(**kw):
    print(json.dumps(kw), flush=True)


def make_prompt(B, S, n_text, dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    tokens = torch.zeros(B, S, bench.NQ + 1, dtype=torch.long)
    mask = torch.zeros(B, S, bench.NQ + 1, dtype=torch.bool)
    tokens[:, :n_text, -1] = torch.randint(0, 128000, (B, n_text), generator=g)
    mask[:, :n_text, -1] = True
    tokens[:, n_text:, :-1] = torch.randint(0, bench.REASON_CARD + bench.SEMANTIC_CARD, (B, S - n_text, bench.NQ), generator=g)
    mask[:, n_text:, :-1] = True
    return tokens.to(dev), mask.to(dev)


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="prefill,caption32,ttm500,codec_sweep")
    a = ap.parse_args()
    only = set(a.only.split(","))
    from uniaudio2_b200 import _lib
    from uniaudio2_b200.llm_models.model_new import Model_stage3

    dev = torch.device("cuda", 0)
    L = _lib.lib()
    if only & {"prefill", "caption32", "ttm500"}:
        with torch.inference_mode():
            model = Model_stage3(bench.model_args(), device=dev)
            bench.init_weights_(model, 0)
            model.setup_caches(32)

            def prefill(tokens, mask):
                B, S = tokens.shape[:2]
                pos = torch.arange(S, device=dev).unsqueeze(0).repeat(B, 1)
                model.reset_caches()
                model.forward_prefix(tokens[:, :-1], None, mask, None, input_pos=pos[:, :-1], input_pos_maxp1=S - 1)

            if "prefill" in only:
                for (B, S, n_text, what) in ((32, 189, 10, "config 3: 32 caption prompts of 10 text + 52 reason + 127 semantic frames"),
                                             (1, 540, 10, "30 s ASR/caption prompt"), (4, 540, 10, "4 x 30 s prompts")):
                    tokens, mask = make_prompt(B, S, n_text, dev)
                    rows = B * (S - 1)
                    for tc in ((0, 1) if L.ua2_set_global_option(b"tc_gemm", 1) == 0 else (0,)):
                        _lib.check(L.ua2_set_global_option(b"tc_gemm", tc))
                        ms = timed(lambda: prefill(tokens, mask), 2)
                        emit(kind="prefill", B=B, S=S, rows=rows, what=what, gemm="tcgen05 3xTF32" if tc else "fp32 SIMT tiles",
                             ms=round(ms, 2), rows_per_s=round(rows / (ms * 1e-3), 1),
                             fp32_equiv_tflops=round(2.0 * rows * P_GLOBAL_LAYERS / (ms * 1e-3) / 1e12, 2), launches=model.last_launch_count())
                    _lib.check(L.ua2_set_global_option(b"tc_gemm", 1))

            if "caption32" in only:
                B, S, NF = 32, 189, 64
                tokens, mask = make_prompt(B, S, 10, dev)
                text_mask = torch.zeros(B, 1, bench.NQ + 1, dtype=torch.bool, device=dev)
                text_mask[..., -1] = True
                have_tc = L.ua2_set_global_option(b"tc_gemm", 1) == 0
                # (tensor cores, rows from which frames use them)
                for (tc, min_rows) in (((0, 128), (1, 128), (1, 16)) if have_tc else ((0, 128),)):
                    _lib.check(L.ua2_set_global_option(b"tc_gemm", tc))
                    _lib.check(L.ua2_set_global_option(b"tc_min_rows", min_rows))

                    def run():
                        prefill(tokens, mask)
                        ct, cm = tokens[:, -1:], mask[:, -1:]
                        for f in range(NF):
                            s = model.generate_frame(ct, cm, input_pos=S - 1 + f, input_pos_maxp1=S + f, temperature=1.0, topk=1,
                                                     forbid_prefix=0)
                            ct = torch.zeros(B, 1, bench.NQ + 1, dtype=torch.long, device=dev)
                            ct[:, 0, -1] = s[:, 0].long()  # ASR/caption loop feeds the text token back (asr_task.py:660-680)
                            cm = text_mask
                    ms = timed(run, 1)
                    pre = timed(lambda: prefill(tokens, mask), 1)
                    emit(kind="caption32", B=B, S=S, frames=NF,
                         gemm=("tcgen05 3xTF32 (weights split on chip) prefill" + (" + frames" if min_rows <= B else ""))
                         if tc else "fp32 SIMT tiles / skinny kernels", total_ms=round(ms, 1),
                         prefill_ms=round(pre, 1), ms_per_frame=round((ms - pre) / NF, 2),
                         text_tokens_per_s=round(B * NF / (ms * 1e-3), 1), clips_per_s=round(B / (ms * 1e-3), 2))
                _lib.check(L.ua2_set_global_option(b"tc_gemm", 1))
                _lib.check(L.ua2_set_global_option(b"tc_min_rows", 32))

            if "ttm500" in only:
                tp, text = bench.synthetic_prompt(0)
                from uniaudio2_b200.evaluation.tts_task import Generator, default_train_args
                gen = Generator(model, default_train_args(bench.REASON_CARD, bench.SEMANTIC_CARD), is_cfg=False, tag="caption")
                torch.manual_seed(888)
                for _ in range(2):
                    t0 = time.perf_counter()
                    r, s = gen.generate_tts(tp, "TTM", text_token=text, temperature=bench.TEMPERATURE, topk=bench.TOPK,
                                            fixed_schedule=(125, 375))
                    torch.cuda.synchronize()
                    dt = time.perf_counter() - t0
                emit(kind="ttm500", frames=500, audio_tokens=4000, seconds=round(dt, 3), audio_tokens_per_s=round(4000 / dt, 1),
                     note="config 4 gives each of 8 ranks one such prompt; weak scaling, no data-path collective")
            del model
            torch.cuda.empty_cache()

    if "codec_sweep" in only:
        m = bench.make_codec(dev)
        g = torch.Generator().manual_seed(0)
        for clip_s in (1, 5, 30):
            for batch in (1, 4, 16, 64):
                if clip_s * batch > 640:
                    continue
                wav = (torch.randn(batch, 1, clip_s * 24000, generator=g) * 0.1).to(dev)
                codes = m.encode(wav)
                enc = timed(lambda: m.encode(wav), 3)
                dec = timed(lambda: m.decode(codes), 3)
                emit(kind="codec_sweep", clip_s=clip_s, batch=batch, encode_ms=round(enc, 2), decode_ms=round(dec, 2),
                     rtf_x_realtime=round(clip_s * batch / ((enc + dec) * 1e-3), 1))


if __name__ == "__main__":
    main()
