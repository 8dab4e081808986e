"""Whisper-medium encoder (24 layers, d 1024, 16 x 64 heads, 30 s window = 1500 frames) at the reference's batch of 6 windows
(reason_tokenizer.py:86 batch_size=6), random weights: time per call in both arithmetic modes and tensor TFLOP/s."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.modeling_whisper import WhisperConfig, WhisperModel  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    m = WhisperModel(WhisperConfig(), device=dev).encoder
    mel = torch.randn(B, 80, 3000, device=dev)
    P, D, F, L = 1500, 1024, 4096, 24
    flop = B * (2.0 * 3000 * 240 * D + 2.0 * P * 3 * D * D + L * (2.0 * P * (4 * D * D + 2 * D * F) + 4.0 * P * P * D))
    res = {}
    for mode, bf in (("fp32_class", 0), ("bf16", 1)):
        m.set_option("bf16", bf)
        for _ in range(2):
            y = m(mel).last_hidden_state
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5
        e0.record()
        for _ in range(n):
            y = m(mel).last_hidden_state
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        res[mode] = y
        print(json.dumps({"what": "Whisper-medium encoder, %d x 30 s windows, %s" % (B, mode), "ms": round(ms, 3), "launches": m.last_launch_count(),
                          "algorithmic_TFLOP": round(flop / 1e12, 3), "TFLOPs": round(flop / ms / 1e9, 1),
                          "x_realtime": round(B * 30.0 / (ms * 1e-3), 1)}))
    print(json.dumps({"bf16_vs_fp32_class_max_abs": float((res["bf16"] - res["fp32_class"]).abs().max()),
                      "out_scale": float(res["fp32_class"].abs().max())}))


if __name__ == "__main__":
    main()
