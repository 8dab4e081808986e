"""Reasoning encoder (`AudioThinking`: dim 768, 6 heads of 128, 5 blocks) at the reference's batch of 6 windows of 30 s: 1500 Whisper
frames + 750 BEST-RQ frames per window -> 900 rows -> 150 query tokens; random weights; encoder alone and with the 8-level quantiser."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.audio_thinking import AudioThinking  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    m = AudioThinking(device=dev)
    with torch.no_grad():
        for p in m.reasoning_vq.parameters():
            p.normal_()
    whisper, mu = torch.randn(B, 1024, 1500, device=dev), torch.randn(B, 1024, 750, device=dev)
    T, D, F, L = 750, 768, 3072, 5
    Tn = T + T // 5
    flop = B * (2.0 * T * 1024 * 2048 + 2.0 * T * 2048 * D + L * (2.0 * Tn * (4 * D * D + 3 * D * F) + 4.0 * Tn * Tn * D))
    for what, fn in (("query tokens", lambda: m.query_tokens(whisper, mu)), ("query tokens + residual VQ", lambda: m.encode_reasoning_part(whisper, mu)[0])):
        for _ in range(2):
            y = fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5
        e0.record()
        for _ in range(n):
            y = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(json.dumps({"what": "AudioThinking encoder, %d x 30 s windows, fp32 class: %s" % (B, what), "ms": round(ms, 3), "launches": m.last_launch_count(),
                          "algorithmic_TFLOP": round(flop / 1e12, 3), "TFLOPs_fp32_equivalent": round(flop / ms / 1e9, 1),
                          "x_realtime": round(B * 30.0 / (ms * 1e-3), 1), "finite": bool(torch.isfinite(y).all())}))


if __name__ == "__main__":
    main()
