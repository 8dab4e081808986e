"""Streaming-step latency of the Moshi-family transformer drop-in and of sample_token on one B200 (CUDA events, after
warm-up, L2 flushed between steps by streaming weights larger than L2 where the model is small).  Prints one JSON line per
shape: time per step, algorithmic bytes (weights touched once + KV read, fp32) and the implied GB/s.

    python tools/measure_stream.py [--steps 200]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from uniaudio2_b200.llm_modules.transformer import StreamingTransformer  # noqa: E402
from uniaudio2_b200.llm_utils.sampling import sample_token  # noqa: E402

SHAPES = {
    # Mimi's encoder/decoder transformer (MimiCodec.py:54-58) run one 12.5 Hz frame at a time
    "mimi_transformer": dict(d_model=512, num_heads=8, num_layers=8, dim_feedforward=2048, causal=True, context=250,
                             positional_embedding="rope", norm="layer_norm", layer_scale=0.01, gating="none"),
    # depformer-shaped: per-step weights, capacity = weights_per_step
    "depformer_like": dict(d_model=1024, num_heads=16, num_layers=6, dim_feedforward=4096, causal=True, context=None,
                           positional_embedding="none", norm="rms_norm_f32", gating="silu", weights_per_step=8),
    # temporal-transformer-shaped slice (8 of 32 layers of a d = 4096 model), context 3000
    "temporal_like_8L": dict(d_model=4096, num_heads=32, num_layers=8, dim_feedforward=16384, causal=True, context=3000,
                             positional_embedding="rope", norm="rms_norm_f32", gating="silu"),
}


def step_bytes(m: StreamingTransformer, B: int, keys: int) -> int:
    per_step = sum(p.numel() for p in m.parameters()) // (m.weights_per_step or 1)
    kv = 2 * B * m.d_model * keys * m.num_layers
    return 4 * (per_step + kv)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--batch", type=int, default=1)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    for name, kw, graph in [(n, k, g) for n, k in SHAPES.items() for g in (0, 1)]:
        m = StreamingTransformer(device=dev, **kw)
        m.set_option("graph", graph)
        B = a.batch
        x = torch.randn(B, 1, kw["d_model"], device=dev)
        wps = kw.get("weights_per_step", 0)
        cap = kw["context"] or wps
        with m.streaming(B):
            n_warm = cap + 8 if not wps else wps
            for i in range(n_warm):  # fill the ring so that every timed step reads `cap` keys
                if wps and i % wps == 0 and i:
                    m.reset_streaming()
                m(x)
            if wps:
                m.reset_streaming()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * a.steps)]
            for i in range(a.steps):
                if wps and i % wps == 0 and i:
                    m.reset_streaming()
                flush.zero_()
                ev[2 * i].record()
                m(x)
                ev[2 * i + 1].record()
            torch.cuda.synchronize()
        ts = sorted(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(a.steps))
        med = ts[len(ts) // 2]
        keys = cap if not wps else (wps + 1) // 2
        by = step_bytes(m, B, keys)
        print(json.dumps(dict(shape=name, graph_pdl=graph, launches=m.last_launch_count(), batch=B, ms_per_step=round(med, 4),
                              p10=round(ts[len(ts) // 10], 4),
                              algorithmic_MB=round(by / 1e6, 2), GBps=round(by / med / 1e6, 1), l2="flushed between steps")))
        del m
        torch.cuda.empty_cache()
    for rows, V, k in ((8, 2048, 250), (1, 32000, 25), (1, 128256, 50)):
        lg = torch.randn(rows, V, device=dev) * 2
        for _ in range(5):
            sample_token(lg, use_sampling=True, temp=0.8, top_k=k)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * a.steps)]
        for i in range(a.steps):
            ev[2 * i].record()
            sample_token(lg, use_sampling=True, temp=0.8, top_k=k)
            ev[2 * i + 1].record()
        torch.cuda.synchronize()
        ts = sorted(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(a.steps))
        print(json.dumps(dict(shape=f"sample_token rows={rows} V={V} top_k={k}", ms_per_call=round(ts[len(ts) // 2], 4),
                              note="includes torch's exponential_ draw of the (rows, k) noise")))


if __name__ == "__main__":
    main()
