"""Flow-matching decoder at production size on one B200: the DiT estimator of models/model_config.json (32 layers, 24 x 64
heads, in 1040 -> out 136) on the classifier-free-guidance batch of a 20 s window (2 x 500 frames = 1000 rows) and the Euler
solver (reason_tokenizer.py:273: guidance 1.5; test.sh: 10 steps).  CUDA events on the launching stream after warm-up; random
weights (the checkpoint is not in the repository).  Prints JSON lines.

    python tools/measure_dit.py [--reps 10]      (the CPU-oracle baseline of the same call is a leg of bench.py)
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from uniaudio2_b200 import _lib  # noqa: E402
from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.AudioDiffusion1D import BASECFM  # noqa: E402
from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.transformer_1d_flow import Transformer1DModel  # noqa: E402

PROD = dict(num_attention_heads=24, attention_head_dim=64, in_channels=1040, out_channels=136, num_layers=32, attention_bias=True,
            activation_fn="gelu-approximate", norm_type="ada_norm_single", norm_elementwise_affine=False, norm_eps=1e-6,
            num_embeds_ada_norm=1000)


def flops(B, T, layers=32, D=1536, I=1040, O=136):
    M = B * T
    lin = layers * 24 * D * D * M                      # qkv 3 D^2 + out D^2 + ff 8 D^2, 2 FLOP per MAC
    attn = layers * 4 * T * T * D * B
    proj = 2 * M * (3 * I * D + D * D + 3 * D * O + O * O)
    return lin + attn + proj


def timed(fn, reps):
    for _ in range(3):
        fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * reps)]
    for i in range(reps):
        ev[2 * i].record()
        fn()
        ev[2 * i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(reps))
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--frames", type=int, default=500)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--once", action="store_true", help="one estimator call only (profiling driver)")
    ap.add_argument("--bf16", action="store_true", help="also time the bf16 option (many-row linears on bf16 operands)")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    m = Transformer1DModel(device=dev, **PROD)
    T = a.frames
    x = torch.randn(2, T, 1040, device=dev)
    t = torch.full((2,), 0.35, device=dev)
    if a.once:
        if a.bf16:
            m.set_option("bf16", 1)
        m(x, timestep=t)
        m(x, timestep=t)
        torch.cuda.synchronize()
        print("launches per estimator call:", m.last_launch_count())
        return
    fl = flops(2, T)
    if True:
        ms = timed(lambda: m(x, timestep=t), a.reps)
        print(json.dumps(dict(what="estimator call, CFG batch 2 x %d frames, 3xTF32 (fp32 weights split on chip)" % T,
                              launches=m.last_launch_count(), ms=round(ms, 3), algorithmic_TFLOP=round(fl / 1e12, 3),
                              fp32_equiv_TFLOPs=round(fl / ms / 1e9, 1), tf32_mma_TFLOPs=round(3 * fl / ms / 1e9, 1))))
    if a.bf16:
        ref = m(x, timestep=t).sample
        m.set_option("bf16", 1)
        ms = timed(lambda: m(x, timestep=t), a.reps)
        got = m(x, timestep=t).sample
        m.set_option("flash_attn", 0)
        ms_simt = timed(lambda: m(x, timestep=t), a.reps)
        m.set_option("flash_attn", 1)
        m.set_option("bf16", 0)
        print(json.dumps(dict(what="estimator call, bf16 option with the fp32 SIMT attention", ms=round(ms_simt, 3))))
        print(json.dumps(dict(what="estimator call, bf16 option (tensor-core attention)", launches=m.last_launch_count(), ms=round(ms, 3),
                              bf16_mma_TFLOPs=round(fl / ms / 1e9, 1), max_abs_diff_vs_3xtf32=float((got - ref).abs().max()),
                              out_scale=float(ref.abs().max()))))
    cfm = BASECFM(m)
    z = torch.randn(1, T, 136, device=dev)
    ic = torch.zeros(1, T, 136, device=dev)
    mu = torch.randn(1, T, 768, device=dev)
    t_span = torch.linspace(0, 1, a.steps + 1)
    ms = timed(lambda: cfm.solve_euler(z, ic, 0, t_span, mu, None, 1.5), max(3, a.reps // 3))
    print(json.dumps(dict(what="solve_euler, %d steps, 20 s window" % a.steps, ms=round(ms, 2), audio_seconds=T / 25.0,
                          x_realtime=round(T / 25.0 / (ms / 1e3), 1))))


if __name__ == "__main__":
    main()
