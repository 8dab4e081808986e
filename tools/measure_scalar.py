"""SQ-codec wave decoder (ScalarModel.decode, scalar24k.py:403-407) on one 20 s window - the last step of `--stage all`'s tokens -> wav
path after the flow-matching decoder: latent (B, 136, 500) -> (B, 1, 480 000).  Geometry of the shipped sqcodec config as restated in
oracle/scalar_oracle.py's defaults (init_channel 48, up-sampling 6 x 5 x 4 x 4 x 2, res kernel 7 with dilations 1 3 5 7 9); random weights."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.scalar24k import ScalarModel  # noqa: E402

CFG = dict(num_bands=1, sample_rate=24000, causal=True, num_samples=1, downsample_factors=[2, 4, 4, 5, 6], downsample_kernel_sizes=[4, 8, 8, 10, 12],
           upsample_factors=[6, 5, 4, 4, 2], upsample_kernel_sizes=[12, 10, 8, 8, 4], latent_hidden_dim=136, default_kernel_size=7,
           delay_kernel_size=5, init_channel=48, res_kernel_size=7)


def build(dev):
    torch.manual_seed(0)
    m = ScalarModel(device=dev, **CFG)
    with torch.no_grad():
        for k, p in m.named_parameters():
            if k.endswith("weight_v"):
                fan = p[0].numel()
                p.copy_(torch.randn_like(p) / fan ** 0.5)
            elif k.endswith("weight_g"):
                p.fill_(0.7)
            elif k.endswith("activation1.weight") or k.endswith("activation2.weight") or k.endswith("activation.weight"):
                p.fill_(0.25)
            else:
                p.copy_(torch.randn_like(p) * 0.05)
    return m


def flops(B, T):
    c0, ups = CFG["init_channel"], CFG["upsample_factors"]
    n = len(ups)
    fl, t, ch = 2.0 * 136 * 5 * c0 * 2 ** n * T, T, c0 * 2 ** n
    for i, s in enumerate(ups):
        co = ch // 2
        t *= s
        fl += 2.0 * ch * co * 2 * t            # transposed conv: 2 taps per output sample
        fl += 5 * (2.0 * 7 * co * co + 2.0 * co * co) * t
        ch = co
    fl += 2.0 * 7 * ch * t
    return B * fl


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    once = "--once" in sys.argv
    dev = torch.device("cuda", 0)
    m = build(dev)
    x = torch.randn(B, 136, 500, device=dev) * 0.5
    y = m.decode(x)
    if once:
        torch.cuda.synchronize()
        print("out", tuple(y.shape))
        return
    for _ in range(2):
        y = m.decode(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        y = m.decode(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = flops(B, 500)
    print(json.dumps({"what": "ScalarModel.decode, %d x 20 s window" % B, "ms": round(ms, 2), "GFLOP": round(fl / 1e9, 1),
                      "fp32_equiv_TFLOPs": round(fl / ms / 1e9, 1), "x_realtime": round(B * 20.0 / (ms * 1e-3), 1), "out": list(y.shape),
                      "finite": bool(torch.isfinite(y).all())}))


if __name__ == "__main__":
    main()
