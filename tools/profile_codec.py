"""Profiling driver (run under ncu): full-size Mimi-twin codec, batch x clip encode + decode.  Not a benchmark."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import codec_oracle as CO  # noqa: E402  (seeded weights / shapes only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--seconds", type=float, default=5.0)
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    from uniaudio2_b200.tools.tokenizer.MimiCodec.mimi_codec import MimiCodec

    cfg = CO.MimiCfg()
    sd = CO.random_mimi_state_dict(cfg, seed=7)
    m = MimiCodec(n_filters=cfg.n_filters, encoder_rates=cfg.encoder_rates, latent_dim=cfg.latent_dim, codebook_size=cfg.codebook_size,
                  codebook_dim=cfg.codebook_dim, rvq_layers=cfg.rvq_layers, num_heads=cfg.num_heads, num_layers=cfg.num_layers,
                  layer_scale=cfg.layer_scale, context=cfg.context, device="cuda")
    full = m.state_dict()
    full.update({k: v.cuda() for k, v in sd.items()})
    m.load_state_dict(full, strict=True)
    wav = torch.randn(a.batch, 1, int(a.seconds * 24000), device="cuda") * 0.1
    for _ in range(a.reps):
        codes = m.encode(wav)
        out = m.decode(codes)
    torch.cuda.synchronize()
    print("codes", tuple(codes.shape), "out", tuple(out.shape))


if __name__ == "__main__":
    main()
