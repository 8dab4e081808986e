"""Generate tests/golden/codec_golden.pt from the UNMODIFIED reference codec (tools/tokenizer/MimiCodec, the importable
twin of llm_modules/*) and assert oracle/codec_oracle.py is bit-identical to it on CPU.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_codec
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import codec_oracle as CO  # noqa: E402
from oracle.ref_shims import REF_ROOT  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "codec_golden.pt")


def codec_cfgs():
    return {
        # hs = 32 heads, resample stride 4 (k=8), 6 quantizers
        "tiny": CO.MimiCfg(n_filters=8, encoder_rates=[8, 5, 4, 3], latent_dim=128, codebook_size=64, codebook_dim=32,
                           rvq_layers=6, num_heads=4, num_layers=2, context=20),
        # hs = 64 like the full config, stride 2 (k=4), context shorter than the sequence
        "mid": CO.MimiCfg(n_filters=16, encoder_rates=[8, 6, 5, 4], latent_dim=256, codebook_size=256, codebook_dim=64,
                          rvq_layers=8, num_heads=4, num_layers=2, context=12),
    }


def build_reference(cfg: CO.MimiCfg, sd):
    os.environ.setdefault("NO_TORCH_COMPILE", "1")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from tools.tokenizer.MimiCodec.model.models.MimiCodec import MimiCodec

    m = MimiCodec(sample_rate=cfg.sample_rate, n_filters=cfg.n_filters, encoder_rates=cfg.encoder_rates, compress=cfg.compress,
                  latent_dim=cfg.latent_dim, codebook_size=cfg.codebook_size, codebook_dim=cfg.codebook_dim,
                  rvq_layers=cfg.rvq_layers, num_heads=cfg.num_heads, num_layers=cfg.num_layers, layer_scale=cfg.layer_scale,
                  context=cfg.context, target_frame_rate=cfg.target_frame_rate)
    full = m.state_dict()
    for k, v in sd.items():
        assert k in full and full[k].shape == v.shape, (k, v.shape, full.get(k, torch.empty(0)).shape)
        full[k] = v
    m.load_state_dict(full, strict=True)
    return m.eval()


def main():
    torch.set_num_threads(8)
    out = {}
    for name, cfg in codec_cfgs().items():
        sd = CO.random_mimi_state_dict(cfg, seed=4321)
        ref = build_reference(cfg, sd)
        orc = CO.MimiOracle(cfg, sd)
        g = torch.Generator().manual_seed(11)
        for B, T in ((2, 3 * cfg.hop_length * cfg.resample_stride + 37), (1, 9 * cfg.hop_length * cfg.resample_stride)):
            wav = torch.randn(B, 1, T, generator=g) * 0.3
            with torch.no_grad():
                codes_ref = ref.encode(wav)
                wav_ref = ref.decode(codes_ref)
                codes_o = orc.encode(wav)
                wav_o = orc.decode(codes_o)
                lat = orc.encode_latent(wav)
            assert torch.equal(codes_ref, codes_o), f"{name}: oracle codes != reference codes"
            assert torch.equal(wav_ref, wav_o), f"{name}: oracle waveform != reference ({(wav_ref - wav_o).abs().max()})"
            # margin of every argmin (distance gap between best and second-best code) - slack for fp32 reordering
            out[f"{name}_B{B}_T{T}"] = dict(cfg_name=name, wav=wav, codes=codes_ref, recon=wav_ref, latent=lat)
            print(f"[ok] {name} B={B} T={T}: codes {tuple(codes_ref.shape)} recon {tuple(wav_ref.shape)} bit-exact")
        out[f"__checksum_{name}"] = {k: float(v.double().sum()) for k, v in sd.items()}
    torch.save(out, GOLDEN)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN) / 1e6, "MB")


if __name__ == "__main__":
    main()
