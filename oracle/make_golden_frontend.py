"""TEST INFRASTRUCTURE ONLY.  Writes tests/golden/frontend_golden.pt: outputs of the REAL third-party classes behind the waveform
front-end and the WavLM encoder of ReasoningCodec_film's tokenize direction, run in this container -

    torchaudio.transforms.Resample(24000, 16000)            reason_tokenizer.py:37, AudioDiffusion1D.py:227
    transformers.WhisperFeatureExtractor()                  reason_tokenizer.py:36, :67-72  (defaults = whisper-medium's preprocessor_config)
    transformers.WavLMModel(WavLMConfig(...))               AudioDiffusion1D.py:226, :359-370 (random weights: the checkpoints are not here)

so that the GPU tests (and the CPU tests of the oracles) compare against the classes the reference calls, not only against this
repo's restatements of them.  Run from the repo root:  python -m oracle.make_golden_frontend
Versions in this image: torchaudio 2.11.0, transformers 5.5.0 (the reference pins transformers==4.57.0, pyproject.toml:25).
"""
import os
import warnings

import torch

SMALL = dict(hidden_size=64, num_attention_heads=2, intermediate_size=128, num_hidden_layers=3, conv_dim=(64, 64, 64), conv_kernel=(10, 3, 2),
             conv_stride=(5, 2, 2), conv_bias=False, num_conv_pos_embeddings=16, num_conv_pos_embedding_groups=4, num_buckets=320,
             max_bucket_distance=800, layer_norm_eps=1e-5)
# head size 64 and 48 channels per positional-convolution group like the checkpoints, k = 128 taps, two strided k = 3 convolutions, conv biases
MID = dict(hidden_size=192, num_attention_heads=3, intermediate_size=256, num_hidden_layers=2, conv_dim=(64, 64, 64, 64), conv_kernel=(10, 3, 3, 2),
           conv_stride=(5, 2, 2, 2), conv_bias=True, num_conv_pos_embeddings=128, num_conv_pos_embedding_groups=4, num_buckets=320,
           max_bucket_distance=800, layer_norm_eps=1e-5)


def build_hf_wavlm(cfg, seed):
    """transformers.WavLMModel with the seeded weights of oracle/wavlm_oracle.py::random_state_dict (strict load: the key set and the
    shapes of the restatement are the real class's)."""
    from transformers import WavLMConfig, WavLMModel

    from oracle import wavlm_oracle as WO

    hc = WavLMConfig(**{k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()}, feat_extract_norm="group", do_stable_layer_norm=False,
                     hidden_dropout=0.0, attention_dropout=0.0, activation_dropout=0.0, feat_proj_dropout=0.0, layerdrop=0.0, apply_spec_augment=False)
    m = WavLMModel(hc).eval()
    m.load_state_dict(WO.random_state_dict(cfg, seed), strict=True)
    return m


def main():
    import torchaudio
    from transformers import WhisperFeatureExtractor

    warnings.filterwarnings("ignore")
    out = {}
    g = torch.Generator().manual_seed(2024)
    # ---- resampler: ragged lengths (not multiples of 3), both directions of the pair the reference uses
    x = torch.randn(3, 4001, generator=g) * 0.2
    out["resample"] = {"x": x, "y_24k_16k": torchaudio.transforms.Resample(24000, 16000)(x), "y_16k_24k": torchaudio.transforms.Resample(16000, 24000)(x[:, :1000].contiguous())}
    # ---- log-mel: a noise clip and a tonal clip (most mel bins on the max - 8 floor), shorter than 30 s -> zero padded by the extractor
    n = 52000
    wav = torch.randn(2, n, generator=g) * 0.1
    wav[1] = 0.4 * torch.sin(torch.arange(n) * 0.07) + 0.2 * torch.sin(torch.arange(n) * 0.9) + 1e-3 * wav[1]
    feats = WhisperFeatureExtractor()(wav.numpy(), sampling_rate=16000, return_tensors="pt")["input_features"]
    assert feats.shape == (2, 80, 3000)
    out["logmel"] = {"wav16": wav, "frames": torch.arange(0, 3000, 7), "features_strided": feats[:, :, ::7].contiguous(),
                     "features_head": feats[:, :, :340].contiguous()}
    # ---- WavLM
    for name, cfg, seed, L in (("small", SMALL, 11, 4000), ("mid", MID, 12, 6000)):
        m = build_hf_wavlm(cfg, seed)
        wav16 = torch.randn(2, L, generator=g) * 0.3
        with torch.no_grad():
            hs = m(wav16, output_hidden_states=True).hidden_states
        out["wavlm_" + name] = {"cfg": cfg, "seed": seed, "wav16": wav16, "hidden_states": [h.clone() for h in hs]}
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "frontend_golden.pt")
    torch.save(out, path)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
