"""Tokenize front-ends at the reference's batch (reason_tokenizer.py:86 batch_size = 6 windows of 30 s + 240 samples at 24 kHz), random
weights: (1) get_whisper_features on the device (resampler + log-mel kernels) next to the reference's route (torchaudio resample on
the device, D2H, WhisperFeatureExtractor on the host, H2D); (2) the WavLM base-plus encoder up to hidden state 9 (the mean of hidden
states 6..9 that AudioDiffusion1D.get_wavlm_feature takes), resampling and the 160 appended zeros included."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film import frontend as FE  # noqa: E402
from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.modeling_wavlm import WavLMConfig, WavLMModel  # noqa: E402


def timed(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    audio = (torch.randn(B, 720240) * 0.1).to(dev)
    rs, lm = FE.Resample(24000, 16000), FE.WhisperLogMel()

    def ours():
        return lm(rs(audio, pad_to=lm.n_samples))["input_features"]

    ms, feats = timed(ours)
    ms_rs, _ = timed(lambda: rs(audio, pad_to=lm.n_samples))
    line = {"what": "get_whisper_features on the device, %d x 30 s" % B, "ms": round(ms, 3), "resampler_ms": round(ms_rs, 3), "launches": 3,
            "x_realtime": round(B * 30.0 / (ms * 1e-3), 1),
            "logmel_GFMA_f64": round(B * 3000 * 201 * 400 * 2 / 1e9, 2)}
    try:  # the reference's route, for scale (host STFT between two copies)
        import torchaudio
        from transformers import WhisperFeatureExtractor

        t16 = torchaudio.transforms.Resample(24000, 16000).to(dev)
        fe = WhisperFeatureExtractor()

        def ref():
            return fe(t16(audio).detach().cpu().numpy(), sampling_rate=16000, return_tensors="pt")["input_features"].to(dev)

        for _ in range(2):
            r = ref()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            r = ref()
        torch.cuda.synchronize()
        line["reference_route_ms"] = round((time.perf_counter() - t0) / 3 * 1e3, 2)
        line["max_abs_vs_reference_route"] = float((r - feats).abs().max())
    except Exception as e:  # noqa: BLE001
        line["reference_route_ms"] = None
        line["reference_route_error"] = repr(e)[:200]
    print(json.dumps(line))

    m = WavLMModel(WavLMConfig(), device=dev)
    m.MAX_BATCH = B
    c = m.config
    w16 = FE.Resample(24000, 16000)

    def wavlm():
        x = w16(audio, pad_to=w16.out_length(audio.shape[-1]) + 160)
        return m.hidden_states_mean(x, 6, 10)

    res = {}
    for mode, bf in (("bf16", 1), ("fp32 class", 0)):
        m.set_option("bf16", bf)
        ms, out = timed(wavlm, n=3, warm=2)
        res[mode] = (ms, out, m.last_launch_count())
    ms16, out16, n16 = res["bf16"]
    T = out.shape[1]
    t = 480160
    conv = 0.0
    ts = []
    for k, s in zip(c.conv_kernel, c.conv_stride):
        t = (t - k) // s + 1
        ts.append(t)
    for i in range(1, len(ts)):
        conv += 2.0 * ts[i] * c.conv_dim[i] * c.conv_kernel[i] * c.conv_dim[i - 1]
    D, Fi = c.hidden_size, c.intermediate_size
    pos = 2.0 * T * D * (D // c.num_conv_pos_embedding_groups) * c.num_conv_pos_embeddings
    layers = 9 * (2.0 * T * (4 * D * D + 2 * D * Fi) + 4.0 * T * T * D)
    flop = B * (conv + pos + 2.0 * T * D * c.conv_dim[-1] + layers)
    print(json.dumps({"what": "WavLM base-plus, %d x 30 s, hidden states 6..9 (9 layers), bf16 mode" % B, "ms": round(ms16, 3), "launches": n16,
                      "TFLOPs": round(flop / ms16 / 1e9, 1), "x_realtime": round(B * 30.0 / (ms16 * 1e-3), 1), "finite": bool(torch.isfinite(out16).all()),
                      "max_abs_vs_fp32_class": float((out16 - out).abs().max())}))
    print(json.dumps({"what": "WavLM base-plus, %d x 30 s, hidden states 6..9 (9 layers), fp32 class" % B, "ms": round(ms, 3), "frames": T,
                      "launches": m.last_launch_count(), "algorithmic_TFLOP": round(flop / 1e12, 3), "TFLOPs_fp32_equivalent": round(flop / ms / 1e9, 1),
                      "x_realtime": round(B * 30.0 / (ms * 1e-3), 1), "finite": bool(torch.isfinite(out).all()), "out_scale": float(out.abs().max())}))


if __name__ == "__main__":
    main()
