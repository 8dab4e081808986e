"""Profiling driver (run under ncu with --profile-from-start off): one get_whisper_features call and one WavLM base-plus forward
(hidden states 6..9) at the reference's batch of 6 windows of 30 s, after a warm-up outside the profiled region.  Not a benchmark.

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_frontend/launches.csv \
        python tools/profile_frontend.py 6
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film import frontend as FE  # noqa: E402
from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.modeling_wavlm import WavLMConfig, WavLMModel  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 6
dev = torch.device("cuda", 0)
torch.manual_seed(0)
audio = (torch.randn(B, 720240) * 0.1).to(dev)
rs, lm = FE.Resample(24000, 16000), FE.WhisperLogMel()
m = WavLMModel(WavLMConfig(), device=dev)
m.MAX_BATCH = B
if len(sys.argv) > 2 and sys.argv[2] == "bf16":
    m.set_option("bf16", 1)


def step():
    feats = lm(rs(audio, pad_to=lm.n_samples))["input_features"]
    x = rs(audio, pad_to=rs.out_length(audio.shape[-1]) + 160)
    return feats, m.hidden_states_mean(x, 6, 10)


step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
