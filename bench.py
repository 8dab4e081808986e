#!/usr/bin/env python
"""bench.py - headline benchmark of the UniAudio2 AR-decode hot path on B200.

metric   : audio tokens/s of the autoregressive decode (BASELINE.json), TTS-10 s workload (configs[1]):
           40-position text prompt prefill + fixed 179-frame schedule (52 reason-phase frames, forbid_prefix=0;
           127 semantic-phase frames, forbid_prefix=4100) = 1432 audio tokens per utterance,
           full-size model (Llama-3.2-3B backbone + 3 L / 2 L experts + Llama-3.2-300M local decoder, 4.86 B params),
           fp32 like the reference (multi_task_inference.py:181-183), random-init weights, synthetic tokens.
step     : one whole utterance (reset_caches + forward_prefix + 179 x generate_frame) per GPU.
value    : whole-job audio tokens/s, inputs resident in HBM, no host sync inside the timed region.
e2e      : same metric through the public API (evaluation.tts_task.Generator.generate_tts) with HOST buffers:
           prompt H2D from pinned memory and one D2H of the sampled frame per AR step inside the timed region.
N > 1    : one replica per GPU (torchrun), one utterance per rank per step (weak scaling), a single NCCL
           all_gather of the generated tokens at the end of each step; time = max over ranks.

  python bench.py --gpus 1 --steps 3 --warmup 3
  python bench.py --impl reference ...      # the reference's algorithm on the host CPU (oracle port, torch CPU fp32)
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "audio_tokens_per_s_ar_decode_tts10s"
UNIT = "audio tokens/s"
PROMPT_LEN = 40
N_REASON, N_SEMANTIC = 52, 127
N_FRAMES = N_REASON + N_SEMANTIC
NQ = 8
REASON_CARD, SEMANTIC_CARD = 4100, 8200
TEMPERATURE, TOPK = 0.9, 50
# parameters touched once per frame / per local step
P_GLOB, P_LOC = 3.715e9, 2.748e8


def frame_bytes(S):
    """Algorithmic HBM bytes of one generate_frame at context S (fp32): all weights once + KV read + logits."""
    kv = S * 33 * 2 * 8 * 128 * 4
    return 4.0 * (P_GLOB + NQ * P_LOC) + kv


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.idx = gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 8]
        sm, reasons, mx = [], set(), None
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        os.unlink(self.f.name)
        return out


# ---------------------------------------------------------------------------------------------- model construction
def model_args():
    from uniaudio2_b200.llm_models.model_new import ModelArgs

    return ModelArgs(llm_name="Llama-3.2-3B", decoder_name="Llama-3.2-300M", llm_pretrained_model="", audio_embeddings_path="",
                     audio_understanding_expert_path="", audio_semantic_vocab_size=SEMANTIC_CARD,
                     audio_reason_vocab_size=REASON_CARD, audio_num_codebooks=NQ)


def init_weights_(model, seed=0):
    """Seeded synthetic weights directly on the device (no checkpoints offline): Linear ~ U(+-1/sqrt(fan_in)),
    embeddings ~ N(0,1), norms ~ 1 + 0.1 N(0,1), audio_head ~ N(0, 0.02^2)."""
    dev = next(model.parameters()).device
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    for name, p in model.named_parameters():
        if "norm_" in name or "ln_f" in name:
            p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g, device=dev))
        elif name == "audio_head":
            p.normal_(0.0, 0.02, generator=g)
        elif name.endswith("wte.weight") or name == "audio_embeddings.weight":
            p.normal_(0.0, 1.0, generator=g)
        else:
            p.uniform_(-1.0, 1.0, generator=g).mul_(1.0 / math.sqrt(p.shape[-1]))


def synthetic_prompt(rank, seed=888):
    g = torch.Generator().manual_seed(seed + rank)
    task_prompt = torch.randint(0, 128000, (10,), generator=g)
    text = torch.randint(0, 128000, (PROMPT_LEN - 10 - 2,), generator=g)  # + <transcription> </transcription> = 40 rows
    return task_prompt, text


# ---------------------------------------------------------------------------------------------- GPU arm
def run_utterance_device(model, tokens, mask, pos):
    """One step with everything resident on the device and no host sync: prefill + 179 frames."""
    S = tokens.size(1)
    model.reset_caches()
    model.forward_prefix(tokens[:, :-1], labels=None, tokens_mask=mask, loss_mask=None, input_pos=pos[:, :-1], input_pos_maxp1=S - 1)
    curr_tokens, curr_mask = tokens[:, -1:], mask[:, -1:]
    audio_mask = torch.cat([torch.ones(1, 1, NQ, dtype=torch.bool), torch.zeros(1, 1, 1, dtype=torch.bool)], -1).to(tokens.device)
    frames = []
    launches = model.last_launch_count()
    for f in range(N_FRAMES):
        forbid = 0 if f < N_REASON else REASON_CARD
        s = model.generate_frame(curr_tokens, curr_mask, input_pos=S - 1 + f, input_pos_maxp1=S + f, temperature=TEMPERATURE,
                                 topk=TOPK, forbid_prefix=forbid)
        launches += model.last_launch_count()
        frames.append(s)
        sl = s.long()
        curr_tokens = torch.cat([sl[:, 1:], sl[:, 0:1]], dim=-1).unsqueeze(1)
        curr_mask = audio_mask
    return torch.stack(frames), launches


# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu --set full capture (per launch)
NCU_TRAFFIC_BYTES = 205.39e6


def time_dominant_kernel(model, hbm_peak, reps=3):
    """Roofline of the dominant kernel: the fused RMSNorm -> fc_1|fc_2 -> SiLU*mul skinny linear (gemv_kernel<1,RMSNORM,
    SWIGLU>) of the backbone MLP, 2 x 8192 x 3072 fp32 weights = 201.3 MB algorithmic bytes per launch; 28+3+2 such
    launches per frame = 28% of all frame bytes.  Timed alone with CUDA events on the launching stream, cycling through
    the 28 backbone layers' weights (5.6 GB >> 126 MB L2, so every launch streams from HBM)."""
    from uniaudio2_b200 import _lib

    L = _lib.lib()
    dev = next(model.parameters()).device
    blocks = model.backbone.transformer.h
    D, Fi = model.backbone.config.n_embd, model.backbone.config.intermediate_size
    x = torch.randn(1, D, device=dev)
    y = torch.empty(1, Fi, device=dev)
    st = _lib.current_stream()

    def one_pass():
        for b in blocks:
            _lib.check(L.ua2_swiglu_f32(_lib.ptr(x), _lib.ptr(b.mlp.fc_1.weight), _lib.ptr(b.mlp.fc_2.weight), _lib.ptr(b.norm_2.weight),
                                        1e-5, _lib.ptr(y), 1, Fi, D, st))

    for _ in range(3):
        one_pass()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        one_pass()
    e1.record()
    torch.cuda.synchronize()
    n = reps * len(blocks)
    us = e0.elapsed_time(e1) * 1e3 / n
    bytes_per_launch = 2.0 * Fi * D * 4 + (D + Fi + D) * 4
    achieved = bytes_per_launch / (us * 1e-6) / 1e9
    return {"bound": "hbm", "kernel": "gemv3_kernel<1,PRO_RMSNORM,EPI_SWIGLU> (backbone mlp fc_1|fc_2, N=8192 K=3072; same template serves every linear of the frame)",
            "achieved": round(achieved, 1), "peak": hbm_peak, "unit": "GB/s", "frac": round(achieved / hbm_peak, 4),
            "traffic": NCU_TRAFFIC_BYTES, "traffic_source": "profiles/r1_ncu_full_gemv3.md (ncu --set full, dram read+write per launch)",
            "launch_us": round(us, 2), "bytes_per_launch": bytes_per_launch, "launches_timed": n}


def host_threads():
    """Host threads the CPU arm may use: the cores this process is allowed on, NOT torch's default - torchrun exports
    OMP_NUM_THREADS=1 to its workers, which made the round-1 reference arm run on one core for N > 1."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def oracle_greedy_frames(orc, tokens, mask, pos, n_frames):
    """Full-size parity reference (BASELINE.md section 3 "parity alongside timing"): greedy (topk = 1) prefill + n_frames frames on the
    CPU oracle; returns the sampled ids (n_frames, 9) and the smallest top-1 / top-2 logit margin of every sampled head."""
    S = tokens.size(1)
    audio_mask = torch.cat([torch.ones(1, 1, NQ, dtype=torch.bool), torch.zeros(1, 1, 1, dtype=torch.bool)], -1)
    frames, margin = [], float("inf")
    with torch.inference_mode():
        orc.reset_caches()
        orc.forward_prefix(tokens[:, :-1], mask, pos[:, :-1])
        curr_tokens, curr_mask = tokens[:, -1:], mask[:, -1:]
        for f in range(n_frames):
            dbg = {}
            s = orc.generate_frame(curr_tokens, curr_mask, torch.tensor([S - 1 + f]), S + f, 1.0, 1, 0, debug=dbg)
            for lg in [dbg["text_logits"]] + list(dbg["ci_logits"]):
                t2 = lg.float().topk(2, dim=-1)[0]
                margin = min(margin, float((t2[..., 0] - t2[..., 1]).min()))
            frames.append(s[0].clone())
            sl = s.long()
            curr_tokens = torch.cat([sl[:, 1:], sl[:, 0:1]], dim=-1).unsqueeze(1)
            curr_mask = audio_mask
    return torch.stack(frames), margin


def gpu_greedy_frames_teacher_forced(model, tokens, mask, pos, ref_frames):
    """The same greedy frames on the GPU model; every frame is fed the ORACLE's previous ids so that each frame compares on its own."""
    dev = next(model.parameters()).device
    S = tokens.size(1)
    tok, msk, ps = tokens.to(dev), mask.to(dev), pos.to(dev)
    audio_mask = torch.cat([torch.ones(1, 1, NQ, dtype=torch.bool), torch.zeros(1, 1, 1, dtype=torch.bool)], -1).to(dev)
    model.reset_caches()
    model.forward_prefix(tok[:, :-1], labels=None, tokens_mask=msk, loss_mask=None, input_pos=ps[:, :-1], input_pos_maxp1=S - 1)
    curr_tokens, curr_mask = tok[:, -1:], msk[:, -1:]
    out = []
    for f in range(ref_frames.size(0)):
        s = model.generate_frame(curr_tokens, curr_mask, input_pos=S - 1 + f, input_pos_maxp1=S + f, temperature=1.0, topk=1, forbid_prefix=0)
        out.append(s[0].cpu())
        sl = ref_frames[f:f + 1].long().to(dev)
        curr_tokens = torch.cat([sl[:, 1:], sl[:, 0:1]], dim=-1).unsqueeze(1)
        curr_mask = audio_mask
    return torch.stack(out)


def cpu_baseline_sample(state_dict_cpu, n_frames=8, threads=None, parity_frames=0):
    """The reference's algorithm (oracle port: torch CPU fp32, same ATen ops as the reference) on the host cores.
    Bounded sample: one 39-position prefill + n_frames generate_frame calls at the start of the TTS-10 s schedule;
    the utterance time is extrapolated as prefill + 179 x mean frame time."""
    from oracle import llm_oracle as O

    torch.set_num_threads(threads if threads else host_threads())
    cfg = O.full_size_cfg(REASON_CARD, SEMANTIC_CARD)
    orc = O.Stage3Oracle(cfg, state_dict_cpu)
    orc.setup_caches(1)
    task_prompt, text = synthetic_prompt(0)
    seq = torch.cat([task_prompt, torch.tensor([128011]), text, torch.tensor([128012])])
    S = seq.numel()
    tokens = torch.zeros(1, S, NQ + 1, dtype=torch.long)
    tokens[0, :, -1] = seq
    mask = torch.zeros(1, S, NQ + 1, dtype=torch.bool)
    mask[..., -1] = True
    pos = torch.arange(S).unsqueeze(0)
    torch.manual_seed(888)
    with torch.inference_mode():
        orc.reset_caches()
        t0 = time.perf_counter()
        orc.forward_prefix(tokens[:, :-1], mask, pos[:, :-1])
        t_prefill = time.perf_counter() - t0
        curr_tokens, curr_mask = tokens[:, -1:], mask[:, -1:]
        audio_mask = torch.cat([torch.ones(1, 1, NQ, dtype=torch.bool), torch.zeros(1, 1, 1, dtype=torch.bool)], -1)
        f = 0

        def one_frame():
            nonlocal curr_tokens, curr_mask, f
            t0 = time.perf_counter()
            s = orc.generate_frame(curr_tokens, curr_mask, torch.tensor([S - 1 + f]), S + f, TEMPERATURE, TOPK, 0)
            dt = time.perf_counter() - t0
            sl = s.long()
            curr_tokens = torch.cat([sl[:, 1:], sl[:, 0:1]], dim=-1).unsqueeze(1)
            curr_mask = audio_mask
            f += 1
            return dt

        times = [one_frame()]  # warm-up
        # the frame is a memory-bound GEMV chain: more threads than memory channels can be slower, so give the CPU arm the
        # thread count it runs fastest with (one frame per candidate), then time n_frames with it
        tried = {}
        if threads is None:
            full = torch.get_num_threads()
            for n in sorted({full, max(1, full // 2), max(1, full // 4), min(full, 16), min(full, 8)}, reverse=True):
                torch.set_num_threads(n)
                one_frame()
                tried[n] = round(one_frame() * 1e3, 1)
            torch.set_num_threads(min(tried, key=tried.get))
            # re-time the prefill with the chosen thread count (the first one ran before the calibration)
            orc.reset_caches()
            t0 = time.perf_counter()
            orc.forward_prefix(tokens[:, :-1], mask, pos[:, :-1])
            t_prefill = time.perf_counter() - t0
            curr_tokens, curr_mask, f = tokens[:, -1:], mask[:, -1:], 0
            one_frame()
        for _ in range(n_frames):
            times.append(one_frame())
    frame_t = sum(times[1:]) / len(times[1:])  # first frame = warm-up
    est = t_prefill + N_FRAMES * frame_t
    if parity_frames:
        ref_frames, margin = oracle_greedy_frames(orc, tokens, mask, pos, parity_frames)
        cpu_baseline_sample.parity_ref = dict(tokens=tokens, mask=mask, pos=pos, frames=ref_frames, min_margin=margin)
    return {"value": round(NQ * N_FRAMES / est, 2), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"1 prefill({S - 1} pos, {t_prefill:.2f}s) + {n_frames} frames ({frame_t * 1e3:.0f} ms/frame, 1 warm-up frame discarded); "
                      f"utterance extrapolated to prefill + {N_FRAMES} frames = {est:.1f}s", "frame_ms": round(frame_t * 1e3, 1),
            "prefill_s": round(t_prefill, 3), "threads_tried_ms_per_frame": tried}, est


MIMI_GEOMETRY = dict(n_filters=64, encoder_rates=[8, 6, 5, 4], latent_dim=512, codebook_size=2048, codebook_dim=256, rvq_layers=32,
                     num_heads=8, num_layers=8, layer_scale=0.01, context=250)  # mimi_config.yaml + MimiCodec.py:47-58 defaults


def init_codec_weights_(m, seed=7):
    """Seeded synthetic codec weights written straight into the product module (no checkpoints offline): conv / linear
    ~ U(+-1/sqrt(fan_in)), biases ~ 0.1 N(0,1), norms ~ 1 + 0.1 N(0,1), layer scales 0.3-0.4 (large enough for the
    transformer to matter), codebook sums ~ N(0,1) with usage in [0.5, 2]."""
    g = torch.Generator().manual_seed(seed)
    sd = m.state_dict()
    for name in sorted(sd.keys()):
        t = sd[name]
        if not torch.is_floating_point(t) or name.endswith("_initialized"):
            continue
        shape = tuple(t.shape)
        if name.endswith("cluster_usage"):
            v = 0.5 + 1.5 * torch.rand(shape, generator=g)
        elif name.endswith("embedding_sum"):
            v = torch.randn(shape, generator=g)
        elif name.endswith("layer_scale_1.scale") or name.endswith("layer_scale_2.scale"):
            v = 0.3 + 0.1 * torch.rand(shape, generator=g)
        elif ".norm" in name and name.endswith(".weight"):
            v = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            v = 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = math.prod(shape[1:]) if len(shape) > 1 else shape[0]
            if "convtr" in name and len(shape) == 3:
                fan_in = shape[0] * shape[2]
            v = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(max(1.0, fan_in))
        sd[name] = v.float().to(t.device)
    m.load_state_dict(sd, strict=True)


def make_codec(dev):
    """The product's MimiCodec drop-in in the mimi_config.yaml geometry with synthetic weights (shared by the bench and the tools)."""
    from uniaudio2_b200.tools.tokenizer.MimiCodec.mimi_codec import MimiCodec

    m = MimiCodec(device=dev, **MIMI_GEOMETRY)
    init_codec_weights_(m)
    return m


def codec_conv_roofline(dev, batch=16, clip_s=10.0, reps=5):
    """Roofline of the dominant SEANet kernel of an encode + decode at this size: the strided down-sampling convolution 64 -> 128
    (k 8, s 4) at the 24 kHz rate, conv_umma_kernel<128> (csrc/ua2_convumma.cu: implicit GEMM on tcgen05 straight from (B, C, T)).
    HBM-bound by its arithmetic (algorithmic bytes = input + output activations, 4 B each; 85 FLOP/B against a 3xTF32 ridge of ~35)."""
    from uniaudio2_b200 import _lib

    L, P = _lib.lib(), _lib.ptr
    T = int(clip_s * 24000)
    x = torch.randn(batch, 64, T, device=dev)
    w = torch.randn(128, 64, 8, device=dev) / (64 * 8) ** 0.5
    b = torch.zeros(128, device=dev)
    y = torch.empty(batch, 128, T // 4, device=dev)
    st = _lib.current_stream()

    def run():
        _lib.check(L.ua2_conv1d_causal_gemm_f32(P(x), P(w), P(b), None, P(y), batch, 64, 128, T, 8, 4, 1, 1, 0, st))

    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    by = 4.0 * (x.numel() + y.numel())
    flop = 2.0 * 64 * 8 * 128 * batch * (T // 4)
    hbm_peak, hbm_src = load_peaks()
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    tf32_peak = float(pk.get("bf16_tflops_sustained", 0.0)) / 2.0  # tf32 runs at half the bf16 rate
    if tf32_peak <= 0.0:
        tf32_peak = 1125.0  # nominal dense tf32, B200_PROFILING.md
    mma = 3.0 * flop / ms / 1e9  # tf32 tensor-core TFLOP/s issued: 3 MMAs per fp32 product (operand splitting keeps fp32-class accuracy)
    # report max(bytes / BW, flops / peak).  85 FLOP per byte: 3 * flop / tf32 peak = 0.55 ms against bytes / HBM peak =
    # 0.23 ms, so the arithmetic binds - on the fp32 SIMT pipe (74 TFLOP/s measured) the same layer cannot run below 1.69 ms.
    return {"bound": "tensor", "kernel": "conv_umma_kernel<128> + weight repack (encoder down-sampling conv 64 -> 128, k 8 s 4, 24 kHz)",
            "achieved": round(mma, 1), "peak": round(tf32_peak, 1), "unit": "TFLOP/s", "frac": round(mma / tf32_peak, 4), "traffic": None,
            "launch_ms": round(ms, 3), "flop_per_launch": flop, "fp32_equivalent_tflops": round(flop / ms / 1e9, 1),
            "bytes_per_launch": by, "hbm_gbs": round(by / ms / 1e6, 1), "hbm_frac": round(by / ms / 1e6 / hbm_peak, 4),
            "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained / 2 (tf32); HBM: " + hbm_src,
            "note": "tf32 MMAs issued (3 per fp32 product); the layer moves 983 MB in + 492 MB out per launch; limited by the producer warps "
                    "that gather, activate and split the activations into tensor memory (profiles/r2_kernel_rooflines.md)"}


def codec_sweep(dev, rank, world, clips=(1.0, 5.0, 30.0), batches=(1, 4, 16, 64, 256), budget_samples=16 * 240000):
    """BASELINE.json config 5: encode + decode RTF over clip length x batch.  The `batch` clips of a point are dealt to the ranks (clip i
    -> rank i mod W, strong scaling); a rank runs its share in sub-batches of at most `budget_samples` samples.  Returns this rank's
    milliseconds per point (the caller takes the max over ranks)."""
    m = make_codec(dev)
    out = []
    for clip_s in clips:
        T = int(clip_s * 24000)
        for batch in batches:
            mine = len(range(rank, batch, world))
            sub = max(1, min(mine, budget_samples // T)) if mine else 0
            wav = torch.randn(max(sub, 1), 1, T, device=dev) * 0.1
            ms = 0.0
            if mine:
                m.decode(m.encode(wav[:sub]))  # warm-up of this shape
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                done = 0
                while done < mine:
                    n = min(sub, mine - done)
                    m.decode(m.encode(wav[:n]))
                    done += n
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
            out.append((clip_s, batch, ms))
    del m
    torch.cuda.empty_cache()
    return out


def bench_codec(dev, batch=16, clip_s=10.0, reps=3, cpu=True, roofline=True):
    """Codec real-time factor (BASELINE.json 'codec RTF'): SEANet + 8-layer transformer + 32 x 2048 x 256 residual VQ
    (mimi_config.yaml geometry, the in-repo twin of llm_modules/{seanet,conv,resample,transformer}.py), encode + decode of
    `batch` synthetic clips of `clip_s` seconds at 24 kHz, fp32, random weights.  RTF = audio seconds / wall seconds."""
    m = make_codec(dev)
    T = int(clip_s * 24000)
    g = torch.Generator().manual_seed(0)
    wav_h = (torch.randn(batch, 1, T, generator=g) * 0.1).pin_memory()
    wav = wav_h.to(dev)
    for _ in range(2):
        codes = m.encode(wav)
        out = m.decode(codes)
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    enc_ms = dec_ms = 0.0
    for _ in range(reps):
        e0.record()
        codes = m.encode(wav)
        e1.record()
        out = m.decode(codes)
        e2.record()
        torch.cuda.synchronize()
        enc_ms += e0.elapsed_time(e1)
        dec_ms += e1.elapsed_time(e2)
    enc_ms /= reps
    dec_ms /= reps
    # end to end with host buffers: H2D of the waveform, D2H of codes and of the reconstruction
    t0 = time.perf_counter()
    for _ in range(reps):
        c = m.encode(wav_h.to(dev, non_blocking=True))
        ch = c.cpu()
        oh = m.decode(ch.to(dev)).cpu()
    e2e_s = (time.perf_counter() - t0) / reps
    audio_s = batch * clip_s
    res = {"config": f"Mimi-twin codec, batch {batch} x {clip_s:.0f} s @24 kHz, fp32", "encode_ms": round(enc_ms, 2), "decode_ms": round(dec_ms, 2),
           "rtf_x_realtime": round(audio_s / ((enc_ms + dec_ms) * 1e-3), 1), "encode_x_realtime": round(audio_s / (enc_ms * 1e-3), 1),
           "decode_x_realtime": round(audio_s / (dec_ms * 1e-3), 1), "e2e_x_realtime": round(audio_s / e2e_s, 1),
           "codes_shape": list(codes.shape), "gflop_per_audio_s": 11.0}
    if roofline:
        res["roofline"] = codec_conv_roofline(dev, batch, clip_s)
    if cpu:
        from oracle import codec_oracle as CO  # the CPU-baseline leg: the oracle port runs the same weights on the host cores

        cfg = CO.MimiCfg()
        shapes = CO.mimi_param_shapes(cfg)
        sd = {k: v.detach().cpu() for k, v in m.state_dict().items() if k in shapes}
        orc = CO.MimiOracle(cfg, sd)
        w1 = wav_h[:1, :, : 2 * 24000].clone()
        with torch.inference_mode():
            orc.decode(orc.encode(w1[..., :12000]))  # warm-up
            t0 = time.perf_counter()
            cc = orc.encode(w1)
            t1 = time.perf_counter()
            orc.decode(cc)
            t2 = time.perf_counter()
        res["cpu_baseline"] = {"kind": "port", "cores": torch.get_num_threads(), "sample": "1 clip x 2 s encode+decode",
                               "rtf_x_realtime": round(2.0 / (t2 - t0), 2), "encode_s": round(t1 - t0, 3), "decode_s": round(t2 - t1, 3)}
    del m
    torch.cuda.empty_cache()
    return res


DIT_PROD = dict(num_attention_heads=24, attention_head_dim=64, in_channels=1040, out_channels=136, num_layers=32, attention_bias=True,
                activation_fn="gelu-approximate", norm_type="ada_norm_single", norm_elementwise_affine=False, norm_eps=1e-6,
                num_embeds_ada_norm=1000)  # tools/tokenizer/ReasoningCodec_film/models/model_config.json


def bench_flow_decoder(dev, frames=500, steps=10, reps=3, cpu=True):
    """Tokens -> latent of `--stage all`: the flow-matching decoder of ReasoningCodec_film on one
    20 s window - Euler solver, `steps` steps (test.sh: 10), classifier-free guidance 1.5 (reason_tokenizer.py:273), DiT of
    model_config.json with random weights.  1.93 TFLOP per estimator call on the 2 x 500-row CFG batch."""
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.AudioDiffusion1D import BASECFM
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.transformer_1d_flow import Transformer1DModel

    torch.manual_seed(0)
    m = Transformer1DModel(device=dev, **DIT_PROD)
    cfm = BASECFM(m)
    T, D, I, O, L = frames, 1536, 1040, 136, 32
    flop = L * 24 * D * D * 2 * T + L * 4 * T * T * D * 2 + 2 * 2 * T * (3 * I * D + D * D + 3 * D * O + O * O)
    g = torch.Generator().manual_seed(1)
    z_h, mu_h = torch.randn(1, T, O, generator=g).pin_memory(), torch.randn(1, T, I - 2 * O, generator=g).pin_memory()
    ic = torch.zeros(1, T, O, device=dev)
    t_span = torch.linspace(0, 1, steps + 1)
    z, mu = z_h.to(dev), mu_h.to(dev)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    bf16_peak = float(peaks.get("bf16_tflops_sustained", 0.0))
    audio_s = T / 25.0
    modes = {}
    # bf16 = the reference's own arithmetic for this block (torch.autocast(bfloat16), reason_tokenizer.py:265) and the default of the
    # product's ReasoningTokenizer; fp32_class = 3xTF32 (what the 1e-4 parity tests run)
    for mode, bf16 in (("bf16", 1), ("fp32_class", 0)):
        m.set_option("bf16", bf16)
        for _ in range(2):
            cfm.solve_euler(z, ic, 0, t_span, mu, None, 1.5)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = cfm.solve_euler(z, ic, 0, t_span, mu, None, 1.5)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        t0 = time.perf_counter()
        for _ in range(reps):  # end to end with host buffers: H2D of noise and condition, D2H of the latent
            lat_h = cfm.solve_euler(z_h.to(dev, non_blocking=True), ic, 0, t_span, mu_h.to(dev, non_blocking=True), None, 1.5).cpu()
        e2e_s = (time.perf_counter() - t0) / reps
        mma = (1 if bf16 else 3) * flop * steps / ms / 1e9  # tensor-core TFLOP/s actually issued (3 tf32 MMAs per fp32 product)
        peak = bf16_peak / (1 if bf16 else 2)
        modes[mode] = {"solve_ms": round(ms, 1), "x_realtime": round(audio_s / (ms * 1e-3), 1), "e2e_x_realtime": round(audio_s / e2e_s, 1),
                       "estimator_ms": round(ms / steps, 2), "mma_tflops": round(mma, 1),
                       "roofline": {"bound": "tensor", "achieved": round(mma, 1), "peak": round(peak, 1) if peak else None, "unit": "TFLOP/s",
                                    "frac": round(mma / peak, 4) if peak else None,
                                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" + ("" if bf16 else " / 2 (tf32 runs at half the bf16 rate)"),
                                    "note": "whole estimator call: linears on the hand-written tcgen05 mainloop + fp32 SIMT attention + glue kernels"}}
    res = {"config": f"DiT 32 x (24 x 64), CFG batch 2 x {T} frames (20 s window), {steps} Euler steps", "tflop_per_estimator_call": round(flop / 1e12, 3),
           "default_mode": "bf16 (the reference's autocast)", **modes["bf16"], "fp32_class": modes["fp32_class"], "latent_shape": list(lat_h.shape)}
    if cpu:
        from oracle import dit_oracle as DO  # the CPU-baseline leg: the oracle port runs one estimator call on the host cores

        sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
        orc = DO.DitOracle(DO.DitCfg(), sd)
        x = torch.randn(2, T, I, generator=g)
        tt = torch.full((2,), 0.35)
        with torch.inference_mode():
            t0 = time.perf_counter()
            orc.forward(x, tt)
            dt = time.perf_counter() - t0
        res["cpu_baseline"] = {"kind": "port", "cores": torch.get_num_threads(), "sample": "1 estimator call (of the 10 a solve makes)",
                               "estimator_s": round(dt, 2), "x_realtime": round(audio_s / (dt * steps), 2)}
    del cfm, m
    torch.cuda.empty_cache()
    return res


def bench_whisper_encoder(dev, batch=6, reps=3, cpu=True):
    """First SSL front-end of ReasoningCodec_film's tokenize: the Whisper-medium encoder on `batch`
    windows of 30 s (reason_tokenizer.py:86 batch_size=6), random weights, bf16 mode = the reference's autocast arithmetic
    (reason_tokenizer.py:114-118) and fp32 class.  1.14 TFLOP per window."""
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.modeling_whisper import WhisperConfig, WhisperModel

    torch.manual_seed(0)
    m = WhisperModel(WhisperConfig(), device=dev).encoder
    P, D, F, L = 1500, 1024, 4096, 24
    flop = batch * (2.0 * 3000 * 240 * D + 2.0 * P * 3 * D * D + L * (2.0 * P * (4 * D * D + 2 * D * F) + 4.0 * P * P * D))
    g = torch.Generator().manual_seed(2)
    mel_h = torch.randn(batch, 80, 3000, generator=g).pin_memory()
    mel = mel_h.to(dev)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    bf16_peak = float(peaks.get("bf16_tflops_sustained", 0.0))
    modes = {}
    for mode, bf16 in (("bf16", 1), ("fp32_class", 0)):
        m.set_option("bf16", bf16)
        for _ in range(2):
            y = m(mel).last_hidden_state
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            y = m(mel).last_hidden_state
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        t0 = time.perf_counter()
        for _ in range(reps):  # end to end with host buffers: H2D of the log-mel windows, D2H of the features
            y_h = m(mel_h.to(dev, non_blocking=True)).last_hidden_state.cpu()
        e2e_s = (time.perf_counter() - t0) / reps
        mma = (1 if bf16 else 3) * flop / ms / 1e9
        peak = bf16_peak / (1 if bf16 else 2)
        modes[mode] = {"ms": round(ms, 2), "x_realtime": round(batch * 30.0 / (ms * 1e-3), 1), "e2e_x_realtime": round(batch * 30.0 / e2e_s, 1),
                       "mma_tflops": round(mma, 1),
                       "roofline": {"bound": "tensor", "achieved": round(mma, 1), "peak": round(peak, 1) if peak else None, "unit": "TFLOP/s",
                                    "frac": round(mma / peak, 4) if peak else None,
                                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" + ("" if bf16 else " / 2 (tf32)")}}
    res = {"config": f"Whisper-medium encoder (24 x (16 x 64), d 1024), {batch} x 30 s windows, random weights",
           "tflop_per_call": round(flop / 1e12, 3), "default_mode": "bf16 (the reference's autocast)", **modes["bf16"],
           "fp32_class": modes["fp32_class"], "out_shape": list(y_h.shape)}
    if cpu:
        from oracle import whisper_oracle as WO  # the CPU-baseline leg: the oracle port encodes one window on the host cores

        sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
        orc = WO.WhisperEncoderOracle(WO.WhisperCfg(), sd)
        with torch.inference_mode():
            t0 = time.perf_counter()
            ref = orc.forward(mel_h[:1])
            dt = time.perf_counter() - t0
        m.set_option("bf16", 0)
        got = m(mel[:1]).last_hidden_state.cpu()
        res["cpu_baseline"] = {"kind": "port", "cores": torch.get_num_threads(), "sample": "1 window of 30 s", "s": round(dt, 2),
                               "x_realtime": round(30.0 / dt, 1)}
        res["parity"] = {"fp32_class_vs_cpu_oracle_max_abs": float((got - ref).abs().max()), "out_scale": float(ref.abs().max())}
    del m
    torch.cuda.empty_cache()
    return res


def bench_tokenize_frontends(dev, batch=6, reps=3, cpu=True):
    """The waveform side of ReasoningCodec_film's tokenize at the reference's batch of `batch` windows of
    30 s + 240 samples (reason_tokenizer.py:86-110): get_whisper_features (resampler + log-mel, :67-72) on the device next to the
    reference's own route for that step (device resample -> host WhisperFeatureExtractor -> device), and the WavLM base-plus encoder
    up to hidden state 9 (AudioDiffusion1D.get_wavlm_feature :359-370, resampling and the 160 appended zeros included) in bf16 mode
    (the reference's autocast) and fp32 class.  Random weights."""
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film import frontend as FE
    from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.modeling_wavlm import WavLMConfig, WavLMModel

    def timed(fn, n):
        for _ in range(2):
            out = fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, out

    g = torch.Generator().manual_seed(3)
    audio_h = (torch.randn(batch, 720240, generator=g) * 0.1).pin_memory()
    audio = audio_h.to(dev)
    rs, lm = FE.Resample(24000, 16000), FE.WhisperLogMel()
    ms_fe, feats = timed(lambda: lm(rs(audio, pad_to=lm.n_samples))["input_features"], reps)
    res = {"config": f"{batch} windows of 30 s + 240 samples at 24 kHz, random weights",
           "whisper_features": {"ms": round(ms_fe, 3), "x_realtime": round(batch * 30.0 / (ms_fe * 1e-3), 1), "launches": 3, "dtype": "f64 accumulation"}}
    if cpu:
        try:  # the reference's route for this step: third-party classes, present in this image
            import torchaudio
            from transformers import WhisperFeatureExtractor

            t16, fe = torchaudio.transforms.Resample(24000, 16000).to(dev), WhisperFeatureExtractor()

            def ref_route():
                return fe(t16(audio).detach().cpu().numpy(), sampling_rate=16000, return_tensors="pt")["input_features"].to(dev)

            ref = ref_route()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                ref = ref_route()
            torch.cuda.synchronize()
            res["whisper_features"]["reference_route"] = {"kind": "reference", "ms": round((time.perf_counter() - t0) / reps * 1e3, 2),
                                                          "what": "torchaudio Resample on the device, D2H, transformers WhisperFeatureExtractor on the host, H2D"}
            res["whisper_features"]["parity_max_abs"] = float((ref - feats).abs().max())
        except Exception as e:  # noqa: BLE001
            res["whisper_features"]["reference_route"] = {"error": f"{type(e).__name__}: {e}"}
    torch.manual_seed(0)
    m = WavLMModel(WavLMConfig(), device=dev)
    m.MAX_BATCH = batch
    c = m.config
    t, ts = 480160, []
    for k, st in zip(c.conv_kernel, c.conv_stride):
        t = (t - k) // st + 1
        ts.append(t)
    T, D, Fi = ts[-1], c.hidden_size, c.intermediate_size
    flop = sum(2.0 * ts[i] * c.conv_dim[i] * c.conv_kernel[i] * c.conv_dim[i - 1] for i in range(1, len(ts)))
    flop += 2.0 * T * D * (D // c.num_conv_pos_embedding_groups) * c.num_conv_pos_embeddings + 2.0 * T * D * c.conv_dim[-1]
    flop = batch * (flop + 9 * (2.0 * T * (4 * D * D + 2 * D * Fi) + 4.0 * T * T * D))

    def wavlm():
        return m.hidden_states_mean(rs(audio, pad_to=rs.out_length(audio.shape[-1]) + 160), 6, 10)

    modes = {}
    for mode, bf16 in (("bf16", 1), ("fp32_class", 0)):
        m.set_option("bf16", bf16)
        ms, out = timed(wavlm, reps)
        modes[mode] = {"ms": round(ms, 2), "x_realtime": round(batch * 30.0 / (ms * 1e-3), 1), "tflops": round(flop / ms / 1e9, 1), "launches": m.last_launch_count()}
    res["wavlm"] = {"config": "WavLM base-plus geometry (12 x 64 heads, d 768, 7 convolutions), hidden states 6..9 = 9 layers", "frames": T,
                    "tflop_per_call": round(flop / 1e12, 3), "default_mode": "fp32 class (bf16 = the reference's autocast, option)", **modes["bf16"],
                    "fp32_class": modes["fp32_class"]}
    if cpu:
        from oracle import wavlm_oracle as WO  # CPU-baseline leg: the oracle port (pinned to transformers.WavLMModel) on one 10 s clip

        cfg = dict(WO.BASE_PLUS, num_hidden_layers=9)
        sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
        clip = rs(audio[:1, :240000], pad_to=160160)
        with torch.inference_mode():
            t0 = time.perf_counter()
            hs = WO.hidden_states(sd, cfg, clip.cpu())
            dt = time.perf_counter() - t0
        got = m.hidden_states_mean(clip, 6, 10).cpu()
        ref = torch.stack(hs, 1)[:, 6:10].mean(1)
        res["wavlm"]["cpu_baseline"] = {"kind": "port", "cores": torch.get_num_threads(), "sample": "1 clip of 10 s", "s": round(dt, 2),
                                        "x_realtime": round(10.0 / dt, 1)}
        res["wavlm"]["parity"] = {"fp32_class_vs_cpu_oracle_max_abs": float((got - ref).abs().max()), "out_scale": float(ref.abs().max())}
    del m
    torch.cuda.empty_cache()
    try:  # reasoning encoder (AudioThinking, AudioDiffusion1D.py:169-188, :372-390): 1500 Whisper + 750 BEST-RQ frames per window -> 150 query tokens
        from uniaudio2_b200.tools.tokenizer.ReasoningCodec_film.models.audio_thinking import AudioThinking

        at = AudioThinking(device=dev)
        with torch.no_grad():
            for prm in at.reasoning_vq.parameters():
                prm.normal_()
        wh, mu = torch.randn(batch, 1024, 1500, device=dev), torch.randn(batch, 1024, 750, device=dev)
        ms_at, q = timed(lambda: at.encode_reasoning_part(wh, mu)[0], reps)
        Tn, Dm, Ff = 900, 768, 3072
        fl = batch * (2.0 * 750 * 1024 * 2048 + 2.0 * 750 * 2048 * Dm + 5 * (2.0 * Tn * (4 * Dm * Dm + 3 * Dm * Ff) + 4.0 * Tn * Tn * Dm))
        res["audio_thinking"] = {"config": "AudioThinking (dim 768, 6 x 128 heads, 5 blocks) + 8-level residual VQ, fp32 class", "ms": round(ms_at, 2),
                                 "x_realtime": round(batch * 30.0 / (ms_at * 1e-3), 1), "tflops_fp32_equivalent": round(fl / ms_at / 1e9, 1),
                                 "launches": at.last_launch_count(), "out_shape": list(q.shape)}
        if cpu:
            from oracle import thinking_oracle as TO  # CPU-baseline leg: the oracle port (bit-equal to the reference source) on one window

            sd = {k: v.detach().cpu() for k, v in at.state_dict().items() if not k.startswith("reasoning_vq.")}
            with torch.inference_mode():
                t0 = time.perf_counter()
                ref = TO.encode(sd, TO.CFG, wh[:1].cpu(), mu[:1].cpu())
                dt = time.perf_counter() - t0
            got = at.query_tokens(wh[:1], mu[:1]).cpu()
            res["audio_thinking"]["cpu_baseline"] = {"kind": "port", "cores": torch.get_num_threads(), "sample": "1 window of 30 s", "s": round(dt, 2),
                                                     "x_realtime": round(30.0 / dt, 1)}
            res["audio_thinking"]["parity"] = {"query_tokens_vs_cpu_oracle_max_abs": float((got - ref).abs().max()), "out_scale": float(ref.abs().max())}
        del at
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        res["audio_thinking"] = {"error": f"{type(e).__name__}: {e}"}
    return res


def state_dict_to_cpu(model):
    return {k: v.detach().to("cpu") for k, v in model.state_dict().items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ua2", choices=["ua2", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-frames", type=int, default=8)
    ap.add_argument("--parity-frames", type=int, default=8, help="greedy frames compared id-for-id with the CPU oracle at full size (N = 1)")
    ap.add_argument("--no-codec", action="store_true")
    ap.add_argument("--no-flow-decoder", action="store_true")
    ap.add_argument("--codec-sweep", action="store_true", help="BASELINE.json config 5: codec RTF over 1/5/30 s clips x batch 1..256, clips dealt to the ranks")
    ap.add_argument("--v3-cps", type=int, default=0)
    ap.add_argument("--v3-stages", type=int, default=0)
    ap.add_argument("--v3-kcw", type=int, default=0)
    ap.add_argument("--v3-balance", type=int, default=-1)
    ap.add_argument("--v3-budget", type=int, default=0)
    ap.add_argument("--attn-direct", type=int, default=-1, help="local-decoder attention inside the proj prologue (library default 1)")
    ap.add_argument("--pf-mb", type=int, default=-1, help="tail L2 prefetch budget per linear, MB (-1: library default)")
    ap.add_argument("--pf-idle-mb", type=int, default=-1, help="extra prefetch budget before attention / sampler kernels, MB")
    ap.add_argument("--pdl", type=int, default=int(os.environ.get("UA2_PDL", "1")))
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    warmup = max(args.warmup, 3) if args.impl == "ua2" else args.warmup
    config = {"workload": "TTS-10s AR decode: 40-pos prefill + 179 frames (52 reason + 127 semantic), B=1 per GPU, "
                          "Llama-3.2-3B backbone + 3L/2L experts + 300M local decoder (4.86B params), topk=50 T=0.9",
              "frames_per_step": N_FRAMES, "audio_tokens_per_step": NQ * N_FRAMES, "prompt_len": PROMPT_LEN,
              "parallelism": f"replica x{world} (one utterance per GPU)", "l2_policy": "inputs larger than L2 (19.5 GB weights streamed per frame)"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        dev = "cuda:0" if torch.cuda.is_available() else "cpu"
        from uniaudio2_b200.llm_models.model_new import Model_stage3

        with torch.inference_mode():
            m = Model_stage3(model_args(), device=dev)
            init_weights_(m, 0)
            sd = state_dict_to_cpu(m)
        del m
        if dev != "cpu":
            torch.cuda.empty_cache()
        ests, last = [], None
        threads = None  # the first sample picks the fastest thread count, the others reuse it
        for i in range(args.warmup + args.steps):
            cb, est = cpu_baseline_sample(sd, n_frames=args.cpu_frames, threads=threads)
            threads = cb["cores"]
            if i >= args.warmup:
                ests.append(est)
                last = cb
        est = sum(ests) / len(ests)
        v = NQ * N_FRAMES / est
        last["value"] = round(v, 2)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": round(v, 2), "unit": UNIT, "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(est * 1e3, 1), "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic", "config": config,
                          "cpu_baseline": last, "e2e": {"value": round(v, 2), "unit": UNIT, "h2d_bytes_per_step": 0,
                                                        "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------------ product arm (GPU)
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback on the product path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    from uniaudio2_b200 import _lib
    from uniaudio2_b200.evaluation.tts_task import Generator, default_train_args
    from uniaudio2_b200.llm_models.model_new import Model_stage3

    if args.v3_balance >= 0:
        _lib.check(_lib.lib().ua2_set_global_option(b"gemv3_balance_grid", args.v3_balance))
    if args.v3_kcw:
        _lib.check(_lib.lib().ua2_set_global_option(b"gemv3_kcw", args.v3_kcw))
    if args.v3_budget:
        _lib.check(_lib.lib().ua2_set_global_option(b"gemv3_budget_kb", args.v3_budget))
    if args.pf_mb >= 0:
        _lib.check(_lib.lib().ua2_set_global_option(b"gemv3_prefetch_mb", args.pf_mb))
    if args.pf_idle_mb >= 0:
        _lib.check(_lib.lib().ua2_set_global_option(b"gemv3_prefetch_idle_mb", args.pf_idle_mb))
    if args.v3_stages:
        _lib.check(_lib.lib().ua2_set_global_option(b"gemv3_max_stages", args.v3_stages))
    if args.v3_cps:
        _lib.check(_lib.lib().ua2_set_global_option(b"gemv3_ctas_per_sm", args.v3_cps))
    with torch.inference_mode():
        model = Model_stage3(model_args(), device=dev)
        init_weights_(model, 0)  # same weights on every replica
        gen = Generator(model, default_train_args(REASON_CARD, SEMANTIC_CARD), is_cfg=False)  # setup_caches(1)
        model.set_option("pdl", int(args.pdl))
        if args.attn_direct >= 0:
            model.set_option("attn_direct", int(args.attn_direct))
        task_prompt, text = synthetic_prompt(rank)
        tokens, mask = gen.prepare_tts_task(task_prompt, text)
        assert tokens.size(0) == PROMPT_LEN
        tokens_d = tokens.unsqueeze(0).to(dev)
        mask_d = mask.bool().unsqueeze(0).to(dev)
        pos_d = torch.arange(PROMPT_LEN, device=dev).unsqueeze(0)
        from uniaudio2_b200.distributed import gather_variable

        def step_device():
            frames, launches = run_utterance_device(model, tokens_d, mask_d, pos_d)
            if world > 1:
                # the single exchange of the path: every rank's generated (9, 179) token matrix, gathered over NCCL
                gather_variable([frames[:, 0, :].t().contiguous()], world)
            return frames, launches

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        torch.manual_seed(888 + rank)
        for _ in range(warmup):
            step_device()
        barrier()
        clocks = ClockSampler(local_rank)
        if rank == 0:
            clocks.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches = 0
        barrier()
        e0.record()
        for _ in range(args.steps):
            _, l = step_device()
            launches += l
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        clk = clocks.stop() if rank == 0 else None
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        ms_per_step = ms / args.steps
        value = world * NQ * N_FRAMES * args.steps / (ms * 1e-3)

        # ---- e2e through the public API with host buffers (prompt H2D, per-frame D2H of the sample)
        e2e = None
        if not args.no_e2e:
            for _ in range(2):
                gen.generate_tts(task_prompt, "TTS", text_token=text, temperature=TEMPERATURE, topk=TOPK, fixed_schedule=(N_REASON, N_SEMANTIC), device_loop=True)
            barrier()
            e0.record()
            for _ in range(args.steps):
                r, s = gen.generate_tts(task_prompt, "TTS", text_token=text, temperature=TEMPERATURE, topk=TOPK,
                                        fixed_schedule=(N_REASON, N_SEMANTIC), device_loop=True)
                if world > 1:
                    gather_variable([torch.cat([r, s], dim=1)], world)
            e1.record()
            barrier()
            ems = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ems], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ems = float(t.item())
            assert gen.n_frames == N_FRAMES and r.shape == (NQ, N_REASON - 2) and s.shape == (NQ, N_SEMANTIC - 1), (gen.n_frames, r.shape, s.shape)
            e2e = {"value": round(world * NQ * N_FRAMES * args.steps / (ems * 1e-3), 1), "unit": UNIT,
                   "h2d_bytes_per_step": int(gen.h2d_bytes), "d2h_bytes_per_step": int(gen.d2h_bytes),
                   "ms_per_step": round(ems / args.steps, 2)}

        hbm_peak, peak_src = load_peaks()
        roofline = cpu_base = codec = flow = parity = whisper = frontends = None
        if rank == 0:
            roofline = time_dominant_kernel(model, hbm_peak)
            roofline["peak_source"] = peak_src
            mean_bytes = sum(frame_bytes(PROMPT_LEN + f) for f in range(N_FRAMES)) / N_FRAMES
            frame_ms = ms_per_step / N_FRAMES  # includes the prefill (1 of 180 calls)
            roofline["frame_effective_gbs"] = round(mean_bytes / (frame_ms * 1e-3) / 1e9, 1)
            roofline["frame_frac_of_peak"] = round(roofline["frame_effective_gbs"] / hbm_peak, 4)
            roofline["frame_algorithmic_bytes"] = mean_bytes
            if world == 1 and not args.no_cpu_baseline:
                try:
                    sd = state_dict_to_cpu(model)
                    cpu_base, _ = cpu_baseline_sample(sd, n_frames=args.cpu_frames, parity_frames=args.parity_frames)
                    del sd
                    ref = getattr(cpu_baseline_sample, "parity_ref", None)
                    if ref is not None:  # full-size parity: the 4.86 B-parameter GPU model against the CPU oracle, id for id
                        got = gpu_greedy_frames_teacher_forced(model, ref["tokens"], ref["mask"], ref["pos"], ref["frames"])
                        neq = (got != ref["frames"]).any(dim=1)
                        parity = {"frames": int(ref["frames"].size(0)), "ids_equal": bool(not neq.any()), "mismatched_frames": int(neq.sum()),
                                  "ids_compared": int(ref["frames"].numel()), "min_margin": round(ref["min_margin"], 6),
                                  "mode": "greedy topk=1, teacher-forced with the oracle's ids, prefill 39 positions"}
                except Exception as e:  # noqa: BLE001  (e.g. host memory): report it instead of losing the GPU measurement
                    cpu_base = {"error": f"{type(e).__name__}: {e}", "kind": "port"}
            if world == 1 and not args.no_codec:
                try:  # secondary metric: never let it take the headline line down
                    codec = bench_codec(dev, cpu=not args.no_cpu_baseline)
                except Exception as e:  # noqa: BLE001
                    codec = {"error": f"{type(e).__name__}: {e}"}
            if world == 1 and not args.no_flow_decoder:
                try:  # secondary metric (tokens -> latent of --stage all), same rule
                    flow = bench_flow_decoder(dev, cpu=not args.no_cpu_baseline)
                except Exception as e:  # noqa: BLE001
                    flow = {"error": f"{type(e).__name__}: {e}"}
                try:  # secondary metric (first SSL front-end of tokenize), same rule
                    whisper = bench_whisper_encoder(dev, cpu=not args.no_cpu_baseline)
                except Exception as e:  # noqa: BLE001
                    whisper = {"error": f"{type(e).__name__}: {e}"}
                try:  # secondary metric (waveform front-end + second SSL encoder of tokenize), same rule
                    frontends = bench_tokenize_frontends(dev, cpu=not args.no_cpu_baseline)
                except Exception as e:  # noqa: BLE001
                    frontends = {"error": f"{type(e).__name__}: {e}"}
    # ---- codec at N > 1: every rank encodes + decodes its own batch of 16 x 10 s clips (replicas, weak scaling), time = max over ranks;
    #      --codec-sweep: the clip length x batch grid of BASELINE.json config 5 with the clips of a point dealt to the ranks
    if world > 1 and not args.no_codec:
        with torch.inference_mode():
            try:
                c = bench_codec(dev, cpu=False, roofline=False)
                t = torch.tensor([c["encode_ms"] + c["decode_ms"]], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                if rank == 0:
                    codec = {"config": c["config"] + f" per GPU x {world} GPUs (replicas)", "ms_max_over_ranks": round(float(t.item()), 2),
                             "rtf_x_realtime": round(world * 160.0 / (float(t.item()) * 1e-3), 1)}
            except Exception as e:  # noqa: BLE001
                codec = {"error": f"{type(e).__name__}: {e}"}
    sweep = None
    if args.codec_sweep:
        with torch.inference_mode():
            pts = codec_sweep(dev, rank, world)
            t = torch.tensor([p[2] for p in pts], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sweep = [{"clip_s": c, "batch": b, "ms": round(float(ms), 2), "x_realtime": round(c * b / (float(ms) * 1e-3), 1) if float(ms) > 0 else None}
                     for (c, b, _), ms in zip(pts, t.tolist())]
    if rank == 0:
        print(json.dumps({"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                          "warmup": warmup, "ms_per_step": round(ms_per_step, 2), "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "fp32", "data": "synthetic", "config": config, "e2e": e2e,
                          "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "cpu_baseline": cpu_base, "parity": parity,
                          "codec": codec, "codec_sweep": sweep, "flow_decoder": flow, "whisper_encoder": whisper, "tokenize_frontends": frontends}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
