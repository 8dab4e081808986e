"""CPU suite: the kernel SOURCES of csrc/ua2_frontend.cu (resampler, log-mel features) and csrc/ua2_wavlm.cu (the WavLM encoder's own
kernels) plus the bias variant of csrc/ua2_dit.cu's attention kernel, compiled with g++ against the thread-per-CUDA-thread shim of
tests/cpu_shim/ (see tests/test_kernels_on_cpu_shim.py) and checked against oracle/frontend_oracle.py and oracle/wavlm_oracle.py -
which are themselves pinned against torchaudio / transformers (tests/test_frontend_oracle.py, tests/test_wavlm_oracle.py).
Shared-memory indexing, barriers, reflect / zero padding, tap loops and reductions as written; outside the shim: the GEMMs."""
import ctypes as C
import math
import os
import re
import subprocess

import pytest
import torch
import torch.nn.functional as F

from oracle import frontend_oracle as FO
from oracle import wavlm_oracle as WO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "tests", "cpu_shim")
CSRC = os.path.join(ROOT, "uniaudio2_b200", "csrc")
GXX = ["g++", "-std=c++20", "-O1", "-shared", "-fPIC", "-pthread"] + (["-fsanitize=address", "-fno-omit-frame-pointer", "-g"]
                                                                        if os.environ.get("UA2_SHIM_ASAN") == "1" else [])
STRIP = r'#include ["<](\.\./\.\./include/ua2_b200\.h|ua2_kernels\.cuh|ua2_umma\.cuh|ua2_enc_dev\.cuh|ua2_philox\.cuh|cuda_bf16\.h)[">]'


def _kernel_part(name, d):
    src = open(os.path.join(CSRC, name + ".cu")).read()
    src = src[:src.index("\nusing namespace ua2;")]  # kernels + launchers; the C-ABI / handle code behind it needs the CUDA runtime
    src = re.sub(r"extern __shared__[^;]*;", "", src)
    open(os.path.join(d, name + "_kernels.inc"), "w").write(re.sub(STRIP, "", src))


def _build(d, harness, out):
    so = os.path.join(d, out)
    r = subprocess.run(GXX + ["-DUA2_CPU_SHIM", "-I", d, "-I", SHIM, "-I", CSRC, os.path.join(SHIM, harness), "-o", so], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    return C.CDLL(so)


@pytest.fixture(scope="module")
def fe(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("shim_fe"))
    _kernel_part("ua2_frontend", d)
    _kernel_part("ua2_wavlm", d)
    return _build(d, "harness_frontend.cpp", "libshim_fe.so")


@pytest.fixture(scope="module")
def dit(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("shim_dit"))
    _kernel_part("ua2_stream", d)
    _kernel_part("ua2_dit", d)
    return _build(d, "harness_stream_dit.cpp", "libshim_dit.so")


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _rel(a, b):
    return float((a - b).abs().max()) / max(1.0, float(b.abs().max()))


@pytest.mark.parametrize("orig,new,L,extra", [(24000, 16000, 1000, 7), (16000, 24000, 333, 0), (24000, 16000, 31, 160)])
def test_resampler_source_on_cpu(fe, orig, new, L, extra):
    g = torch.Generator().manual_seed(L)
    x = torch.randn(2, L, generator=g)
    kern, width, o, n = FO.sinc_resample_kernel(orig, new)
    ref = FO.resample(x, orig, new)
    n_valid = ref.shape[-1]
    y = torch.full((2, n_valid + extra), float("nan"))
    assert fe.shim_fe_resample(_p(x), L, _p(kern.reshape(n, -1).contiguous()), _p(y), n_valid + extra, 2, L, n_valid, n_valid + extra, o, n, width) == 0
    assert float((y[:, :n_valid] - ref).abs().max()) < 2e-6 * max(1.0, float(ref.abs().max()))
    assert bool((y[:, n_valid:] == 0).all())  # the padding region is written as zeros


def test_log_mel_source_on_cpu(fe):
    """A short clip (29 frames: 4 CTAs per clip, a ragged last one) with the production transform sizes, against the torch.stft
    restatement of WhisperFeatureExtractor."""
    g = torch.Generator().manual_seed(3)
    L = 29 * FO.HOP
    wav = torch.randn(2, L, generator=g) * 0.2
    wav[1] = torch.sin(torch.arange(L) * 0.11) * 0.5 + 1e-3 * wav[1]  # tonal: most bins sit on the max - 8 floor
    win = torch.hann_window(FO.N_FFT)
    filt = torch.from_numpy(FO.mel_filter_bank()).float().contiguous()
    out = torch.full((2, FO.N_MELS, 29), float("nan"))
    assert fe.shim_fe_logmel(_p(wav), L, _p(win), _p(filt), _p(out), 2, L, FO.N_FFT, FO.HOP, FO.N_MELS, 29) == 0
    # the reference arithmetic on the same (unpadded) clip
    stft = torch.stft(wav, FO.N_FFT, FO.HOP, window=win, return_complex=True)
    mel = filt.T @ (stft[..., :-1].abs() ** 2)
    ls = torch.clamp(mel, min=1e-10).log10()
    ls = (torch.maximum(ls, ls.amax(dim=(1, 2), keepdim=True) - 8.0) + 4.0) / 4.0
    assert float((out - ls).abs().max()) < 2e-4


def test_wavlm_conv0_groupnorm_source_on_cpu(fe):
    g = torch.Generator().manual_seed(5)
    B, L, C0, k0, s0 = 2, 1500, 24, 10, 5
    T0 = (L - k0) // s0 + 1  # 299 frames: two chunks of 256
    x = torch.randn(B, L, generator=g) * 0.3
    w = torch.randn(C0, 1, k0, generator=g) / math.sqrt(k0)
    gamma, beta = 1 + 0.2 * torch.randn(C0, generator=g), 0.1 * torch.randn(C0, generator=g)
    ref = F.gelu(F.group_norm(F.conv1d(x[:, None], w, stride=s0), C0, gamma, beta, 1e-5)).transpose(1, 2)
    n_chunks = (T0 + 255) // 256
    part = torch.zeros(B * n_chunks * C0 * 2, dtype=torch.float64)
    stat = torch.zeros(B * C0 * 2)
    out = torch.full((B, T0, C0), float("nan"))
    assert fe.shim_wl_conv0(_p(x), L, _p(w.contiguous()), None, _p(gamma), _p(beta), _p(part), _p(stat), _p(out), B, L, T0, C0, k0, s0, C.c_float(1e-5)) == 0
    assert _rel(out, ref) < 2e-6
    bias = 0.2 * torch.randn(C0, generator=g)
    ref_b = F.gelu(F.group_norm(F.conv1d(x[:, None], w, bias, stride=s0), C0, gamma, beta, 1e-5)).transpose(1, 2)
    assert fe.shim_wl_conv0(_p(x), L, _p(w.contiguous()), _p(bias), _p(gamma), _p(beta), _p(part), _p(stat), _p(out), B, L, T0, C0, k0, s0, C.c_float(1e-5)) == 0
    assert _rel(out, ref_b) < 2e-6


@pytest.mark.parametrize("k,s", [(3, 2), (2, 2)])
def test_wavlm_im2col_and_repack_source_on_cpu(fe, k, s):
    """conv as GEMM: im2col rows of a channels-last tensor times the repacked weight = F.conv1d."""
    g = torch.Generator().manual_seed(k)
    B, Tin, Cin, Cout = 2, 37, 8, 12
    Tout = (Tin - k) // s + 1
    x = torch.randn(B, Tin, Cin, generator=g)
    w = torch.randn(Cout, Cin, k, generator=g)
    col = torch.full((B * Tout, k * Cin), float("nan"))
    wr = torch.full((Cout, k * Cin), float("nan"))
    assert fe.shim_wl_im2col(_p(x), _p(col), B, Tin, Tout, Cin, k, s) == 0
    assert fe.shim_wl_repack_conv(_p(w), _p(wr), Cout, Cin, k) == 0
    ref = F.conv1d(x.transpose(1, 2), w, stride=s).transpose(1, 2).reshape(B * Tout, Cout)
    assert _rel(col @ wr.T, ref) < 1e-5


@pytest.mark.parametrize("D,groups,K,T", [(64, 4, 16, 45), (96, 2, 8, 33), (48, 4, 128, 20)])
def test_wavlm_positional_conv_source_on_cpu(fe, D, groups, K, T):
    g = torch.Generator().manual_seed(K)
    B, cg = 2, D // groups
    h = torch.randn(B, T, D, generator=g)
    w = torch.randn(D, cg, K, generator=g) / math.sqrt(cg * K)
    bias = 0.1 * torch.randn(D, generator=g)
    ref = F.gelu(F.conv1d(h.transpose(1, 2), w, bias, padding=K // 2, groups=groups)[:, :, :-1]).transpose(1, 2)
    wr = torch.full((D * cg * K,), float("nan"))
    p = torch.full((B, T, D), float("nan"))
    assert fe.shim_wl_posconv(_p(h), _p(w), _p(bias), _p(wr), _p(p), B, T, D, cg, K) == 0
    assert _rel(p, ref) < 1e-5


def _gate_inputs(g, B, T, H, hs):
    h = torch.randn(B, T, H * hs, generator=g)
    sd = {"gru_rel_pos_linear.weight": torch.randn(8, hs, generator=g) / math.sqrt(hs), "gru_rel_pos_linear.bias": 0.1 * torch.randn(8, generator=g),
          "gru_rel_pos_const": 1 + 0.3 * torch.randn(1, H, 1, 1, generator=g)}
    return h, sd


def test_wavlm_gate_and_bias_table_source_on_cpu(fe):
    g = torch.Generator().manual_seed(9)
    B, T, H, hs = 2, 11, 3, 64
    h, sd = _gate_inputs(g, B, T, H, hs)
    ref = WO.gate(sd, {"num_attention_heads": H}, "", h)
    gate = torch.full((B, H, T), float("nan"))
    assert fe.shim_wl_gate(_p(h), _p(sd["gru_rel_pos_linear.weight"]), _p(sd["gru_rel_pos_linear.bias"]), _p(sd["gru_rel_pos_const"].reshape(H).contiguous()),
                           _p(gate), B, T, H, hs) == 0
    assert _rel(gate, ref) < 1e-6
    emb = torch.randn(320, H, generator=g)
    Tt = 200  # reaches the logarithmic buckets
    tab = torch.full((H, 2 * Tt - 1), float("nan"))
    assert fe.shim_wl_bias_table(_p(emb), _p(tab), H, Tt, 320, 800) == 0
    pb = WO.position_bias({"encoder.layers.0.attention.rel_attn_embed.weight": emb}, {"num_buckets": 320, "max_bucket_distance": 800}, Tt)
    i = torch.arange(Tt)[:, None]
    j = torch.arange(Tt)[None, :]
    assert torch.equal(tab[:, (j - i + Tt - 1)], pb)


def test_wavlm_axpy_source_on_cpu(fe):
    g = torch.Generator().manual_seed(1)
    a, b = torch.randn(40, generator=g), torch.randn(40, generator=g)
    out = torch.full((40,), float("nan"))
    assert fe.shim_wl_axpy(_p(a), _p(out), C.c_float(0.25), 1, 10) == 0
    assert fe.shim_wl_axpy(_p(b), _p(out), C.c_float(0.25), 0, 10) == 0
    assert torch.allclose(out, 0.25 * a + 0.25 * b, atol=1e-7)


@pytest.mark.parametrize("hs,T,B", [(32, 45, 2), (64, 33, 1), (128, 20, 1)])
def test_biased_attention_source_on_cpu(dit, hs, T, B):
    """dit_attn_kernel<.., BIAS = true>: scores + gate[b, h, i] * tab[h, j - i + T - 1], against softmax with the materialised bias."""
    H = 2
    g = torch.Generator().manual_seed(T)
    q = torch.randn(B, T, H, hs, generator=g)
    k = torch.randn(B, H, T, hs, generator=g).contiguous()
    v = torch.randn(B, H, T, hs, generator=g).contiguous()
    gate = (1 + torch.rand(B, H, T, generator=g)).contiguous()
    tab = torch.randn(H, 2 * T - 1, generator=g).contiguous()
    i = torch.arange(T)[:, None]
    j = torch.arange(T)[None, :]
    bias = gate[..., None] * tab[:, (j - i + T - 1)][None]
    ref = F.scaled_dot_product_attention(q.permute(0, 2, 1, 3), k, v, attn_mask=bias).permute(0, 2, 1, 3).reshape(B * T, H * hs)
    out = torch.full((B * T, H * hs), float("nan"))
    assert dit.shim_dit_attn_bias(_p(q.reshape(B * T, H * hs).contiguous()), _p(k), _p(v), _p(out), B, T, H, hs, _p(gate), _p(tab)) == 0
    assert _rel(out, ref) < 1e-5


def test_whole_wavlm_handle_on_cpu_against_transformers_golden(encoder_handles_shim):
    """csrc/ua2_wavlm.cu as shipped - parameter loading, weight repacks, workspace sizing, the feature encoder, feature projection,
    positional convolution, post-LayerNorm layers with the gated relative position bias, hidden-state outputs and their mean - on CPU
    tensors through the shim, against the hidden states of the real transformers.WavLMModel (tests/golden/frontend_golden.pt)."""
    from uniaudio2_b200 import _lib

    lib = encoder_handles_shim
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "frontend_golden.pt"), weights_only=False)
    c = gold["wavlm_small"]
    cfg = c["cfg"]
    sd = WO.random_state_dict(cfg, c["seed"])
    n = len(cfg["conv_dim"])
    arr = lambda v: (C.c_int32 * 8)(*(list(v) + [0] * (8 - n)))
    ccfg = _lib.WavLMCfg(cfg["hidden_size"], cfg["num_attention_heads"], cfg["intermediate_size"], cfg["num_hidden_layers"], n, arr(cfg["conv_dim"]),
                         arr(cfg["conv_kernel"]), arr(cfg["conv_stride"]), 1 if cfg["conv_bias"] else 0, cfg["num_conv_pos_embeddings"],
                         cfg["num_conv_pos_embedding_groups"], cfg["num_buckets"], cfg["max_bucket_distance"], cfg["layer_norm_eps"])
    h = C.c_void_p()
    assert lib.ua2_wavlm_create(C.byref(ccfg), C.byref(h)) == 0, lib.ua2_last_error()
    sd = dict(sd)
    pre = "encoder.pos_conv_embed.conv."
    g, v = sd.pop(pre + "parametrizations.weight.original0"), sd.pop(pre + "parametrizations.weight.original1")
    sd[pre + "weight"] = (g * (v / v.norm(dim=(0, 1), keepdim=True))).contiguous()
    sd.pop("masked_spec_embed")
    keep = []
    for key, t in sd.items():
        t = t.contiguous()
        keep.append(t)
        shape = (C.c_int64 * t.dim())(*t.shape)
        assert lib.ua2_wavlm_load_weight(h, key.encode(), _p(t), shape, t.dim()) == 0, (key, lib.ua2_last_error())
    assert lib.ua2_wavlm_finalize(h, None) == 0, lib.ua2_last_error()
    wav = c["wav16"][:, :1000].contiguous()  # a shorter clip keeps the OS-thread emulation to seconds; reference below from the oracle
    with torch.no_grad():
        ref = WO.hidden_states(sd | {}, cfg, wav)
    B, L = wav.shape
    T = int(lib.ua2_wavlm_frames(h, L))
    assert T == ref[0].shape[1]
    nl = cfg["num_hidden_layers"]
    out = torch.full((B, T, cfg["hidden_size"]), float("nan"))
    allh = torch.full((nl + 1, B, T, cfg["hidden_size"]), float("nan"))
    assert lib.ua2_wavlm_forward(h, _p(wav), L, B, L, 1, nl + 1, _p(out), _p(allh), None) == 0, lib.ua2_last_error()
    for i in range(nl + 1):
        assert _rel(allh[i], ref[i]) < 2e-5, (i, _rel(allh[i], ref[i]))
    assert _rel(out, torch.stack(ref, 1)[:, 1:nl + 1].mean(1)) < 2e-5
    # only the needed layers run, and a second call (cached bias table, reused workspace) gives the same result
    out2 = torch.full_like(out, float("nan"))
    assert lib.ua2_wavlm_forward(h, _p(wav), L, B, L, 0, 2, _p(out2), None, None) == 0, lib.ua2_last_error()
    assert _rel(out2, (ref[0] + ref[1]) / 2) < 2e-5
    assert lib.ua2_wavlm_forward(h, _p(wav), L, B, L, 0, nl + 2, _p(out2), None, None) != 0  # range beyond the layers is refused
    assert lib.ua2_wavlm_destroy(h) == 0
