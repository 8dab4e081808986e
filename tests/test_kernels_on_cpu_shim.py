"""CPU suite: the kernel SOURCES of uniaudio2_b200/csrc compiled with g++ against a thread-per-CUDA-thread shim (tests/cpu_shim/:
one OS thread per CUDA thread, barriers for __syncthreads / __syncwarp, warp shuffles, atomics, software bf16, emulated mbarrier +
bulk copy, exactly-sized dynamic shared memory) and executed through the product's own launchers, with only the tensor-core GEMM
replaced by a CPU GEMM.  Three builds: (1) the conv_tc / resblock kernels behind a small harness, (2) the kernel halves of
ua2_stream.cu / ua2_dit.cu, (3) whole translation units with the real headers (-DUA2_CPU_SHIM) exporting their C-ABI operators,
up to the complete codec handle.  It checks the kernels as written - shared-memory indexing, barriers, halos, ring parities,
chunk loops - against the oracles and the reference's golden fixtures; tools/shim_asan.sh runs it under the address sanitizer.
Outside the shim: tensor cores, clusters / DSMEM, the GPU's weak memory model."""
import ctypes as C
import math
import os
import re
import subprocess

import pytest
import torch
import torch.nn.functional as F

from oracle import codec_oracle as CO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "tests", "cpu_shim")
CSRC = os.path.join(ROOT, "uniaudio2_b200", "csrc")
# UA2_SHIM_ASAN=1: build the harnesses with the address sanitizer (run pytest under LD_PRELOAD=$(gcc -print-file-name=libasan.so)
# ASAN_OPTIONS=detect_leaks=0): global-memory overruns of the kernels land in the red zones of the torch CPU allocations, shared-
# memory overruns in those of the static arrays / of the exactly-sized dynamic block.  tools/shim_asan.sh is the whole command.
GXX = ["g++", "-std=c++20", "-O1", "-shared", "-fPIC", "-pthread"] + (["-fsanitize=address", "-fno-omit-frame-pointer", "-g"]
                                                                        if os.environ.get("UA2_SHIM_ASAN") == "1" else [])


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("shim"))
    for name in ("ua2_convtc", "ua2_resblock"):
        src = open(os.path.join(CSRC, name + ".cu")).read()
        # the dynamic shared-memory array is defined (aligned) by the harness: drop the kernel-local declaration
        src = re.sub(r"extern __shared__[^;]*;", "", src)
        src = re.sub(r'#include "ua2_kernels.cuh"', "", src)
        open(os.path.join(d, name + "_shim.inc"), "w").write(src)
    so = os.path.join(d, "libshim.so")
    cmd = GXX + [ "-I", d, "-I", SHIM, os.path.join(SHIM, "harness.cpp"), "-o", so]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return C.CDLL(so)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


@pytest.mark.parametrize("B,Cin,Cout,T,K,stride,elu,use_res,rep", [(4, 128, 12, 161, 8, 4, 1, 0, 0), (2, 352, 8, 70, 3, 1, 1, 1, 0),
                                                                     (2, 256, 8, 131, 4, 2, 0, 0, 1)])
def test_conv_tc_source_on_cpu(shim, B, Cin, Cout, T, K, stride, elu, use_res, rep):
    g = torch.Generator().manual_seed(T)
    x = torch.randn(B, Cin, T, generator=g)
    w = (torch.randn(Cout, Cin, K, generator=g) / math.sqrt(Cin * K)).contiguous()
    b = None if rep else torch.randn(Cout, generator=g) * 0.1
    ref = CO.conv1d_causal(F.elu(x) if elu else x, w, b, stride=stride, dilation=1, pad_mode="replicate" if rep else "constant")
    r = torch.randn_like(ref) if use_res else None
    if use_res:
        ref = r + ref
    y = torch.full_like(ref, float("nan"))
    rc = shim.shim_conv1d_tc(_p(x), _p(w), _p(b), _p(r), _p(y), B, Cin, Cout, T, K, stride, 1, elu, rep)
    assert rc == 0, rc
    assert float((y - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("B,Cin,Cout,T,stride", [(2, 128, 12, 70, 4), (1, 160, 8, 129, 5), (3, 128, 4, 44, 8)])
def test_convtr_tc_source_on_cpu(shim, B, Cin, Cout, T, stride):
    g = torch.Generator().manual_seed(T + stride)
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn(Cin, Cout, 2 * stride, generator=g) / math.sqrt(2 * Cin)
    b = torch.randn(Cout, generator=g) * 0.1
    ref = CO.convtr1d_causal(F.elu(x), w, b, stride)
    # w_phase[ph][co][ci][tap] = w[ci][co][ph + tap * s]   (repack_convtr_phase_kernel, csrc/ua2_codec.cu)
    wp = torch.stack([torch.stack([w[:, :, ph + tap * stride].t() for tap in range(2)], dim=-1) for ph in range(stride)], dim=0).contiguous()
    assert wp.shape == (stride, Cout, Cin, 2)
    y = torch.full_like(ref, float("nan"))
    rc = shim.shim_convtr1d_tc(_p(x), _p(wp), _p(b), _p(y), B, Cin, Cout, T, stride, 1)
    assert rc == 0, rc
    assert float((y - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("B,T", [(1, 128), (2, 300), (1, 5)])
def test_resblock_source_on_cpu(shim, B, T):
    g = torch.Generator().manual_seed(T)
    C_, H = 64, 32
    x = torch.randn(B, C_, T, generator=g)
    w1 = (torch.randn(H, C_, 3, generator=g) / math.sqrt(C_ * 3)).contiguous()
    b1 = torch.randn(H, generator=g) * 0.1
    w2 = (torch.randn(C_, H, 1, generator=g) / math.sqrt(H)).contiguous()
    b2 = torch.randn(C_, generator=g) * 0.1
    ref = x + CO.conv1d_causal(F.elu(CO.conv1d_causal(F.elu(x), w1, b1)), w2, b2)
    y = torch.full_like(x, float("nan"))
    rc = shim.shim_resblock(_p(x), _p(w1), _p(b1), _p(w2), _p(b2), _p(y), B, C_, H, T)
    assert rc == 0, rc
    assert float((y - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))
    assert shim.shim_resblock(_p(x), _p(w1), _p(b1), _p(w2), _p(b2), _p(y), B, 128, 64, T) == 801  # not served: cudaErrorNotSupported


# ------------------------------------------------------------------------------------------------------------------------
# kernels that ARE parity-green on the GPU (ua2_stream.cu, ua2_dit.cu): their source also runs on the shim, so that the CPU-only
# test tier notices regressions in them
@pytest.fixture(scope="module")
def shim2(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("shim2"))
    for name in ("ua2_stream", "ua2_dit"):
        src = open(os.path.join(CSRC, name + ".cu")).read()
        src = src[:src.index("\nusing namespace ua2;")]  # kernels + launchers; the handle code behind it needs the CUDA runtime
        src = re.sub(r"extern __shared__[^;]*;", "", src)
        src = re.sub(r'#include ["<](\.\./\.\./include/ua2_b200\.h|ua2_kernels\.cuh|ua2_umma\.cuh|ua2_philox\.cuh|cuda_bf16\.h)[">]', "", src)
        open(os.path.join(d, name + "_kernels.inc"), "w").write(src)
    so = os.path.join(d, "libshim2.so")
    cmd = GXX + ["-DUA2_CPU_SHIM", "-I", d, "-I", SHIM, "-I", CSRC,  # the macro drops the tensor-core-only epilogue of ua2_dit.cu
           os.path.join(SHIM, "harness_stream_dit.cpp"), "-o", so]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    return C.CDLL(so)


def _rel(a, b):
    return float((a - b).abs().max()) / max(1.0, float(b.abs().max()))


@pytest.mark.parametrize("hs,n_splits", [(32, 1), (64, 1), (32, 2)])
def test_ring_attention_source_on_cpu(shim2, hs, n_splits):
    """rope_ring_append_kernel + ring_attn_kernel (+ combine): wrapped ring, T = 3 query rows per sequence - the shapes of
    tests/test_zz_moshi_gpu.py::test_ring_attention_operator, smaller."""
    from oracle import moshi_oracle as MO

    B, H, cap, T, offset, context = 2, 2, 9, 3, 20, 7
    Cd = H * hs
    g = torch.Generator().manual_seed(hs)
    ring = MO.RingKV(B, H, hs, cap)
    for t0 in range(0, offset, 4):
        n = min(4, offset - t0)
        ring.complete(torch.randn(B, H, n, hs, generator=g), torch.randn(B, H, n, hs, generator=g))
    kc, vc = ring.cache[0].clone().contiguous(), ring.cache[1].clone().contiguous()
    qkv = torch.randn(B, T, 3, H, hs, generator=g)
    q, k, v = qkv.permute(2, 0, 3, 1, 4)
    off_t = torch.full((1,), offset, dtype=torch.long)
    qr, kr = MO.apply_rope(q, k, off_t, 10000.0)
    keys, values, pos_k = ring.complete(kr, v)
    delta = (off_t + torch.arange(T).view(-1, 1)) - pos_k.view(1, -1)
    bias = (pos_k.view(1, -1) >= 0) & (delta >= 0) & (delta < context)
    ref = F.scaled_dot_product_attention(qr, keys, values, bias).permute(0, 2, 1, 3).reshape(B * T, Cd)
    pos = (offset + torch.arange(B * T) % T).int()
    bidx = (torch.arange(B * T) // T).int()
    freqs = MO.rope_freqs(hs, 10000.0).contiguous()
    qkv_f = qkv.reshape(B * T, 3 * Cd).contiguous()
    q_out, y = torch.empty(B * T, Cd), torch.full((B * T, Cd), float("nan"))
    assert shim2.shim_rope_ring_append(_p(qkv_f), 3 * Cd, _p(pos), _p(bidx), _p(freqs), _p(q_out), _p(kc), _p(vc), B * T, H, hs, cap, 1) == 0
    assert _rel(kc, ring.cache[0]) < 1e-5 and torch.equal(vc, ring.cache[1])
    part_ml, part_acc = torch.zeros(B * T * H * n_splits * 2), torch.zeros(B * T * H * n_splits * hs)
    rc = shim2.shim_ring_attn(_p(q_out), _p(kc), _p(vc), _p(pos), _p(bidx), _p(y), B * T, H, hs, cap, C.c_longlong(offset + T), 1, 1, context,
                              n_splits, _p(part_ml), _p(part_acc))
    assert rc == 0
    assert _rel(y, ref) < 1e-4


def test_sample_token_source_on_cpu(shim2):
    from oracle import moshi_oracle as MO

    g = torch.Generator().manual_seed(4)
    logits = (torch.randn(3, 200, generator=g) * 2.5).contiguous()

    def run(use_sampling, temp, top_k, top_p, end_token, noise):
        out = torch.full((3,), -1, dtype=torch.int64)
        rc = shim2.shim_sample_token(_p(logits), 3, 200, use_sampling, C.c_float(temp), top_k, C.c_float(top_p), end_token, _p(noise), _p(out))
        assert rc == 0
        return out

    assert torch.equal(run(0, 1.0, 0, 0.0, -1, None), logits.argmax(-1))
    q = torch.empty(3, 200).exponential_(1, generator=g)
    assert torch.equal(run(1, 0.8, 0, 0.0, -1, q), MO.sample_token(logits, True, 0.8, q=q))
    qk = torch.empty(3, 17).exponential_(1, generator=g)
    assert torch.equal(run(1, 0.9, 17, 0.0, -1, qk), MO.sample_token(logits, True, 0.9, top_k=17, q=qk))
    assert torch.equal(run(1, 1.1, 0, 0.7, -1, q), MO.sample_token(logits, True, 1.1, top_p=0.7, q=q))
    assert torch.equal(run(1, 0.9, 17, 0.0, 120, qk), MO.sample_token(logits, True, 0.9, top_k=17, q=qk, end_token=120))
    tied = torch.randint(0, 3, (3, 200), generator=g).float().contiguous()  # ties at the threshold: any of the k largest is valid
    out = torch.full((3,), -1, dtype=torch.int64)
    assert shim2.shim_sample_token(_p(tied), 3, 200, 1, C.c_float(1.0), 9, C.c_float(0.0), -1, _p(qk[:, :9].contiguous()), _p(out)) == 0
    assert bool((tied.gather(1, out[:, None])[:, 0] >= torch.topk(tied, 9).values[:, -1]).all())


@pytest.mark.parametrize("hs,T,B", [(32, 45, 2), (64, 33, 1), (128, 20, 1)])
def test_dit_attention_source_on_cpu(shim2, hs, T, B):
    H = 2
    g = torch.Generator().manual_seed(T)
    q = torch.randn(B, T, H, hs, generator=g)
    k = torch.randn(B, H, T, hs, generator=g).contiguous()
    v = torch.randn(B, H, T, hs, generator=g).contiguous()
    ref = F.scaled_dot_product_attention(q.permute(0, 2, 1, 3), k, v).permute(0, 2, 1, 3).reshape(B * T, H * hs)
    out = torch.full((B * T, H * hs), float("nan"))
    assert shim2.shim_dit_attn(_p(q.reshape(B * T, H * hs).contiguous()), _p(k), _p(v), _p(out), B, T, H, hs) == 0
    assert _rel(out, ref) < 1e-5


def test_dit_glue_sources_on_cpu(shim2):
    """LayerNorm + adaLN modulation, the k = 3 im2col, bias / gate / GELU / head-split epilogues and the Euler glue."""
    g = torch.Generator().manual_seed(8)
    B, T, D, H, hs = 2, 7, 64, 2, 32
    M = B * T
    x = torch.randn(M, D, generator=g)
    table = torch.randn(6, D, generator=g)
    t6 = torch.randn(B, 6 * D, generator=g)
    out = torch.empty(M, D)
    assert shim2.shim_dit_ln_mod(_p(x), _p(out), _p(table), _p(t6), 6 * D, D, 3, 4, C.c_float(1e-6), M, T, D) == 0
    mod = table[None] + t6.reshape(B, 6, D)
    ref = F.layer_norm(x.view(B, T, D), (D,), None, None, 1e-6) * (1 + mod[:, 4:5]) + mod[:, 3:4]
    assert _rel(out.view(B, T, D), ref) < 1e-5
    # im2col of the k = 3 'same' convolution
    Cc = 5
    xi = torch.randn(B, T, Cc, generator=g).contiguous()
    col = torch.empty(M, 3 * Cc)
    assert shim2.shim_dit_im2col3(_p(xi), _p(col), B, T, Cc) == 0
    w = torch.randn(4, Cc, 3, generator=g)
    ref = F.conv1d(xi.transpose(1, 2), w, None, padding=1).transpose(1, 2).reshape(M, 4)
    # column k * C + c of the im2col matrix multiplies w[:, c, k]  (dit_repack_conv3_kernel)
    assert _rel(col @ w.permute(0, 2, 1).reshape(4, 3 * Cc).t(), ref) < 1e-5
    # gate * (y + bias) + residual
    src, bias, res = torch.randn(M, D, generator=g), torch.randn(D, generator=g), torch.randn(M, D, generator=g)
    res0 = res.clone()
    assert shim2.shim_dit_gate_res(_p(src), _p(bias), _p(res), _p(table), _p(t6), 2, M, D, T) == 0
    assert _rel(res.view(B, T, D), mod[:, 2:3] * (src + bias).view(B, T, D) + res0.view(B, T, D)) < 1e-6
    # q / k / v split with biases, K and V in (B, H, T, hs)
    raw, b3 = torch.randn(M, 3 * D, generator=g), torch.randn(3 * D, generator=g)
    qo, ko, vo = torch.empty(M, D), torch.empty(B, H, T, hs), torch.empty(B, H, T, hs)
    assert shim2.shim_dit_qkv_split(_p(raw), _p(b3), _p(qo), _p(ko), _p(vo), M, D, T, H, hs) == 0
    full = (raw + b3).view(B, T, 3, H, hs)
    assert torch.equal(qo, full[:, :, 0].reshape(M, D)) and torch.equal(ko, full[:, :, 1].permute(0, 2, 1, 3)) and torch.equal(vo, full[:, :, 2].permute(0, 2, 1, 3))
    # GELU (tanh form)
    yg = torch.empty(M, D)
    assert shim2.shim_dit_bias_gelu(_p(src), _p(bias), _p(yg), M, D) == 0
    assert _rel(yg, F.gelu(src + bias, approximate="tanh")) < 1e-6
    # Euler glue: in-context blend + CFG batch assembly, then guidance mix + update
    lat, cond, ic, tt = 6, 4, 3, 0.3
    xs, noise, inc, mu = (torch.randn(T, n, generator=g) for n in (lat, lat, lat, cond))
    xs0 = xs.clone()
    inp = torch.empty(2, T, 2 * lat + cond)
    assert shim2.shim_dit_euler(_p(xs), _p(noise), _p(inc), _p(mu), _p(inp), None, T, lat, cond, ic, C.c_float(tt), C.c_float(0.9999), C.c_float(0), C.c_float(0), 0) == 0
    xb = xs0.clone()
    xb[:ic] = (1 - torch.tensor(0.9999) * tt) * noise[:ic] + tt * inc[:ic]
    assert torch.allclose(xs, xb, atol=1e-6)
    assert torch.allclose(inp[0], torch.cat([xb, inc, torch.zeros(T, cond)], 1), atol=1e-6) and torch.allclose(inp[1], torch.cat([xb, inc, mu], 1), atol=1e-6)
    d = torch.randn(2, T, lat, generator=g).contiguous()
    x1 = xs.clone()
    assert shim2.shim_dit_euler(_p(x1), None, None, None, None, _p(d), T, lat, cond, ic, C.c_float(0), C.c_float(0), C.c_float(1.5), C.c_float(0.25), 1) == 0
    assert torch.allclose(x1, xs + 0.25 * (d[0] + 1.5 * (d[1] - d[0])), atol=1e-6)


# ------------------------------------------------------------------------------------------------------------------------
# whole translation units with the REAL headers: csrc/ua2_codec.cu + ua2_sgemm.cu (+ ua2_convtc.cu, ua2_resblock.cu) build with
# -DUA2_CPU_SHIM (csrc/ua2_common.cuh swaps its three PTX helpers and `launch` for shim versions) and export their C-ABI
# operators unchanged - the operator tests of tests/test_codec_gpu.py, at small shapes, on the CPU
@pytest.fixture(scope="module")
def shim3(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("shim3"))
    srcs = []
    hdr = os.path.join(ROOT, "include", "ua2_b200.h")
    for name in ("ua2_codec", "ua2_sgemm", "ua2_convtc", "ua2_resblock", "ua2_attn", "ua2_gemv3", "ua2_gemv", "ua2_misc", "ua2_codec_model"):
        src = open(os.path.join(CSRC, name + ".cu")).read()
        # dynamic shared memory -> the exactly-sized block that the shim's launch() allocates from the launcher's byte count
        src = re.sub(r"extern __shared__\s+(?:__align__\(\d+\)\s+)?(\w+)\s+(\w+)\[\];", r"\1* \2 = static_cast<\1*>(shim::g_dyn_smem);", src)
        assert "extern __shared__" not in src
        src = src.replace('#include "../../include/ua2_b200.h"', f'#include "{hdr}"')
        open(os.path.join(d, name + ".cpp"), "w").write(src)
        srcs.append(os.path.join(d, name + ".cpp"))
    stub = open(os.path.join(SHIM, "stubs_real_headers.cpp")).read().replace('#include "../../include/ua2_b200.h"', f'#include "{hdr}"')
    open(os.path.join(d, "stubs.cpp"), "w").write(stub)
    so = os.path.join(d, "libshim3.so")
    cmd = GXX + [ "-DUA2_CPU_SHIM", "-DUA2_ATTN_RING_MIN_ITEMS=64", "-I", CSRC, "-I", os.path.join(SHIM, "rt"),
           "-Wl,--no-undefined"] + srcs + [os.path.join(d, "stubs.cpp"), "-o", so]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    lib = C.CDLL(so)
    lib.ua2_last_error.restype = C.c_char_p
    return lib


def _ok(lib, rc):
    assert rc == 0, lib.ua2_last_error()


@pytest.mark.parametrize("B,Cin,Cout,T,K,stride,dil,elu,res,rep", [
    (1, 1, 64, 200, 7, 1, 1, 0, 0, 0), (1, 64, 32, 150, 3, 1, 1, 1, 0, 0), (1, 32, 64, 150, 1, 1, 1, 1, 1, 0), (2, 64, 128, 203, 8, 4, 1, 1, 0, 0),
    (1, 96, 200, 133, 12, 6, 1, 1, 0, 0), (2, 128, 128, 57, 4, 2, 1, 0, 0, 1), (1, 16, 24, 100, 3, 1, 2, 1, 0, 0), (1, 64, 1, 199, 3, 1, 1, 1, 0, 0)])
def test_conv1d_operators_on_cpu(shim3, B, Cin, Cout, T, K, stride, dil, elu, res, rep):
    """ua2_conv1d_causal_f32 (direct kernel) and ua2_conv1d_causal_gemm_f32 (implicit GEMM) - tests/test_codec_gpu.py::test_conv1d_causal."""
    g = torch.Generator().manual_seed(Cin + Cout + T)
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn(Cout, Cin, K, generator=g) / math.sqrt(Cin * K)
    b = None if rep else torch.randn(Cout, generator=g) * 0.1
    ref = CO.conv1d_causal(F.elu(x) if elu else x, w, b, stride=stride, dilation=dil, pad_mode="replicate" if rep else "constant")
    r = torch.randn_like(ref) if res else None
    if res:
        ref = r + ref
    tol = 2e-5 * max(1.0, float(ref.abs().max()))
    y, wd = torch.full_like(ref, float("nan")), w.permute(1, 2, 0).contiguous()
    _ok(shim3, shim3.ua2_conv1d_causal_f32(_p(x), _p(wd), _p(b), _p(r), _p(y), B, Cin, Cout, T, K, stride, dil, elu, rep, None))
    assert float((y - ref).abs().max()) < tol
    y2 = torch.full_like(ref, float("nan"))
    _ok(shim3, shim3.ua2_conv1d_causal_gemm_f32(_p(x), _p(w), _p(b), _p(r), _p(y2), B, Cin, Cout, T, K, stride, dil, elu, rep, None))
    assert float((y2 - ref).abs().max()) < tol


@pytest.mark.parametrize("B,Cin,Cout,T,stride", [(1, 128, 64, 30, 8), (2, 96, 40, 77, 6), (1, 48, 20, 65, 3), (1, 32, 32, 40, 2)])
def test_convtr1d_operators_on_cpu(shim3, B, Cin, Cout, T, stride):
    g = torch.Generator().manual_seed(Cin + T)
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn(Cin, Cout, 2 * stride, generator=g) / math.sqrt(Cin * 2)
    b = torch.randn(Cout, generator=g) * 0.1
    ref = CO.convtr1d_causal(F.elu(x), w, b, stride)
    tol = 2e-5 * max(1.0, float(ref.abs().max()))
    y, wd = torch.full_like(ref, float("nan")), w.permute(0, 2, 1).contiguous()
    _ok(shim3, shim3.ua2_convtr1d_causal_f32(_p(x), _p(wd), _p(b), _p(y), B, Cin, Cout, T, stride, 1, None))
    assert float((y - ref).abs().max()) < tol
    wp = torch.empty(stride * Cout * Cin * 2)
    _ok(shim3, shim3.ua2_convtr1d_repack_phase_f32(_p(w), _p(wp), Cin, Cout, stride, None))
    y2 = torch.full_like(ref, float("nan"))
    _ok(shim3, shim3.ua2_convtr1d_causal_gemm_f32(_p(x), _p(wp), _p(b), _p(y2), B, Cin, Cout, T, stride, 1, None))
    assert float((y2 - ref).abs().max()) < tol


def test_depthwise_film_interp_elementwise_on_cpu(shim3):
    g = torch.Generator().manual_seed(3)
    B, Cc, T, s = 2, 64, 37, 2
    x = torch.randn(B, Cc, T, generator=g)
    w = torch.randn(Cc, 1, 2 * s, generator=g)
    y = torch.empty(B, Cc, T * s)
    _ok(shim3, shim3.ua2_convtr1d_depthwise_f32(_p(x), _p(w), _p(y), B, Cc, T, s, None))
    assert float((y - CO.convtr1d_causal(x, w, None, s, groups=Cc)).abs().max()) < 1e-6
    from oracle import film_oracle as FO

    params, feat = torch.randn(3, 11, 2 * 24, generator=g), torch.randn(3, 11, 24, generator=g)
    mask = torch.tensor([0, 1, 0], dtype=torch.uint8)
    out = torch.empty(3, 11, 24)
    _ok(shim3, shim3.ua2_film_f32(_p(params), _p(feat), _p(mask), _p(out), 3, 11, 24, C.c_float(0.1), None))
    assert float((out - FO.time_film(params, feat, mask, 0.1)).abs().max()) < 1e-6
    r = torch.randn(2, 12, 30, generator=g)
    ref_i = F.interpolate(r, scale_factor=2.5, mode="nearest")
    yi = torch.empty_like(ref_i)
    _ok(shim3, shim3.ua2_interp_nearest_f32(_p(r), _p(yi), 2, 12, 30, ref_i.shape[-1], C.c_float(2.5), None))
    assert torch.equal(yi, ref_i)
    z = torch.randn(1000, generator=g)
    yz = torch.empty_like(z)
    _ok(shim3, shim3.ua2_elementwise_f32(_p(z), _p(yz), C.c_longlong(1000), 0, C.c_float(9.0), None))
    assert torch.equal(yz, torch.round(9 * z) / 9)  # round_func9, scalar24k.py:279-288


@pytest.mark.parametrize("B,D,T,K,n_q", [(2, 32, 21, 64, 4), (1, 64, 33, 300, 3)])
def test_rvq_operators_on_cpu(shim3, B, D, T, K, n_q):
    """ua2_rvq_encode_f32 / ua2_rvq_encode_gemm_f32 / ua2_rvq_decode_f32 - tests/test_codec_gpu.py::test_rvq_encode_decode."""
    g = torch.Generator().manual_seed(D + K)
    x = torch.randn(B, D, T, generator=g)
    emb = torch.randn(n_q, K, D, generator=g)
    residual, ref_codes = x.clone(), []
    for q in range(n_q):
        flat = residual.transpose(1, 2).reshape(-1, D)
        codes = torch.cdist(flat[None], emb[q][None], p=2)[0].argmin(-1).view(B, T)
        residual = residual - F.embedding(codes, emb[q]).transpose(1, 2)
        ref_codes.append(codes)
    ref_codes = torch.stack(ref_codes, 1)
    sq = (emb * emb).sum(-1).contiguous()
    codes = torch.full((B, n_q + 2, T), -1, dtype=torch.int64)
    _ok(shim3, shim3.ua2_rvq_encode_f32(_p(x), _p(emb), _p(sq), _p(codes), B, D, T, K, n_q, n_q + 2, 1, None))
    assert torch.equal(codes[:, 1:1 + n_q], ref_codes)
    assert int(codes[:, 0].max()) == -1 and int(codes[:, -1].max()) == -1
    r_md = x.transpose(1, 2).reshape(B * T, D).contiguous()
    S = torch.empty(B * T, K)
    codes2 = torch.full((B, n_q + 2, T), -1, dtype=torch.int64)
    _ok(shim3, shim3.ua2_rvq_encode_gemm_f32(_p(r_md), _p(emb), _p(sq), _p(S), _p(codes2), B, D, T, K, n_q, n_q + 2, 1, None))
    assert torch.equal(codes2[:, 1:1 + n_q], ref_codes)
    out = torch.empty(B, D, T)
    _ok(shim3, shim3.ua2_rvq_decode_f32(_p(codes), _p(emb), _p(out), B, D, T, K, n_q, n_q + 2, 1, None))
    ref = sum(F.embedding(ref_codes[:, q], emb[q]).transpose(1, 2) for q in range(n_q))
    assert float((out - ref).abs().max()) < 1e-5


@pytest.mark.parametrize("B,Cin,Cout,T,K,stride,dil,pl,pr", [(2, 96, 96, 120, 4, 4, 1, 0, 0), (1, 64, 64, 90, 2, 2, 1, 0, 0), (1, 48, 48, 199, 7, 1, 9, 27, 27),
                                                             (1, 136, 200, 50, 5, 1, 1, 2, 2)])
def test_general_conv1d_operator_on_cpu(shim3, B, Cin, Cout, T, K, stride, dil, pl, pr):
    """ua2_conv1d_f32: the nn.Conv1d forms of the ScalarModel decoder with bias + PReLU + residual fused (tests/test_scalar_gpu.py)."""
    g = torch.Generator().manual_seed(K + T)
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn(Cout, Cin, K, generator=g) / (Cin * K) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    slope = torch.tensor([0.2])
    ref = F.prelu(F.conv1d(F.pad(x, (pl, pr)), w, b, stride=stride, dilation=dil), slope)
    use_res = Cin == Cout and stride == 1 and ref.shape[-1] == T
    if use_res:
        ref = ref + x
    y = torch.full_like(ref, float("nan"))
    _ok(shim3, shim3.ua2_conv1d_f32(_p(x), _p(w), _p(b), _p(slope), _p(x) if use_res else None, _p(y), B, Cin, Cout, T, K, stride, dil, pl, pr, None))
    assert float((y - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("B,Cin,Cout,T,stride,crop", [(1, 96, 48, 40, 5, 2), (2, 64, 32, 33, 4, 0), (1, 48, 24, 50, 2, 1)])
def test_general_convtr1d_operator_on_cpu(shim3, B, Cin, Cout, T, stride, crop):
    """ua2_convtr1d_f32: nn.ConvTranspose1d (kernel 2 * stride) + PReLU, cropped to [crop, crop + T * stride) - ScalarModel upsamplers."""
    g = torch.Generator().manual_seed(Cin + stride)
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn(Cin, Cout, 2 * stride, generator=g) / (2 * Cin) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    slope = torch.tensor([0.3])
    T_out = T * stride
    ref = F.prelu(F.conv_transpose1d(x, w, b, stride=stride), slope)[..., crop:crop + T_out]
    wp = torch.empty(stride * Cout * Cin * 2)
    _ok(shim3, shim3.ua2_convtr1d_repack_phase_f32(_p(w), _p(wp), Cin, Cout, stride, None))
    y = torch.full_like(ref, float("nan"))
    _ok(shim3, shim3.ua2_convtr1d_f32(_p(x), _p(wp), _p(b), _p(slope), _p(y), B, Cin, Cout, T, stride, crop, T_out, None))
    assert float((y - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))
    assert shim3.ua2_convtr1d_f32(_p(x), _p(wp), _p(b), _p(slope), _p(y), B, Cin, Cout, T, stride, stride + 1, T_out, None) != 0  # window too far right


@pytest.mark.parametrize("B,T", [(2, 300), (1, 128), (1, 77)])
def test_resblock_operator_on_cpu(shim3, B, T):
    """ua2_resblock_f32 through its C-ABI entry (argument checks included): y = x + conv_k1(ELU(conv_k3(ELU(x)))), modules/seanet.py:21-94."""
    g = torch.Generator().manual_seed(T)
    x = torch.randn(B, 64, T, generator=g)
    w1, b1 = torch.randn(32, 64, 3, generator=g) / 14, torch.randn(32, generator=g) * 0.1
    w2, b2 = torch.randn(64, 32, 1, generator=g) / 6, torch.randn(64, generator=g) * 0.1
    ref = x + CO.conv1d_causal(F.elu(CO.conv1d_causal(F.elu(x), w1, b1)), w2, b2)
    y = torch.full_like(ref, float("nan"))
    _ok(shim3, shim3.ua2_resblock_f32(_p(x), _p(w1), _p(b1), _p(w2), _p(b2), _p(y), B, 64, 32, T, None))
    assert float((y - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))
    assert shim3.ua2_resblock_f32(_p(x), _p(w1), _p(b1), _p(w2), _p(b2), _p(x), B, 64, 32, T, None) != 0    # aliasing refused
    assert shim3.ua2_resblock_f32(_p(x), _p(w1), _p(b1), _p(w2), _p(b2), _p(y), B, 32, 16, T, None) != 0    # other widths refused


def test_conv_tc_dispatch_through_real_launchers_on_cpu(shim3):
    """Option "conv_tc": the C-ABI conv operators route wide layers through launch_conv1d_tc / launch_convtr1d_tc (host code of
    csrc/ua2_convtc.cu: scratch growth, chunking, raw-product epilogue) - here with the tensor-core GEMM replaced by a CPU GEMM
    (stubs_real_headers.cpp), so everything but the tcgen05 kernel itself is the shipped source.  Narrow layers must fall through."""
    g = torch.Generator().manual_seed(11)
    shim3.shim_set_conv_tc(1)
    try:
        for (B, Cin, Cout, T, K, stride, elu, res, rep) in [(2, 160, 24, 150, 8, 4, 1, 0, 0), (1, 352, 16, 140, 3, 1, 1, 1, 0), (1, 300, 8, 260, 4, 2, 0, 0, 1),
                                                            (1, 32, 16, 200, 3, 1, 1, 0, 0)]:  # the last one (Cin * K < 1024) stays on the SIMT path
            x = torch.randn(B, Cin, T, generator=g)
            w = torch.randn(Cout, Cin, K, generator=g) / math.sqrt(Cin * K)
            b = None if rep else torch.randn(Cout, generator=g) * 0.1
            ref = CO.conv1d_causal(F.elu(x) if elu else x, w, b, stride=stride, pad_mode="replicate" if rep else "constant")
            r = torch.randn_like(ref) if res else None
            if res:
                ref = r + ref
            y = torch.full_like(ref, float("nan"))
            _ok(shim3, shim3.ua2_conv1d_causal_gemm_f32(_p(x), _p(w), _p(b), _p(r), _p(y), B, Cin, Cout, T, K, stride, 1, elu, rep, None))
            assert float((y - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max())), (Cin, K)
        for (B, Cin, Cout, T, stride) in [(2, 160, 12, 70, 4), (1, 128, 8, 131, 8), (1, 64, 8, 150, 2)]:  # the last one (2 * Cin < 256) falls through
            x = torch.randn(B, Cin, T, generator=g)
            w = torch.randn(Cin, Cout, 2 * stride, generator=g) / math.sqrt(Cin * 2)
            b = torch.randn(Cout, generator=g) * 0.1
            ref = CO.convtr1d_causal(F.elu(x), w, b, stride)
            wp = torch.empty(stride * Cout * Cin * 2)
            _ok(shim3, shim3.ua2_convtr1d_repack_phase_f32(_p(w), _p(wp), Cin, Cout, stride, None))
            y = torch.full_like(ref, float("nan"))
            _ok(shim3, shim3.ua2_convtr1d_causal_gemm_f32(_p(x), _p(wp), _p(b), _p(y), B, Cin, Cout, T, stride, 1, None))
            assert float((y - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max())), (Cin, stride)
    finally:
        shim3.shim_set_conv_tc(0)


def _attention_reference(q, kc, vc, pos, bidx, n_head, n_groups, hs, window):
    """lit_model.py:468-532 for one query row per entry: keys 0..pos of cache row bidx (the last `window` of them when window > 0),
    each KV group serving n_head / n_groups query heads, softmax(q k / sqrt(hs)) v."""
    M, qpk = q.shape[0], n_head // n_groups
    out = torch.empty(M, n_head * hs)
    for m in range(M):
        n = int(pos[m]) + 1
        lo = max(0, n - window) if window > 0 else 0
        for h in range(n_head):
            k, v = kc[int(bidx[m]), h // qpk, lo:n].double(), vc[int(bidx[m]), h // qpk, lo:n].double()
            w = torch.softmax(k @ q[m, h * hs:(h + 1) * hs].double() / math.sqrt(hs), 0)
            out[m, h * hs:(h + 1) * hs] = (w @ v).float()
    return out


@pytest.mark.parametrize("hs,n_head,n_groups,M,S_max,window", [(128, 6, 2, 8, 600, 0), (64, 4, 4, 5, 500, 0), (32, 4, 2, 10, 520, 150)])
def test_kv_cache_attention_ring_and_split_sources_on_cpu(shim3, hs, n_head, n_groups, M, S_max, window):
    """csrc/ua2_attn.cu through launch_attn + launch_attn_combine: the shipped one-shot split kernel (GPU-green; here it validates the
    shim's mbarrier / bulk-copy emulation) and the not-yet-run persistent ring kernel (option "attn_ring") on ragged positions -
    rows that end inside a chunk, at a chunk edge, at the last cache slot, a single key; empty splits; permuted cache rows.  The two
    kernels share item arithmetic, so their outputs must be bit-equal; both must match the fp64 reference."""
    shim3.shim_set_sm_count(3)  # 6-12 persistent CTAs: a dozen or more items each, every ring slot reused with both parities
    g = torch.Generator().manual_seed(hs + M)
    B = M
    kc, vc = torch.randn(B, n_groups, S_max, hs, generator=g), torch.randn(B, n_groups, S_max, hs, generator=g)
    q = torch.randn(M, n_head * hs, generator=g)
    pos = torch.randint(0, S_max, (M,), generator=g).to(torch.int32)
    pos[0], pos[1], pos[2], pos[3] = S_max - 1, 0, 63, 64
    bidx = torch.randperm(B, generator=g).to(torch.int32)
    splits = (S_max + 63) // 64
    assert M * n_groups * splits >= 64  # the ring threshold of launch_attn in this build (-DUA2_ATTN_RING_MIN_ITEMS; 592 as shipped)
    ref = _attention_reference(q, kc, vc, pos, bidx, n_head, n_groups, hs, window)
    outs = []
    for ring in (0, 1):
        ws = torch.full((M * n_head * splits * (hs + 2),), float("nan"))
        y = torch.full((M, n_head * hs), float("nan"))
        grid_x = C.c_int(0)
        rc = shim3.shim_attn(_p(q), _p(kc), _p(vc), _p(pos), _p(bidx), _p(y), _p(ws), M, n_head, n_groups, hs, S_max, window, ring, C.byref(grid_x))
        assert rc == 0, rc
        assert grid_x.value == (3 * (2 if hs == 128 else 4) if ring else splits)  # persistent CTAs vs one CTA per item
        assert float((y - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max())), ring
        outs.append(y)
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("M,N,K,norm,res,swiglu", [(1, 384, 3072, True, False, True), (1, 320, 1024, True, False, False), (1, 256, 2048, False, True, False),
                                                   (2, 130, 256, False, False, False), (5, 258, 640, True, True, False), (11, 96, 132, False, False, True)])
def test_decode_linear_gemv3_source_on_cpu(shim3, M, N, K, norm, res, swiglu):
    """gemv3_kernel (csrc/ua2_gemv3.cu + ua2_gemv3_dev.cuh) - the weight-streaming linear of the decode frame and the kernel bench.py
    reports as dominant; GPU-green (tests/test_ops_gpu.py::test_linear / test_swiglu, the shapes and formulas used here).  On the
    shim it runs as written: persistent slab-partitioned CTAs, per-warp bulk-copy rings and mbarrier parities (emulated), K-split
    partial sums exchanged through shared memory, fused RMSNorm prologue / residual / SwiGLU epilogues, row tiles of 1, 2, 4 and 8."""
    from oracle import llm_oracle as O

    shim3.shim_set_sm_count(3)
    g = torch.Generator().manual_seed(M * 1000 + N + K)
    x = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / math.sqrt(K)
    W2 = torch.randn(N, K, generator=g) / math.sqrt(K) if swiglu else None
    nw = 1 + 0.1 * torch.randn(K, generator=g) if norm else None
    r = torch.randn(M, N, generator=g) if res else None
    xn = O.rms_norm(x, nw, 1e-5) if norm else x
    ref = F.silu(F.linear(xn, W)) * F.linear(xn, W2) if swiglu else F.linear(xn, W)
    if res:
        ref = ref + r
    y = torch.full((M, N), float("nan"))
    grid_y = C.c_int(0)
    rc = shim3.shim_gemv3(_p(x), _p(W), _p(W2), _p(nw), C.c_float(1e-5), _p(r), _p(y), M, N, K, C.byref(grid_y))
    assert rc == 0, rc
    assert 1 <= grid_y.value <= 9  # persistent grid: at most 3 CTAs on each of the 3 pretend SMs
    assert float((y - ref).abs().max() / ref.abs().max()) < 2e-5


@pytest.mark.skipif(os.environ.get("UA2_SHIM_FULL") != "1", reason="3 minutes of OS-thread emulation: UA2_SHIM_FULL=1 runs it (last run: passed)")
def test_whole_codec_handle_on_cpu_against_reference_golden(shim3):
    """csrc/ua2_codec_model.cu - the ua2_codec_* handle as shipped (weight-norm folding, layout repacks, SEANet encoder, projected
    transformer with its skinny LayerNorm / interleaved-RoPE / GELU / LayerScale linears and windowed attention, down-sampling,
    residual VQ, and the whole way back) - on CPU tensors through the shim, against tests/golden/codec_golden.pt: the fixtures the
    UNMODIFIED reference codec produced (oracle/make_golden_codec.py).  Same bar as the GPU test: bit-equal VQ indices, waveform
    within 1e-4.  The tensor-core GEMM is declared unavailable, so every linear runs on the fp32 kernels of csrc/."""
    from oracle.make_golden_codec import codec_cfgs
    from uniaudio2_b200 import _lib
    from uniaudio2_b200.tools.tokenizer.MimiCodec.mimi_codec import MimiCodec

    golden = torch.load(os.path.join(ROOT, "tests", "golden", "codec_golden.pt"), weights_only=False)
    cfg = codec_cfgs()["tiny"]
    sd = CO.random_mimi_state_dict(cfg, seed=4321)
    assert {k: float(v.double().sum()) for k, v in sd.items()} == golden["__checksum_tiny"]
    m = MimiCodec(sample_rate=cfg.sample_rate, n_filters=cfg.n_filters, encoder_rates=cfg.encoder_rates, compress=cfg.compress,
                  latent_dim=cfg.latent_dim, codebook_size=cfg.codebook_size, codebook_dim=cfg.codebook_dim, rvq_layers=cfg.rvq_layers,
                  num_heads=cfg.num_heads, num_layers=cfg.num_layers, layer_scale=cfg.layer_scale, context=cfg.context, device="cpu")
    full = m.state_dict()
    full.update(sd)
    m.load_state_dict(full, strict=True)
    shim3.shim_set_tc_available(0)
    shim3.shim_set_sm_count(4)
    try:
        ratios = (C.c_int32 * 8)(*(m.encoder_rates + [0] * (8 - len(m.encoder_rates))))
        c = m.cfg
        ccfg = _lib.CodecCfg(c["n_filters"], ratios, len(m.encoder_rates), c["latent_dim"], c["codebook_size"], c["codebook_dim"], c["rvq_layers"],
                             c["num_heads"], c["num_layers"], c["context"], c["dim_feedforward"], m.resample_stride, 10000.0)
        h = C.c_void_p()
        shim3.ua2_codec_frames.restype = C.c_int64
        _ok(shim3, shim3.ua2_codec_create(C.byref(ccfg), C.byref(h)))
        keep = []
        for key, t in m.state_dict().items():
            t = t.detach().contiguous()
            keep.append(t)
            _ok(shim3, shim3.ua2_codec_load_weight(h, key.encode(), _p(t), (C.c_int64 * t.dim())(*t.shape), t.dim()))
        _ok(shim3, shim3.ua2_codec_finalize(h, None))
        fx = golden["tiny_B2_T5797"]
        wav = fx["wav"].contiguous()
        B, _, T = wav.shape
        Tq = int(shim3.ua2_codec_frames(h, T))
        codes = torch.full((B, cfg.rvq_layers, Tq), -1, dtype=torch.int64)
        _ok(shim3, shim3.ua2_codec_encode(h, _p(wav), B, T, _p(codes), None))
        assert torch.equal(codes, fx["codes"]), "VQ indices differ from the reference"
        recon = torch.full((B, 1, Tq * m.resample_stride * m.hop_length), float("nan"))
        _ok(shim3, shim3.ua2_codec_decode(h, _p(fx["codes"].contiguous()), B, Tq, _p(recon), None))
        assert recon.shape == fx["recon"].shape
        assert float((recon - fx["recon"]).abs().max()) < 1e-4
        shim3.ua2_codec_destroy(h)
    finally:
        shim3.shim_set_tc_available(1)
