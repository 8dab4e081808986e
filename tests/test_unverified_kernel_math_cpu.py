"""CPU suite: index arithmetic of the kernels that were written after round 1's GPU budget ran out (csrc/ua2_convtc.cu,
csrc/ua2_resblock.cu), transcribed statement by statement into numpy and checked against the conv oracle.  This does not
run the kernels - it checks that the gather / scatter formulas they implement compute the convolution they claim to, which is
where an unrun kernel most plausibly goes wrong.  The GPU tests of the same code are opt-in (tests/test_zzz_unverified_gpu.py)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import codec_oracle as CO


def _elu(v):
    return np.where(v > 0, v, np.expm1(v))


def emulate_conv_tc(x, w, bias, res, stride, dilation, pre_elu, replicate, chunk_rows):
    """launch_conv1d_tc: conv_im2col_kernel -> GEMM with W = (Cout, Cin*K) -> conv_tc_epilogue_kernel, in row chunks."""
    B, Cin, T_in = x.shape
    Cout, _, Ktaps = w.shape
    k_eff = (Ktaps - 1) * dilation + 1
    pad_left = k_eff - stride                      # ua2_conv1d_causal_gemm_f32
    T_out = (T_in + stride - 1) // stride
    KT = Cin * Ktaps
    M = B * T_out
    W = w.reshape(Cout, KT)
    y = np.full((B, Cout, T_out), np.nan, dtype=np.float64)
    for m0 in range(0, M, chunk_rows):
        rows = min(chunk_rows, M - m0)
        A = np.zeros((rows, KT))
        for i in range(rows * KT):                 # conv_im2col_kernel, one "thread" per element
            k = i % KT
            m = m0 + i // KT
            b, t = divmod(m, T_out)
            ci, tap = divmod(k, Ktaps)
            pos = t * stride - pad_left + tap * dilation
            inside = 0 <= pos < T_in
            if not inside and replicate:
                pos = 0 if pos < 0 else T_in - 1
                inside = True
            v = 0.0
            if inside:
                v = x[b, ci, pos]
                if pre_elu:
                    v = float(_elu(np.float64(v)))
            A[i // KT, k] = v
        C = A @ W.T
        for r in range(rows):                      # conv_tc_epilogue_kernel
            m = m0 + r
            b, t = divmod(m, T_out)
            for co in range(Cout):
                v = C[r, co] + (bias[co] if bias is not None else 0.0)
                if res is not None:
                    v += res[b, co, t]
                y[b, co, t] = v
    return y


@pytest.mark.parametrize("B,Cin,Cout,T,K,stride,dil,elu,use_res,rep", [
    (2, 6, 5, 23, 8, 4, 1, 1, 0, 0), (1, 4, 7, 17, 3, 1, 1, 1, 1, 0), (2, 5, 5, 19, 4, 2, 1, 0, 0, 1), (1, 3, 4, 30, 3, 1, 2, 1, 0, 0)])
def test_conv_tc_index_math(B, Cin, Cout, T, K, stride, dil, elu, use_res, rep):
    g = torch.Generator().manual_seed(T + K)
    x = torch.randn(B, Cin, T, generator=g, dtype=torch.float64)
    w = torch.randn(Cout, Cin, K, generator=g, dtype=torch.float64) / math.sqrt(Cin * K)
    b = None if rep else torch.randn(Cout, generator=g, dtype=torch.float64)
    ref = CO.conv1d_causal(F.elu(x) if elu else x, w, b, stride=stride, dilation=dil, pad_mode="replicate" if rep else "constant")
    r = torch.randn_like(ref) if use_res else None
    if use_res:
        ref = r + ref
    got = emulate_conv_tc(x.numpy(), w.numpy(), None if b is None else b.numpy(), None if r is None else r.numpy(), stride, dil, elu, rep,
                          chunk_rows=7)
    assert got.shape == tuple(ref.shape) and not np.isnan(got).any()
    assert np.abs(got - ref.numpy()).max() < 1e-10


def emulate_convtr_tc(x, w, bias, stride, pre_elu, chunk_rows, crop_left=0, T_out=None):
    """launch_convtr1d_tc: convtr_im2col_kernel -> GEMM with W = w_phase (s*Cout, Cin*2) -> convtr_tc_epilogue_kernel."""
    B, Cin, T_in = x.shape
    _, Cout, K2 = w.shape
    s = stride
    assert K2 == 2 * s
    T_out = T_in * s if T_out is None else T_out
    # ua2_convtr1d_repack_phase_f32: w_phase[ph][co][ci][tap] = w[ci][co][ph + tap * s]
    wp = np.zeros((s, Cout, Cin, 2))
    for ph in range(s):
        for tap in range(2):
            wp[ph, :, :, tap] = w[:, :, ph + tap * s].T
    W = wp.reshape(s * Cout, Cin * 2)
    Tj = T_in + 1 if crop_left + T_out > T_in * s else T_in
    M = B * Tj
    y = np.full((B, Cout, T_out), np.nan)
    for m0 in range(0, M, chunk_rows):
        rows = min(chunk_rows, M - m0)
        A = np.zeros((rows, 2 * Cin))
        for r in range(rows):                      # convtr_im2col_kernel
            b, j = divmod(m0 + r, Tj)
            for ci in range(Cin):
                a = x[b, ci, j] if j < T_in else 0.0
                bb = x[b, ci, j - 1] if (j >= 1 and j - 1 < T_in) else 0.0
                if pre_elu:
                    a, bb = float(_elu(np.float64(a))), float(_elu(np.float64(bb)))
                A[r, 2 * ci], A[r, 2 * ci + 1] = a, bb
        C = A @ W.T
        for r in range(rows):                      # convtr_tc_epilogue_kernel
            b, j = divmod(m0 + r, Tj)
            for ph in range(s):
                t = j * s + ph - crop_left
                if 0 <= t < T_out:
                    y[b, :, t] = C[r, ph * Cout:(ph + 1) * Cout] + (bias if bias is not None else 0.0)
    return y


@pytest.mark.parametrize("B,Cin,Cout,T,stride", [(2, 5, 3, 9, 4), (1, 4, 6, 13, 2), (1, 3, 2, 7, 5)])
def test_convtr_tc_index_math(B, Cin, Cout, T, stride):
    g = torch.Generator().manual_seed(T + stride)
    x = torch.randn(B, Cin, T, generator=g, dtype=torch.float64)
    w = torch.randn(Cin, Cout, 2 * stride, generator=g, dtype=torch.float64) / math.sqrt(2 * Cin)
    b = torch.randn(Cout, generator=g, dtype=torch.float64)
    ref = CO.convtr1d_causal(F.elu(x), w, b, stride)
    got = emulate_convtr_tc(x.numpy(), w.numpy(), b.numpy(), stride, 1, chunk_rows=5)
    assert got.shape == tuple(ref.shape) and not np.isnan(got).any()
    assert np.abs(got - ref.numpy()).max() < 1e-10


def emulate_resblock64(x, w1, b1, w2, b2, RB_T=128):
    """resblock64_kernel: per CTA (clip b, tile t0) - ELU(x) tile with two left halo columns, stage 1 (thread = 2 hidden x 8
    positions reading columns p + tap), ELU, stage 2 (thread = 4 channels x 8 positions), skip from the raw input."""
    B, C, T = x.shape
    H = w1.shape[0]
    y = np.full_like(x, np.nan)
    for b in range(B):
        for t0 in range(0, T, RB_T):
            xe = np.zeros((C, RB_T + 2))
            for ci in range(C):
                for j in range(RB_T + 2):
                    t = t0 - 2 + j
                    xe[ci, j] = _elu(x[b, ci, t]) if 0 <= t < T else 0.0
            he = np.zeros((H, RB_T))
            for th in range(16):
                for tt in range(16):
                    for r in range(2):
                        h = 2 * th + r
                        for p in range(8):
                            acc = b1[h]
                            for ci in range(C):
                                for tap in range(3):
                                    acc += w1[h, ci, tap] * xe[ci, 8 * tt + p + tap]
                            he[h, 8 * tt + p] = _elu(acc)
            for tc in range(16):
                for tt in range(16):
                    for c in range(4):
                        ch = 4 * tc + c
                        for p in range(8):
                            t = t0 + 8 * tt + p
                            if t < T:
                                y[b, ch, t] = x[b, ch, t] + b2[ch] + float(w2[ch, :, 0] @ he[:, 8 * tt + p])
    return y


def test_resblock_fused_index_math():
    g = torch.Generator().manual_seed(1)
    B, C, H, T = 1, 64, 32, 150  # two tiles, the second one ragged
    x = torch.randn(B, C, T, generator=g, dtype=torch.float64)
    w1 = torch.randn(H, C, 3, generator=g, dtype=torch.float64) / math.sqrt(C * 3)
    b1 = torch.randn(H, generator=g, dtype=torch.float64) * 0.1
    w2 = torch.randn(C, H, 1, generator=g, dtype=torch.float64) / math.sqrt(H)
    b2 = torch.randn(C, generator=g, dtype=torch.float64) * 0.1
    ref = x + CO.conv1d_causal(F.elu(CO.conv1d_causal(F.elu(x), w1, b1)), w2, b2)
    got = emulate_resblock64(x.numpy(), w1.numpy(), b1.numpy(), w2.numpy(), b2.numpy())
    assert not np.isnan(got).any()
    assert np.abs(got - ref.numpy()).max() < 1e-10
