"""CPU suite: the SOURCE of the kernels that have not run on a GPU yet (csrc/ua2_convtc.cu, csrc/ua2_resblock.cu) compiled
with g++ against a thread-per-CUDA-thread shim (tests/cpu_shim/: OS threads, a barrier for __syncthreads, static __shared__)
and executed through the product's own launchers, with the tensor-core GEMM replaced by a CPU GEMM.  Checks the kernels as
written - shared-memory indexing, barriers, halo handling, chunk loops - against the conv oracle.  (Warp shuffles, bulk copies
and tensor cores are outside the shim: kernels that use them are covered on the GPU only.)"""
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
    cmd = ["g++", "-std=c++20", "-O1", "-shared", "-fPIC", "-pthread", "-I", d, "-I", SHIM, os.path.join(SHIM, "harness.cpp"), "-o", so]
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
