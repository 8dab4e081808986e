"""Profiling driver (run under ncu): a few launches of the hand-written tcgen05 GEMM at a decode shape (32 rows against the
128256 x 3072 text head: HBM-bound weight streaming) and a prefill shape (1024 rows, SwiGLU pair: tensor-bound).  Not a benchmark.

    ncu --set full --clock-control none --import-source on -k regex:umma_kernel -s 2 -c 4 -o gpurun_out/umma python tools/profile_umma.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from uniaudio2_b200 import _lib  # noqa: E402

L, P = _lib.lib(), _lib.ptr
dev = torch.device("cuda", 0)
torch.manual_seed(0)
for M, N, K, sw in [(32, 128256, 3072, 0), (1024, 8192, 3072, 1)]:
    x = torch.randn(M, K, device=dev)
    W = torch.randn(N, K, device=dev) / K ** 0.5
    W2 = torch.randn(N, K, device=dev) / K ** 0.5 if sw else None
    y = torch.empty(M, N, device=dev)
    for _ in range(3):
        _lib.check(L.ua2_tc_linear_f32(P(x), P(W), P(W2), None, 1e-5, None, P(y), M, N, K, None))
    torch.cuda.synchronize()
print("done")
