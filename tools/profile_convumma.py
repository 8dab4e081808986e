"""Profiling driver (run under ncu): the tcgen05 implicit-GEMM convolution (csrc/ua2_convumma.cu) on the 24 kHz SEANet layers at batch 16 x 10 s."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from uniaudio2_b200 import _lib  # noqa: E402

L, P = _lib.lib(), _lib.ptr
dev = torch.device("cuda", 0)
torch.manual_seed(0)
Bc, T = 16, 240000
x = torch.randn(Bc, 64, T, device=dev)
hid = torch.empty(Bc, 32, T, device=dev)
yb = torch.empty(Bc, 64, T, device=dev)
w1 = torch.randn(32, 64, 3, device=dev) / (64 * 3) ** 0.5
w2 = torch.randn(64, 32, 1, device=dev) / 32 ** 0.5
b1, b2 = torch.zeros(32, device=dev), torch.zeros(64, device=dev)
for _ in range(2):
    _lib.check(L.ua2_conv1d_causal_gemm_f32(P(x), P(w1), P(b1), None, P(hid), Bc, 64, 32, T, 3, 1, 1, 1, 0, None))
    _lib.check(L.ua2_conv1d_causal_gemm_f32(P(hid), P(w2), P(b2), P(x), P(yb), Bc, 32, 64, T, 1, 1, 1, 1, 0, None))
torch.cuda.synchronize()
print("done")
