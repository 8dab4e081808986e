"""Profiling driver (run under ncu): the KV-cache attention at batch 32 x 2048 keys (Llama-3.2-3B geometry) and the SEANet layers at
batch 16 x 10 s (Mimi geometry): fused 64-channel residual block, strided conv 64 -> 128 (k 8, s 4), transposed conv 128 -> 64.
Not a benchmark.

    ncu --set full --clock-control none --import-source on -k regex:"attn_split_kernel|attn_ring_kernel|resblock64_kernel|sgemm_conv_kernel|umma_kernel" \
        -o gpurun_out/attn_conv python tools/profile_attn_conv.py
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
# ---- attention: one-shot split kernel at 2048 keys, ring kernel at 540 keys
n_head, G, hs, S_max = 24, 8, 128, 2048
for B, S in ((32, 2048), (32, 540)):
    kc, vc = torch.randn(B, G, S_max, hs, device=dev), torch.randn(B, G, S_max, hs, device=dev)
    q = torch.randn(B, n_head * hs, device=dev)
    pos = torch.full((B,), S - 1, dtype=torch.int32, device=dev)
    bidx = torch.arange(B, dtype=torch.int32, device=dev)
    y = torch.empty(B, n_head * hs, device=dev)
    ws = torch.empty(L.ua2_attn_workspace_floats(B, n_head, hs, S_max), device=dev)
    for _ in range(2):
        _lib.check(L.ua2_attn_f32(P(q), P(kc), P(vc), P(pos), P(bidx), P(y), P(ws), B, n_head, G, hs, S_max, None))
    torch.cuda.synchronize()
    del kc, vc
# ---- SEANet layers at 24 kHz, batch 16 x 10 s
Bc, T = 16, 240000
x = torch.randn(Bc, 64, T, device=dev)
w1 = torch.randn(32, 64, 3, device=dev) / (64 * 3) ** 0.5
w2 = torch.randn(64, 32, 1, device=dev) / 32 ** 0.5
b1, b2 = torch.zeros(32, device=dev), torch.zeros(64, device=dev)
yb = torch.empty(Bc, 64, T, device=dev)
for _ in range(2):
    _lib.check(L.ua2_resblock_f32(P(x), P(w1), P(b1), P(w2), P(b2), P(yb), Bc, 64, 32, T, None))
w = torch.randn(128, 64, 8, device=dev) / (64 * 8) ** 0.5
b = torch.zeros(128, device=dev)
yd = torch.empty(Bc, 128, T // 4, device=dev)
for _ in range(2):
    _lib.check(L.ua2_conv1d_causal_gemm_f32(P(x), P(w), P(b), None, P(yd), Bc, 64, 128, T, 8, 4, 1, 1, 0, None))
wt = torch.randn(128, 64, 8, device=dev) / (128 * 2) ** 0.5
wp = torch.empty(4, 64, 128, 2, device=dev)
_lib.check(L.ua2_convtr1d_repack_phase_f32(P(wt), P(wp), 128, 64, 4, None))
bt = torch.zeros(64, device=dev)
for _ in range(2):
    _lib.check(L.ua2_convtr1d_causal_gemm_f32(P(yd), P(wp), P(bt), P(yb), Bc, 128, 64, T // 4, 4, 1, None))
torch.cuda.synchronize()
print("done")
