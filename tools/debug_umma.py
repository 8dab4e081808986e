"""Error map of ua2_tc_linear_f32 per (row tile, weight tile) for a few shapes (debug aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from uniaudio2_b200 import _lib
L, P = _lib.lib(), _lib.ptr
for M, N, K, NT in [(300, 512, 2048, 128), (1024, 5120, 3072, 256), (513, 1344, 512, 128), (256, 256, 2048, 256), (512, 256, 2048, 256), (300, 256, 512, 128)]:
    g = torch.Generator().manual_seed(1)
    x = torch.randn(M, K, generator=g); W = torch.randn(N, K, generator=g) / K ** 0.5
    ref = x.double() @ W.double().t()
    xd, Wd = x.cuda(), W.cuda()
    y = torch.full((M, N), float("nan"), device="cuda")
    _lib.check(L.ua2_tc_linear_f32(P(xd), P(Wd), None, None, 1e-5, None, P(y), M, N, K, None))
    torch.cuda.synchronize()
    err = (y.cpu().double() - ref).abs()
    n_mt, n_nt = (M + NT - 1) // NT, (N + 127) // 128
    print(f"shape M={M} N={N} K={K}: max err {float(err.max()):.3e}")
    for mt in range(n_mt):
        row = []
        for nt in range(min(n_nt, 24)):
            e = err[mt * NT:(mt + 1) * NT, nt * 128:(nt + 1) * 128]
            row.append("%.0e" % float(e.max()))
        print("  mt", mt, " ".join(row))
    bad = (err > 1e-2).nonzero()
    if len(bad):
        print("  first bad (m, n):", bad[:6].tolist(), " bad count", len(bad), " bad rows", sorted(set((bad[:, 0] // 32).tolist()))[:20])
