"""Debug driver: ua2_conv1d_f32 against torch for a list of shapes / options; prints the max error of each."""
import math
import sys

sys.path.insert(0, ".")
import torch
import torch.nn.functional as F

from uniaudio2_b200 import _lib

L = _lib.lib()
staged_modes = [int(a) for a in sys.argv[1:]] or [0, 1]
cases = [  # B, Cin, Cout, T, K, dil, causal, res, prelu
    (3, 64, 32, 1000, 3, 1, 1, 1, 1), (1, 96, 96, 3000, 7, 9, 1, 0, 1), (2, 64, 64, 1028, 2, 1, 1, 0, 0),
]
for staged in staged_modes:
    _lib.check(L.ua2_set_global_option(b"conv_umma_staged", staged))
    for rep, (B, Cin, Cout, T, K, dil, causal, res, pre) in enumerate(cases):
        g = torch.Generator().manual_seed(Cin + Cout + T + K + dil)
        x = torch.randn(B, Cin, T, generator=g)
        w = torch.randn(Cout, Cin, K, generator=g) / math.sqrt(Cin * K)
        b = torch.randn(Cout, generator=g) * 0.1
        slope = torch.tensor([0.25])
        pl, pr = (dil * (K - 1), 0) if causal else ((K * dil - dil) // 2,) * 2
        ref = F.conv1d(F.pad(x, (pl, pr)), w, b, dilation=dil)
        if pre:
            ref = F.prelu(ref, slope)
        r = torch.randn_like(ref) if res else None
        if res:
            ref = ref + r
        xd, wd, bd, sd = x.cuda(), w.contiguous().cuda(), b.cuda(), slope.cuda()
        rd = r.cuda() if res else None
        y = torch.full(tuple(ref.shape), float("nan"), device="cuda")
        _lib.check(L.ua2_conv1d_f32(_lib.ptr(xd), _lib.ptr(wd), _lib.ptr(bd), _lib.ptr(sd) if pre else None, _lib.ptr(rd), _lib.ptr(y), B, Cin, Cout, T, K, 1,
                                    dil, pl, pr, None))
        torch.cuda.synchronize()
        d = (y.cpu() - ref).abs()
        bad = (d > 1e-3).nonzero()
        print("staged", staged, (B, Cin, Cout, T, K, dil, causal, res, pre), "err", float(d.max()), "bad", int(bad.shape[0]), bad[:3].tolist(), flush=True)
