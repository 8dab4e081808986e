"""Hardware check + timing of the hand-written tcgen05 3xTF32 GEMM (csrc/ua2_umma.cu) through ua2_tc_linear_f32.

    python tools/check_umma.py [--time]

Correctness: against an fp64 product of the same fp32 operands; bar = the fp32-class error model of the path (3xTF32 drops
lo*lo ~ 2^-22; the TMEM accumulator rounds toward zero once per k-step of 8: about steps/2 ulp of the running sum).
Timing (--time): decode-shaped (M = 32) weight streaming against the HBM peak and prefill-shaped (M = 1024 .. 6016) against the
tf32 tensor peak / 3, next to the library collective of round 1 (option tc_impl = 0) where it is built."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from uniaudio2_b200 import _lib  # noqa: E402

L = _lib.lib()
P = _lib.ptr


def run(x, W, W2=None, norm_w=None, res=None):
    M, K = x.shape
    N = W.shape[0]
    y = torch.full((M, N), float("nan"), device=x.device)
    _lib.check(L.ua2_tc_linear_f32(P(x), P(W), P(W2) if W2 is not None else None, P(norm_w) if norm_w is not None else None, 1e-5,
                                   P(res) if res is not None else None, P(y), M, N, K, None))
    return y


def check():
    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cpu").manual_seed(0)
    bad = 0
    cases = [  # (M, N, K, swiglu, rmsnorm, residual)
        (32, 256, 64, 0, 0, 0), (32, 128, 32, 0, 0, 0), (1, 128, 128, 0, 0, 0), (7, 260, 100, 0, 0, 0), (32, 5120, 3072, 0, 1, 0),
        (32, 3072, 8192, 0, 0, 1), (32, 8192, 3072, 1, 1, 0), (64, 1024, 512, 0, 0, 0), (50, 12300, 2048, 0, 0, 0),
        (128, 768, 768, 0, 1, 1), (147, 2304, 768, 0, 0, 0), (300, 512, 2048, 1, 0, 0), (1000, 1536, 1040, 0, 0, 0),
        (1024, 5120, 3072, 0, 1, 0), (513, 1344, 512, 1, 0, 0),
    ]
    for M, N, K, sw, rn, rs in cases:
        x = torch.randn(M, K, generator=g)
        W = torch.randn(N, K, generator=g) / K ** 0.5
        W2 = torch.randn(N, K, generator=g) / K ** 0.5 if sw else None
        nw = torch.rand(K, generator=g) + 0.5 if rn else None
        r = torch.randn(M, N, generator=g) if rs else None
        xd = x.double()
        if rn:
            xn = (x.float() * torch.rsqrt((x.float() ** 2).mean(-1, keepdim=True) + 1e-5) * nw).double()
        else:
            xn = xd
        ref = xn @ W.double().t()
        if sw:
            ref = torch.nn.functional.silu(ref) * (xn @ W2.double().t())
        if rs:
            ref = ref + r.double()
        y = run(x.to(dev), W.to(dev), W2.to(dev) if sw else None, nw.to(dev) if rn else None, r.to(dev) if rs else None)
        torch.cuda.synchronize()
        err = float((y.cpu().double() - ref).abs().max())
        scale = max(1.0, float(ref.abs().max()))
        tol = max(4e-6, 1.5 * (3 * K / 8) * 2.0 ** -24) * (2.0 if sw else 1.0)
        ok = err <= tol * scale and bool(torch.isfinite(y).all())
        bad += 0 if ok else 1
        print(json.dumps(dict(M=M, N=N, K=K, swiglu=sw, rmsnorm=rn, residual=rs, max_abs_err=err, rel=err / scale, tol=tol, ok=ok)), flush=True)
    return bad


def timed(fn, reps):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def timing():
    dev = torch.device("cuda", 0)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = float(peaks.get("hbm_gbs", 6534.5))
    for M, N, K, sw in [(32, 5120, 3072, 0), (32, 8192, 3072, 1), (32, 3072, 8192, 0), (32, 128256, 3072, 0), (64, 8192, 3072, 1),
                        (128, 8192, 3072, 1), (1024, 5120, 3072, 0), (1024, 8192, 3072, 1), (1024, 3072, 8192, 0), (6016, 8192, 3072, 1),
                        (1000, 4608, 1536, 0), (1000, 6144, 1536, 0), (1000, 1536, 6144, 0)]:
        n_sets = max(1, int(400e6 // (N * K * 4 * (2 if sw else 1))) + 1) if M <= 128 else 1  # rotate weights past L2 for the HBM-bound shapes
        Ws = [torch.randn(N, K, device=dev) / K ** 0.5 for _ in range(n_sets)]
        W2s = [torch.randn(N, K, device=dev) / K ** 0.5 for _ in range(n_sets)] if sw else None
        x = torch.randn(M, K, device=dev)
        y = torch.empty(M, N, device=dev)
        i = [0]

        def fn():
            j = i[0] % n_sets
            i[0] += 1
            _lib.check(L.ua2_tc_linear_f32(P(x), P(Ws[j]), P(W2s[j]) if sw else None, None, 1e-5, None, P(y), M, N, K, None))

        ms = timed(fn, 20)
        wbytes = N * K * 4 * (2 if sw else 1)
        flops = 2.0 * M * N * K * (2 if sw else 1)
        print(json.dumps(dict(M=M, N=N, K=K, swiglu=sw, ms=round(ms, 4), weight_GBps=round(wbytes / ms / 1e6, 1), frac_hbm=round(wbytes / ms / 1e6 / hbm, 3),
                              fp32_equiv_TFLOPs=round(flops / ms / 1e9, 1), tf32_mma_TFLOPs=round(3 * flops / ms / 1e9, 1),
                              note="includes the activation-split and epilogue kernels (3 launches)")), flush=True)
        del Ws, W2s
        torch.cuda.empty_cache()


if __name__ == "__main__":
    bad = check()
    print("FAILED cases:", bad)
    if "--time" in sys.argv:
        timing()
    sys.exit(1 if bad else 0)
