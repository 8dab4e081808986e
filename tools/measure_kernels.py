"""Roofline of the two kernel families BASELINE.json's north_star singles out - the KV-cache attention and the SEANet
convolutions - measured stand-alone through the C ABI on one B200 (CUDA events on the launching stream, 3 warm-ups, mean of
`reps` launches; operands larger than L2 or rotated through a set of buffers larger than L2).

  attention : ua2_attn_f32 on decode rows (one query per sequence) of the Llama-3.2-3B geometry (24 heads / 8 KV groups x 128);
              algorithmic bytes = 2 * G * S * hs * 4 per sequence (K and V read once), bound = HBM
  conv      : ua2_conv1d_causal_gemm_f32 / ua2_convtr1d_causal_gemm_f32 on the SEANet layers of the Mimi geometry at batch 16 x
              10 s; FLOP = 2 * B * T_out * Cout * Cin * K; bound = fp32 FMA pipe (148 SMs x 128 lanes x 2 x SM clock) for all but
              the first / last layer, whose intensity is low enough for HBM to bind

    python tools/measure_kernels.py [--conv-tc] [--resblock] [--attn-ring]     (also time the not-yet-measured options)
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from uniaudio2_b200 import _lib  # noqa: E402

L = _lib.lib()
P = _lib.ptr


def timed(fn, reps):
    for _ in range(3):
        fn(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = float(peaks.get("hbm_gbs", 6534.5))
    fp32_peak = 148 * 128 * 2 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12  # TFLOP/s
    n_head, G, hs, S_max = 24, 8, 128, 2048
    ring_opts = (0, 1) if "--attn-ring" in sys.argv else (0,)
    for B, S in ((1, 256), (1, 2048), (8, 2048), (32, 2048), (32, 540)):
        kv_bytes = 2 * G * S_max * hs * 4 * B
        n_sets = max(1, int(300e6 // kv_bytes) + 1)  # rotate through > 126 MB of caches so that K/V come from HBM
        kcs = [torch.randn(B, G, S_max, hs, device=dev) for _ in range(n_sets)]
        vcs = [torch.randn(B, G, S_max, hs, device=dev) for _ in range(n_sets)]
        q = torch.randn(B, n_head * hs, device=dev)
        pos = torch.full((B,), S - 1, dtype=torch.int32, device=dev)
        bidx = torch.arange(B, dtype=torch.int32, device=dev)
        y = torch.empty(B, n_head * hs, device=dev)
        ws = torch.empty(L.ua2_attn_workspace_floats(B, n_head, hs, S_max), device=dev)
        by = 2 * G * S * hs * 4 * B
        for ring in ring_opts:
            _lib.check(L.ua2_set_global_option(b"attn_ring", ring))

            def run(i, stream=None):
                j = i % n_sets
                _lib.check(L.ua2_attn_f32(P(q), P(kcs[j]), P(vcs[j]), P(pos), P(bidx), P(y), P(ws), B, n_head, G, hs, S_max, stream))

            ms = timed(run, 50)
            # the same calls replayed from a CUDA graph: without the host cost of the two ctypes launches per call (~50 us), which
            # bounds the eager number for every shape but the largest (profiles/r1_kernel_rooflines.md)
            side = torch.cuda.Stream()
            n_calls = max(n_sets, 8)
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(graph, stream=side):
                for i in range(n_calls):
                    run(i, C.c_void_p(side.cuda_stream))
            ms_graph = timed(lambda i: graph.replay(), 10) / n_calls
            print(json.dumps(dict(kernel="ua2_attn_f32: %s + combine" % ("attn_ring_kernel<128> (where the launcher takes it)" if ring else "attn_split_kernel<128>"),
                                  attn_ring=ring, batch=B, keys=S, us=round(ms * 1e3, 2), us_graph_replay=round(ms_graph * 1e3, 2),
                                  algorithmic_MB=round(by / 1e6, 2), GBps_graph=round(by / ms_graph / 1e6, 1),
                                  frac_of_hbm_peak_graph=round(by / ms_graph / 1e6 / hbm, 3), frac_of_hbm_peak_eager=round(by / ms / 1e6 / hbm, 3),
                                  buffers_rotated=n_sets)))
            del graph
        _lib.check(L.ua2_set_global_option(b"attn_ring", 0))
        del kcs, vcs
        torch.cuda.empty_cache()
    # ---- SEANet layers (encoder: conv; decoder: transposed conv), batch 16 x 10 s at 24 kHz
    Bc, T0 = 16, 240000
    convs = [  # (name, Cin, Cout, K, stride, T_in)
        ("enc conv k7 1->64", 1, 64, 7, 1, T0),
        ("enc res k3 64->32", 64, 32, 3, 1, T0),
        ("enc res k1 32->64", 32, 64, 1, 1, T0),
        ("enc down k8 s4 64->128", 64, 128, 8, 4, T0),
        ("enc res k3 128->64 @6k", 128, 64, 3, 1, T0 // 4),
        ("enc res k1 64->128 @6k", 64, 128, 1, 1, T0 // 4),
        ("enc down k10 s5 128->256", 128, 256, 10, 5, T0 // 4),
        ("enc down k12 s6 256->512", 256, 512, 12, 6, T0 // 20),
        ("enc down k16 s8 512->1024", 512, 1024, 16, 8, T0 // 120),
        ("enc last k3 1024->512", 1024, 512, 3, 1, T0 // 960),
    ]
    conv_opts = (0, 1) if "--conv-tc" in sys.argv else (0,)
    for name, Cin, Cout, K, stride, T, conv_tc in [c + (o,) for c in convs for o in conv_opts]:
        _lib.check(L.ua2_set_global_option(b"conv_tc", conv_tc))
        x = torch.randn(Bc, Cin, T, device=dev)
        w = torch.randn(Cout, Cin, K, device=dev) / (Cin * K) ** 0.5
        b = torch.zeros(Cout, device=dev)
        T_out = -(-T // stride)
        yb = torch.empty(Bc, Cout, T_out, device=dev)

        def run(i):
            _lib.check(L.ua2_conv1d_causal_gemm_f32(P(x), P(w), P(b), None, P(yb), Bc, Cin, Cout, T, K, stride, 1, 1, 0, None))

        ms = timed(run, 10)
        fl = 2.0 * Bc * T_out * Cout * Cin * K
        by = 4.0 * (x.numel() + yb.numel() + w.numel())
        t_roof = max(by / (hbm * 1e9), fl / (fp32_peak * 1e12)) * 1e3
        print(json.dumps(dict(kernel="sgemm_conv_kernel (ua2_conv1d_causal_gemm_f32)", layer=name, conv_tc=conv_tc, ms=round(ms, 3), GFLOP=round(fl / 1e9, 1),
                              TFLOPs=round(fl / ms / 1e9, 1), MB=round(by / 1e6, 1), GBps=round(by / ms / 1e6, 1),
                              flop_per_byte=round(fl / by, 1), bound="hbm" if by / (hbm * 1e9) > fl / (fp32_peak * 1e12) else "fp32",
                              frac_of_roofline=round(t_roof / ms, 3))))
    _lib.check(L.ua2_set_global_option(b"conv_tc", 0))
    for name, Cin, Cout, stride, T, conv_tc in [c + (o,) for c in (("dec up k16 s8 1024->512", 1024, 512, 8, T0 // 960),
                                                                     ("dec up k12 s6 512->256", 512, 256, 6, T0 // 120),
                                                                     ("dec up k10 s5 256->128", 256, 128, 5, T0 // 20),
                                                                     ("dec up k8 s4 128->64", 128, 64, 4, T0 // 4)) for o in conv_opts]:
        _lib.check(L.ua2_set_global_option(b"conv_tc", conv_tc))
        x = torch.randn(Bc, Cin, T, device=dev)
        w = torch.randn(Cin, Cout, 2 * stride, device=dev) / (Cin * 2) ** 0.5
        wp = torch.empty(stride, Cout, Cin, 2, device=dev)
        _lib.check(L.ua2_convtr1d_repack_phase_f32(P(w), P(wp), Cin, Cout, stride, None))
        b = torch.zeros(Cout, device=dev)
        yb = torch.empty(Bc, Cout, T * stride, device=dev)

        def run(i):
            _lib.check(L.ua2_convtr1d_causal_gemm_f32(P(x), P(wp), P(b), P(yb), Bc, Cin, Cout, T, stride, 1, None))

        ms = timed(run, 10)
        fl = 2.0 * Bc * T * stride * Cout * Cin * 2
        by = 4.0 * (x.numel() + yb.numel() + w.numel())
        t_roof = max(by / (hbm * 1e9), fl / (fp32_peak * 1e12)) * 1e3
        print(json.dumps(dict(kernel="sgemm_conv_kernel phase GEMMs (ua2_convtr1d_causal_gemm_f32)", layer=name, conv_tc=conv_tc, ms=round(ms, 3),
                              GFLOP=round(fl / 1e9, 1), TFLOPs=round(fl / ms / 1e9, 1), GBps=round(by / ms / 1e6, 1),
                              flop_per_byte=round(fl / by, 1), bound="hbm" if by / (hbm * 1e9) > fl / (fp32_peak * 1e12) else "fp32",
                              frac_of_roofline=round(t_roof / ms, 3))))
    _lib.check(L.ua2_set_global_option(b"conv_tc", 0))
    if "--resblock" in sys.argv:  # the 64-channel residual block at 24 kHz: two implicit GEMMs vs the fused kernel
        Cc, H, T = 64, 32, T0
        x = torch.randn(Bc, Cc, T, device=dev)
        w1 = torch.randn(H, Cc, 3, device=dev) / (Cc * 3) ** 0.5
        w2 = torch.randn(Cc, H, 1, device=dev) / H ** 0.5
        b1, b2 = torch.zeros(H, device=dev), torch.zeros(Cc, device=dev)
        hid, yb = torch.empty(Bc, H, T, device=dev), torch.empty(Bc, Cc, T, device=dev)

        def two(i):
            _lib.check(L.ua2_conv1d_causal_gemm_f32(P(x), P(w1), P(b1), None, P(hid), Bc, Cc, H, T, 3, 1, 1, 1, 0, None))
            _lib.check(L.ua2_conv1d_causal_gemm_f32(P(hid), P(w2), P(b2), P(x), P(yb), Bc, H, Cc, T, 1, 1, 1, 1, 0, None))

        def fused(i):
            _lib.check(L.ua2_resblock_f32(P(x), P(w1), P(b1), P(w2), P(b2), P(yb), Bc, Cc, H, T, None))

        fl = 2.0 * Bc * T * (Cc * H * 3 + H * Cc)
        for name, fn in (("two implicit GEMMs", two), ("fused (ua2_resblock_f32)", fused)):
            ms = timed(fn, 10)
            print(json.dumps(dict(kernel="SEANet resblock 64 -> 32 -> 64 @ 24 kHz", impl=name, ms=round(ms, 3), GFLOP=round(fl / 1e9, 1),
                                  TFLOPs=round(fl / ms / 1e9, 1), frac_of_fp32_pipe=round(fl / ms / 1e9 / fp32_peak, 3))))
    print(json.dumps(dict(peaks=dict(hbm_GBps=hbm, fp32_fma_TFLOPs=round(fp32_peak, 1)))))


if __name__ == "__main__":
    main()
