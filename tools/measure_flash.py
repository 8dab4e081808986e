"""Time the tensor-core attention (ua2_flash_attn_bf16) on the flow decoder's shape and report TFLOP/s; also check it against fp32 SDPA."""
import json
import sys

import torch

sys.path.insert(0, ".")
from uniaudio2_b200 import _lib  # noqa: E402

L, P = _lib.lib(), _lib.ptr
for (B, H, T) in [(2, 24, 500), (8, 24, 500), (2, 24, 1500), (16, 24, 1000)]:
    q = torch.randn(B, H, T, 64, device="cuda").bfloat16()
    k = torch.randn(B, H, T, 64, device="cuda").bfloat16()
    v = torch.randn(B, H, T, 64, device="cuda").bfloat16()
    out = torch.empty(B, T, H * 64, device="cuda")
    ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float()).permute(0, 2, 1, 3).reshape(B, T, -1)
    for sbuf in (1, 2):
        _lib.check(L.ua2_set_global_option(b"flash_sbuf", sbuf))
        for _ in range(5):
            _lib.check(L.ua2_flash_attn_bf16(P(q), P(k), P(v), P(out), B, T, H, 64, None))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 50
        e0.record()
        for _ in range(n):
            _lib.check(L.ua2_flash_attn_bf16(P(q), P(k), P(v), P(out), B, T, H, 64, None))
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / n * 1e3
        err = float((out - ref).abs().max())
        fl = 4.0 * B * H * T * T * 64
        print(json.dumps({"kernel": "flash_bf16_kernel<%d>" % sbuf, "B": B, "H": H, "T": T, "us": round(us, 2), "TFLOPs": round(fl / us / 1e6, 1),
                          "max_abs_err_vs_fp32_sdpa": err}))
_lib.check(L.ua2_set_global_option(b"flash_sbuf", 0))
