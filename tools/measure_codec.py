"""Codec encode / decode time at BASELINE.json's shape (batch 16 x 10 s) with one process-wide option toggled: the A/B numbers for profiles/."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from uniaudio2_b200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
opts = [a for a in sys.argv[1:] if not a.startswith("-")] or ["conv_pointwise"]
for name in opts:
    for v in (0, 1):
        _lib.check(_lib.lib().ua2_set_global_option(name.encode(), v))
        r = bench.bench_codec(dev, cpu=False, roofline=False)
        print(json.dumps({"option": name, "value": v, "encode_ms": r["encode_ms"], "decode_ms": r["decode_ms"], "x_realtime": r["rtf_x_realtime"]}))
