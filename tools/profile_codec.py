"""Profiling driver (run under ncu): full-size Mimi-twin codec, batch x clip encode + decode.  Not a benchmark."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--seconds", type=float, default=5.0)
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    m = bench.make_codec(torch.device("cuda", 0))
    wav = torch.randn(a.batch, 1, int(a.seconds * 24000), device="cuda") * 0.1
    for _ in range(a.reps):
        codes = m.encode(wav)
        out = m.decode(codes)
    torch.cuda.synchronize()
    print("codes", tuple(codes.shape), "out", tuple(out.shape))


if __name__ == "__main__":
    main()
