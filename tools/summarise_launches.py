"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list into a markdown share table.

usage: python tools/summarise_launches.py gpurun_out/launches.csv [--last N]   (N = launches of the last frame / call to keep)
"""
import argparse
import collections
import csv
import re


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--last", type=int, default=0)
    ap.add_argument("--title", default="")
    ap.add_argument("--only", default="", help="keep kernels whose name contains this (e.g. ua2::)")
    args = ap.parse_args()
    lines = [l for l in open(args.csv) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    rows = [r for r in rows if r.get("Metric Name") == "gpu__time_duration.sum"]
    if args.only:
        rows = [r for r in rows if args.only in r["Kernel Name"]]
    if args.last:
        rows = rows[-args.last:]
    agg = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"^void |\(.*$", "", r["Kernel Name"]).replace("(anonymous namespace)::", "").replace("unnamed>::", "").replace("ua2::<", "")
        if name.startswith("cutlass::device_kernel"):  # keep the MMA atom of the CUTLASS collective, drop the rest of the type
            m = re.search(r"SM100_MMA_\w+<[^>]*>", r["Kernel Name"])
            name = "cutlass::device_kernel<GemmUniversal<TMA + UMMA warp-specialised, " + (m.group(0) if m else "?") + ">>"
        key = (name, r["Grid Size"], r["Block Size"])
        v = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("ns", "nsecond"):
            v /= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            v *= 1e3
        a = agg.setdefault(key, [0.0, 0])
        a[0] += v
        a[1] += 1
    total = sum(a[0] for a in agg.values())
    n = sum(a[1] for a in agg.values())
    if args.title:
        print(f"## {args.title}\n")
    print(f"total {total:.1f} us over {n} launches (ncu-serialised, cold cache)\n")
    print("| share | total us | launches | avg us | kernel | grid | block |\n|---|---|---|---|---|---|---|")
    for (name, grid, block), (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"| {100 * t / total:.1f}% | {t:.1f} | {c} | {t / c:.2f} | `{name}` | {grid} | {block} |")


if __name__ == "__main__":
    main()
