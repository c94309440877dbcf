"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table.

    python tools/summarize_launches.py gpurun_out/launches.csv [--iters N] > profiles/rNN_launches.md

Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's CUDA-event
breakdown, not absolute times.
"""
import argparse
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("<unnamed>::", "").replace("void ", "")
    return name[:90]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--iters", type=int, default=1, help="training iterations covered by the capture")
    ap.add_argument("--top", type=int, default=40)
    a = ap.parse_args()
    lines = [l for l in open(a.csv) if not l.startswith("==")]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        if row["Metric Unit"] in ("us", "usecond"):
            v *= 1e3
        elif row["Metric Unit"] in ("ms", "msecond"):
            v *= 1e6
        k = short(row["Kernel Name"])
        e = agg.setdefault(k, [0, 0.0, row["Grid Size"], row["Block Size"]])
        e[0] += 1
        e[1] += v
        n += 1
    tot = sum(e[1] for e in agg.values())
    ours = sum(e[1] for k, e in agg.items() if not (k.startswith("at::") or "cutlass" in k or "gemv" in k
                                                     or k.startswith("std::") or "cublas" in k.lower()))
    print(f"launches: {n} ({n / a.iters:.0f} per iteration over {a.iters} iterations); "
          f"sum of kernel time {tot / 1e6:.2f} ms ({tot / 1e6 / a.iters:.2f} ms per iteration); "
          f"library-side (torch glue) share {100 * (tot - ours) / tot:.1f}%\n")
    print("| launches/iter | ms/iter | share | grid | block | kernel |")
    print("|---:|---:|---:|---|---|---|")
    for k, e in sorted(agg.items(), key=lambda kv: -kv[1][1])[:a.top]:
        print(f"| {e[0] / a.iters:.1f} | {e[1] / 1e6 / a.iters:.3f} | {100 * e[1] / tot:.1f}% | {e[2]} | {e[3]} | `{k}` |")


if __name__ == "__main__":
    sys.exit(main())
