"""Top source lines by warp-stall samples from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`.

    ncu -i gpurun_out/prof.ncu-rep --page source --print-source cuda,sass --csv > /tmp/src.csv
    python tools/ncu_source_top.py /tmp/src.csv [--top 25] [--file decoder_tc.cu]

The export holds one block per profiled launch; the launch with the most samples is reported.  For every CUDA
line: samples, share, executed warp instructions, shared-memory wavefronts (and the excess from bank conflicts),
global sectors, and the three largest stall reasons.
"""
import argparse
import csv
import sys

csv.field_size_limit(1 << 30)


def blocks(path):
    cur, hdr = None, None
    for row in csv.reader(open(path, newline="")):
        if not row:
            continue
        if row[0] in ("File Path", "File Name"):
            cur = {"name": "", "file": row[1], "rows": []}
            yield_later.append(cur)
            hdr = None
            continue
        if row[0] in ("Function Name", "Kernel Name"):
            if cur is not None:
                cur["name"] = row[1]
            continue
        if row[0] == "Line No":
            hdr = row
            if cur is not None:
                cur["hdr"] = hdr
            continue
        if cur is not None and hdr is not None:
            cur["rows"].append(row)


yield_later = []


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--top", type=int, default=25)
    ap.add_argument("--file", default=None)
    a = ap.parse_args()
    blocks(a.csv)
    best, best_n = None, -1
    for b in yield_later:
        if "hdr" not in b or (a.file and a.file not in (b["file"] or "")):
            continue
        h = b["hdr"]
        i_s = h.index("# Samples")
        n = 0
        for r in b["rows"]:
            if r[0] and len(r) > i_s:
                try:
                    n += int(r[i_s])
                except ValueError:
                    pass
        if n > best_n:
            best, best_n = b, n
    if best is None:
        sys.exit("no source blocks found")
    h = best["hdr"]
    col = {name: i for i, name in enumerate(h)}
    stall_cols = [(n, i) for n, i in col.items() if n.startswith("stall_") and "Not Issued" not in n]
    lines = []
    for r in best["rows"]:
        if not r[0]:
            continue
        try:
            s = int(r[col["# Samples"]])
        except ValueError:
            continue
        lines.append((s, r))
    lines.sort(key=lambda x: -x[0])
    print(f"kernel: {best['name'][:100]}\nfile: {best['file']}\nsamples: {best_n}\n")
    print("| line | samples | share | warp inst | sh.wavefronts (excess) | gl.sectors | top stalls | source |")
    print("|---:|---:|---:|---:|---:|---:|---|---|")

    def num(r, name):
        try:
            return int(r[col[name]])
        except (ValueError, KeyError, IndexError):
            return 0

    for s, r in lines[:a.top]:
        st = sorted(((num(r, n), n[6:]) for n, _ in stall_cols), reverse=True)[:3]
        st = ", ".join(f"{n} {v}" for v, n in st if v > 0)
        print(f"| {r[0]} | {s} | {100.0 * s / max(best_n, 1):.1f}% | {num(r, 'Instructions Executed')} | "
              f"{num(r, 'L1 Wavefronts Shared')} ({num(r, 'L1 Wavefronts Shared Excessive')}) | "
              f"{num(r, 'L2 Theoretical Sectors Global')} | {st} | `{r[1].strip()[:90]}` |")


if __name__ == "__main__":
    main()
