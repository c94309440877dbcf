"""Rank source lines by EXCESS shared-memory wavefronts (bank-conflict replays) per kernel.

    ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > src.csv ; python tools/ncu_conflicts.py src.csv
"""
import csv
import sys

csv.field_size_limit(1 << 30)


def main(path, top=8):
    cur, hdr, blocks = None, None, []
    for row in csv.reader(open(path, newline="")):
        if not row:
            continue
        if row[0] in ("File Path", "File Name"):
            cur = {"file": row[1], "name": "", "rows": []}
            blocks.append(cur)
            hdr = None
        elif row[0] in ("Function Name", "Kernel Name"):
            cur["name"] = row[1]
        elif row[0] == "Line No":
            hdr = row
            cur["hdr"] = hdr
        elif cur is not None and hdr is not None:
            cur["rows"].append(row)
    seen = set()
    for b in blocks:
        if "hdr" not in b:
            continue
        key = (b["name"].replace("<unnamed>::", "").replace("void ", "")[:48], b["file"].split("/")[-1])
        if key in seen:
            continue
        seen.add(key)
        col = {n: i for i, n in enumerate(b["hdr"])}

        def num(r, n):
            try:
                return int(r[col[n]])
            except (ValueError, KeyError, IndexError):
                return 0

        lines = [(num(r, "L1 Wavefronts Shared Excessive"), num(r, "L1 Wavefronts Shared"), num(r, "# Samples"), r[0],
                  r[1].strip()[:110]) for r in b["rows"] if r[0]]
        tot, ex = sum(l[1] for l in lines), sum(l[0] for l in lines)
        if tot == 0 or ex * 50 < tot:
            continue
        print(f"\n== {key[0]} [{key[1]}]: shared wavefronts {tot}, excessive {ex} ({100.0 * ex / tot:.0f}%)")
        for l in sorted(lines, reverse=True)[:top]:
            if l[0]:
                print(f"  line {l[3]:>4}: excess {l[0]:>10} of {l[1]:>10}  samples {l[2]:>6}  {l[4]}")


if __name__ == "__main__":
    main(sys.argv[1])
