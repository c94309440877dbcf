"""Condense an `ncu --page raw --csv` export into one row per distinct kernel (the longest launch of each)."""
import csv
import sys

COLS = [("gpu__time_duration.sum", "ms"), ("dram__bytes_read.sum", "rdGB"), ("dram__bytes_write.sum", "wrGB"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "lsu%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shWave"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shConfl"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math")]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    best = {}
    for r in data:
        name = r[idx["Kernel Name"]].replace("<unnamed>::", "").replace("void ", "").split("(")[0]
        t = float(r[idx["gpu__time_duration.sum"]])
        if units[idx["gpu__time_duration.sum"]] == "us":
            t /= 1e3
        if name not in best or t > best[name][0]:
            best[name] = (t, r)
    print("| kernel | " + " | ".join(c for _, c in COLS) + " |")
    print("|---|" + "---:|" * len(COLS))
    for name, (t, r) in sorted(best.items(), key=lambda kv: -kv[1][0]):
        cells = []
        for m, c in COLS:
            if m not in idx:
                cells.append("-")
                continue
            v = r[idx[m]]
            try:
                f = float(v)
                u = units[idx[m]]
                if c == "ms":
                    f = t
                if c in ("rdGB", "wrGB"):
                    f = f / 1e3 if u == "Mbyte" else (f / 1e6 if u == "Kbyte" else (f / 1e9 if u == "byte" else f))
                cells.append(f"{f:.3g}" if abs(f) < 1e6 else f"{f / 1e6:.1f}M")
            except ValueError:
                cells.append(v)
        print(f"| `{name}` | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
