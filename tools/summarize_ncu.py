#!/usr/bin/env python3
"""Turns ncu CSV exports into the markdown summaries kept under profiles/.

    python tools/summarize_ncu.py launches <log.csv> "<title>"          # --metrics gpu__time_duration.sum --csv --log-file
    python tools/summarize_ncu.py raw <raw.csv> "<title>" [max_rows]    # ncu -i x.ncu-rep --page raw --csv
"""
import csv
import sys

RAW_COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
            "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed"]


def short(name):
    name = name.replace("ccb::", "")
    return name.split("(")[0].strip()


def launches(path, title):
    rows = list(csv.reader(open(path, newline="")))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h = rows[hi]
    ix = {n: i for i, n in enumerate(h)}
    agg, total, n = {}, 0.0, 0
    unit = None
    for r in rows[hi + 1:]:
        if len(r) < len(h):
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        if unit in ("ns", "nsecond"):
            v /= 1e3
        elif unit in ("ms", "msecond"):
            v *= 1e3
        a = agg.setdefault(short(r[ix["Kernel Name"]]), [0, 0.0, 0.0])
        a[0] += 1
        a[1] += v
        a[2] = max(a[2], v)
        total += v
        n += 1
    print(f"# {title}\n\ntotal {total / 1e3:.1f} ms over {n} launches (cold-cache, serialised under ncu: compare SHARES, not absolutes)\n")
    print("| kernel | launches | total us | share | avg us | max us |\n|---|---|---|---|---|---|")
    for k, (c, t, m) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {c} | {t:.1f} | {t / total:.3f} | {t / c:.1f} | {m:.1f} |")


def raw(path, title, max_rows=60):
    rows = list(csv.reader(open(path, newline="")))
    h, units = rows[0], rows[1]
    ix = {n: i for i, n in enumerate(h)}
    cols = [c for c in RAW_COLS if c in ix]
    print(f"# {title}\n")
    print("| kernel | " + " | ".join(f"{c} [{units[ix[c]]}]" for c in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for r in rows[2:2 + max_rows]:
        if len(r) < len(h):
            continue
        print(f"| {short(r[ix['Kernel Name']])} | " + " | ".join(r[ix[c]] for c in cols) + " |")


if __name__ == "__main__":
    kind, path, title = sys.argv[1:4]
    if kind == "launches":
        launches(path, title)
    else:
        raw(path, title, int(sys.argv[4]) if len(sys.argv) > 4 else 60)
