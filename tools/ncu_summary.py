#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list (B200_PROFILING.md recipe) per kernel:
launches, total / mean device time and share of the captured step.  Usage:
    python tools/ncu_summary.py gpurun_out/launches.csv > profiles/rNN_launches.md
ncu serialises launches and runs them cold-cache, so only the SHARES are comparable with the CUDA-event numbers of
bench.py, not the absolutes."""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*$", "", r["Kernel Name"]).replace("void ", "").strip()
        rows.append((name, r["Grid Size"], r["Block Size"], float(r["Metric Value"].replace(",", "")) / 1e3))
    agg = OrderedDict()
    for name, grid, block, us in rows:
        a = agg.setdefault(name, [0, 0.0, block])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    print(f"# ncu launch list: {len(rows)} launches, {total / 1e3:.3f} ms summed device time\n")
    print("| kernel | launches | total ms | mean us | share | block |")
    print("|---|---:|---:|---:|---:|---|")
    for name, (n, us, block) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {n} | {us / 1e3:.3f} | {us / n:.1f} | {100 * us / total:.1f}% | {block} |")


if __name__ == "__main__":
    main(sys.argv[1])
