#!/usr/bin/env python
"""Summarise `ncu --set full` reports (.ncu-rep) as a markdown table, one column per captured launch, and write the
DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) to a JSON that bench.py reads for
`roofline.traffic`.  Usage (here, after gpurun brought the reports back):

    python tools/ncu_full_summary.py profiles/r01e_ncu_full.md profiles/gemm_traffic.json gpurun_out/gemm_prof.ncu-rep ...
"""
import csv
import io
import json
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "smsp__inst_executed.sum",
]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main(md_path, json_path, reports):
    md = ["# ncu --set full --clock-control none captures; one column per captured launch\n"]
    traffic = {}
    for rep in reports:
        hdr, units, data = raw(rep)
        md.append(f"## {rep.split('/')[-1]}")
        md.append("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |")
        md.append("|---|---|" + "---|" * len(data))
        ik = hdr.index("Kernel Name")
        md.append("| Kernel Name |  | " + " | ".join(r[ik][:70] for r in data) + " |")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                md.append(f"| {m} | {units[i]} | " + " | ".join(r[i] for r in data) + " |")
        md.append("")
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        it = hdr.index("gpu__time_duration.sum")
        for n, r in enumerate(data):
            traffic[f"{rep.split('/')[-1]}:{n}"] = {
                "kernel": r[ik][:90], "dram_bytes": to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]),
                "duration_us": float(r[it].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(units[it], 1)}
    open(md_path, "w").write("\n".join(md) + "\n")
    json.dump(traffic, open(json_path, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3:])
