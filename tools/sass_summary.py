#!/usr/bin/env python
"""Per-kernel SASS evidence of the Blackwell-native path: counts of the tcgen05 / TMEM / TMA mnemonics (and of the
legacy HMMA path) in every kernel of the in-tree libvda.so, from `cuobjdump -sass`.

    python tools/sass_summary.py > profiles/sass_summary.txt

UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / .st, UTMALDG / UTMASTG = TMA tensor load / store, UTCBAR =
tcgen05.commit, SYNCS = mbarrier ops, HMMA = mma.sync (only the 32x32 temporal attention, which is HBM-bound),
MUFU.EX2 = exponentials."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "video_depth_anything_b200", "libvda.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "MUFU.EX2", "LDGSTS", "FFMA2"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, cur, order = collections.defaultdict(collections.Counter), None, []
    for line in sass.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = m.group(1)
            order.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + ".") or (k == "UTCHMMA" and op.startswith("UTC") and "MMA" in op):
                    counts[cur][k] += 1
    dm = demangle(order)
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(order)} kernels; columns = SASS instruction counts (static)")
    print(f"{'kernel':100s} " + " ".join(f"{k:>8s}" for k in ["instr"] + KEYS))
    tot = collections.Counter()
    for fn in sorted(order, key=lambda f: -counts[f]["UTCHMMA"] * 10**6 - counts[f]["_total"]):
        c = counts[fn]
        name = re.sub(r"\(.*", "", dm.get(fn, fn)).replace("void ", "")
        print(f"{name[:100]:100s} " + " ".join(f"{c[k]:8d}" for k in ["_total"] + KEYS))
        tot.update(c)
    print(f"{'TOTAL':100s} " + " ".join(f"{tot[k]:8d}" for k in ["_total"] + KEYS))


if __name__ == "__main__":
    main()
