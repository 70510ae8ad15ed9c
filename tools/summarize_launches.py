#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel calls, total time and share of the
profiled window (optionally only launches [first, last]).

    python tools/summarize_launches.py gpurun_out/launches_c2.csv [first last]
"""
import csv
import re
import sys
from collections import OrderedDict


def load(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr, rows = rows[0], rows[1:]
    ki, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
    gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
    out = []
    for r in rows:
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("<unnamed>::", "")
        name = re.sub(r"at::native::|at::", "", name)
        out.append((int(r[ii]), name[:60], float(r[vi].replace(",", "")) / 1e3, r[gi], r[bi]))
    return out


def main():
    rows = load(sys.argv[1])
    if len(sys.argv) >= 4:
        a, b = int(sys.argv[2]), int(sys.argv[3])
        rows = [r for r in rows if a <= r[0] <= b]
    agg = OrderedDict()
    for _, name, us, grid, block in rows:
        e = agg.setdefault(name, [0, 0.0, grid, block])
        e[0] += 1
        e[1] += us
    total = sum(e[1] for e in agg.values())
    print(f"| kernel | launches | total us | avg us | share | grid | block |")
    print("|---|---:|---:|---:|---:|---|---|")
    for name, (n, us, grid, block) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {n} | {us:.1f} | {us / n:.1f} | {us / total:.3f} | {grid} | {block} |")
    print(f"| **total** | {len(rows)} | {total:.1f} | | 1.000 | | |")


if __name__ == "__main__":
    main()
