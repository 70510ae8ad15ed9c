#!/usr/bin/env python
"""DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of every kernel in an `ncu --set full`
report, averaged over its launches, merged into profiles/ncu_traffic.json under a workload key -- the figure bench.py
reports as roofline.traffic.

    python tools/ncu_traffic.py gpurun_out/r01b_full_c2.ncu-rep c2
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"Tbyte": 1e12, "Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    rep, key = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    acc = {}
    for r in data:
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "")
        name = re.sub(r"\(bool\)|\(int\)", "", name)
        tot = 0.0
        if "dram__bytes_read.sum" in ix:
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(r[ix[m]].replace(",", "")) * SCALE[units[ix[m]]]
        else:
            # light section sets carry the rate only: bytes = dram__bytes.sum.per_second x duration
            rate_u = units[ix["dram__bytes.sum.per_second"]].split("/")[0]
            rate = float(r[ix["dram__bytes.sum.per_second"]].replace(",", "")) * SCALE[rate_u]
            dur_u = units[ix["gpu__time_duration.sum"]]
            dur = float(r[ix["gpu__time_duration.sum"]].replace(",", "")) * {"s": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9}[dur_u]
            tot = rate * dur
        grid = r[ix["Grid Size"]]
        a = acc.setdefault(name, [0, 0.0, 0.0])
        # keep the largest-grid launches of a name (the small warm-up launches of the same kernel are not the bench's)
        a[0] += 1
        a[1] += tot
        a[2] = max(a[2], tot)
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    tbl = json.load(open(path)) if os.path.exists(path) else {}
    tbl[key] = {k: round(v[2] if "gae_scan" in k or "rows_to_bf16" in k else v[1] / v[0]) for k, v in sorted(acc.items())}
    tbl.setdefault("_source", {})[key] = os.path.basename(rep)
    json.dump(tbl, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(tbl[key], indent=1))


if __name__ == "__main__":
    main()
