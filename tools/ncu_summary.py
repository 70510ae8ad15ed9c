#!/usr/bin/env python
"""Per-launch summary of an `ncu --set full` report (read here, on the CPU box, with `ncu -i`).

    python tools/ncu_summary.py gpurun_out/r01b_full_c2.ncu-rep > profiles/r01b_full_c2_summary.md

One row per profiled launch: duration, DRAM bytes read/written (the `roofline.traffic` figure), DRAM / tensor-pipe /
SM throughput as % of peak, achieved occupancy, registers, shared-memory bank-conflict ratio.
"""
import csv
import io
import re
import subprocess
import sys

COLS = [
    ("us", "gpu__time_duration.sum"),
    ("dram_rd_MB", "dram__bytes_read.sum"),
    ("dram_wr_MB", "dram__bytes_write.sum"),
    ("dram_%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor_%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("sm_%", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l1tex_%", "l1tex__throughput.avg.pct_of_peak_sustained_active"),
    ("warps_%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("ipc", "sm__inst_executed.avg.per_cycle_active"),
    ("regs", "launch__registers_per_thread"),
    ("smem_ld_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum"),
    ("smem_ld_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"),
]
SCALE = {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6, "ms": 1e3, "us": 1.0, "ns": 1e-3, "ms ": 1e3}


def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    names = [c for c, m in COLS if m in ix]
    print("| # | kernel | grid | block | " + " | ".join(names) + " |")
    print("|---|---|---|---|" + "---:|" * len(names))
    for n, r in enumerate(data):
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "")
        name = re.sub(r"at::native::|at::", "", name)[:48]
        vals = []
        for c, m in COLS:
            if m not in ix:
                continue
            try:
                v = float(r[ix[m]].replace(",", "")) * SCALE.get(units[ix[m]], 1.0)
                vals.append(f"{v:.3f}" if abs(v) < 100 else f"{v:.1f}")
            except ValueError:
                vals.append(r[ix[m]])
        print(f"| {n} | `{name}` | {r[ix['Grid Size']]} | {r[ix['Block Size']]} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
