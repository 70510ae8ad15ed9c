#!/usr/bin/env python
"""Per-kernel SASS opcode summary of librlppo_b200.so (no GPU needed): which kernels carry the Blackwell-specific
instructions -- UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA tensor load / store),
UBLKCP (cp.async.bulk), SYNCS (mbarrier), UTCBAR (tcgen05.commit), griddepcontrol (ACQBULK / PDL) -- and how many
instructions each kernel has in total.

    python tools/sass_summary.py > profiles/r02_sass_opcodes.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "rlgym_ppo_b200", "librlppo_b200.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "UTMAPF", "REDG", "RED", "ATOMG",
        "MUFU", "F2FP", "HMMA", "FFMA", "DFMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    filt = subprocess.run(["cu++filt", "-p"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True)
    names = filt.stdout.split("\n") if filt.returncode == 0 else []
    kernels, cur = [], None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = [m.group(1), collections.Counter(), 0]
            kernels.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur[1][m.group(1)] += 1
            cur[2] += 1
    print("# SASS opcode summary of `rlgym_ppo_b200/librlppo_b200.so` (sm_100a; `cuobjdump -sass`, tools/sass_summary.py)\n")
    print("| kernel | instrs | " + " | ".join(KEYS) + " |")
    print("|---|---:|" + "---:|" * len(KEYS))
    for i, (mangled, cnt, total) in enumerate(kernels):
        name = names[i] if i < len(names) and names[i] else mangled
        name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
        name = re.sub(r"\((bool|int)\)", "", name)[:70]
        print(f"| `{name}` | {total} | " + " | ".join(str(cnt.get(k, 0)) for k in KEYS) + " |")


if __name__ == "__main__":
    sys.exit(main())
