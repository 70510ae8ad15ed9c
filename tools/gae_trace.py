#!/usr/bin/env python
"""Event trace of the persistent GAE scan (RLPPO_GAE_TRACE): per tile of three CTAs (first, middle, last), the
globaltimer stamps of  0 inputs landed | 1 aggregate handed to the look-back warp | 2 carry received | 3 outputs
stored | 4 look-back: aggregate published | 5 look-back: all records arrived | 6 look-back: carry handed over.

    python tools/gae_trace.py [--log2 24] > gpurun_out/gae_trace.txt
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PATH = "/tmp/rlppo_gae_trace.txt"
os.environ["RLPPO_GAE_TRACE"] = PATH

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2", type=int, default=24)
    args = ap.parse_args()
    from rlgym_ppo_b200 import _lib, ops
    _lib.require_device()
    dev = "cuda:0"
    n = 1 << args.log2
    g = torch.Generator(device=dev).manual_seed(1)
    rew = torch.randn(n, device=dev, generator=g) * 0.1
    done = (torch.rand(n, device=dev, generator=g) < 1 / 300).float()
    tr = torch.zeros(n, device=dev, dtype=torch.float64)
    val = torch.randn(n + 1, device=dev, generator=g)
    std = torch.tensor([0.7], device=dev)
    out = tuple(torch.empty(n, device=dev) for _ in range(3))
    for _ in range(3):
        ops.gae(rew, done, tr, val, 0.99, 0.95, std, out=out)
    torch.cuda.synchronize()
    rows = [list(map(int, l.split())) for l in open(PATH)]
    t0 = min(v for r in rows for v in r[2:9] if v > 0)
    names = ["in", "agg", "carry", "out", "lb_pub", "lb_recs", "lb_carry"]
    print("# us since the first stamp;  cta_slot k " + " ".join(names))
    for r in rows:
        if r[2] == 0:
            continue
        print(r[0], r[1], " ".join(f"{(v - t0) / 1e3:8.2f}" if v else "       -" for v in r[2:9]))


if __name__ == "__main__":
    main()
