#!/usr/bin/env python
"""GAE / normalisation bandwidth sweep (BASELINE.json configs[4], SURVEY.md 8d "C5"): flat rollouts of 2^20 .. 2^max
timesteps with random done masks, device-timed with CUDA events, inputs larger than L2 or L2 flushed between launches.

    python tools/gae_sweep.py [--max-log2 28] [--reps 5] [--trunc f64|f32] [--check]

Prints one JSON line per (n, p_done): algorithmic GB/s = 28 B/step / kernel time, the fraction of the
measured HBM peak, and (with --check) two size-independent properties: linearity in the rewards with std=None, and
agreement of a two-chunk sharded scan (rlppo_gae_chunk_summary + carry_in) with the one-launch scan.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min-log2", type=int, default=20)
    ap.add_argument("--max-log2", type=int, default=28)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--trunc", default="f64", choices=["f64", "f32"])
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    from rlgym_ppo_b200 import _lib, ops
    _lib.require_device()
    dev = torch.device("cuda:0")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    std = torch.tensor([0.7], device=dev)
    for lg in range(args.min_log2, args.max_log2 + 1, 2):
        n = 1 << lg
        for p_done in (0.0, 1 / 300, 0.1, 1.0):
            g = torch.Generator(device=dev)
            g.manual_seed(lg * 7 + int(p_done * 1000))
            rew = torch.randn(n, device=dev, generator=g) * 0.1
            done = (torch.rand(n, device=dev, generator=g) < p_done).float()
            tr = ((torch.rand(n, device=dev, generator=g) < 1 / 1500).float() * (1 - done))
            tr[-1] = 1 - done[-1]
            tr = tr.double() if args.trunc == "f64" else tr
            val = torch.randn(n + 1, device=dev, generator=g)
            out = tuple(torch.empty(n, device=dev) for _ in range(3))
            for _ in range(3):
                ops.gae(rew, done, tr, val, 0.99, 0.95, std, out=out)
            ts = []
            for _ in range(args.reps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.gae(rew, done, tr, val, 0.99, 0.95, std, out=out)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = float(np.median(ts))
            line = {"n": n, "p_done": round(p_done, 5), "trunc": args.trunc, "ms": round(ms, 4),
                    "algorithmic_GBps": round(28 * n / ms / 1e6, 1), "frac_of_measured_hbm": round(28 * n / ms / 1e6 / hbm, 4),
                    "actual_bytes_per_step": 32 if args.trunc == "f64" else 28,
                    "timesteps_per_s": round(n / ms * 1e3, 0)}
            if args.check:
                vt, adv, ret = (t.clone() for t in out)
                # (1) linearity in the rewards without normalisation: gae(2r, 2V) == 2 gae(r, V) exactly (power of two)
                o1 = ops.gae(rew, done, tr, val, 0.99, 0.95, None)
                o2 = ops.gae(rew * 2, done, tr, val * 2, 0.99, 0.95, None)
                line["linear_exact"] = bool(all(torch.equal(a * 2, b) for a, b in zip(o1, o2)))
                # (2) two-chunk sharded scan == one launch
                cut = n // 2 + 17
                s = ops.gae_chunk_summary(rew[cut:].contiguous(), done[cut:].contiguous(), tr[cut:].contiguous(),
                                          val[cut:].contiguous(), 0.99, 0.95, std)
                carry = torch.stack([s[1], s[3]])          # right chunk's first A and R (carry into it is 0)
                left = ops.gae(rew[:cut].contiguous(), done[:cut].contiguous(), tr[:cut].contiguous(),
                               val[:cut + 1].contiguous(), 0.99, 0.95, std, carry_in=carry)
                line["sharded_max_abs_diff"] = float(max((a - b[:cut]).abs().max() for a, b in zip(left, (vt, adv, ret))))
                del o1, o2, left
            print(json.dumps(line), flush=True)
            del rew, done, tr, val, out
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
