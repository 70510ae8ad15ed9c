#!/usr/bin/env python
"""Device timeline of one steady-state learner iteration (no nsys in the image: torch.profiler / CUPTI activity records).

    python tools/timeline.py [--workload c2|c3] [--host] > gpurun_out/timeline.txt

Prints every kernel / memcpy / memset of one iteration with its start offset, duration and the idle gap before it, then
the totals: busy time, idle time, span.  `--host` feeds the step pinned HOST arrays (the e2e path) instead of device-
resident ones.
"""
import argparse
import contextlib
import io
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--host", action="store_true")
    ap.add_argument("--iters", type=int, default=2)
    args = ap.parse_args()
    import bench
    from rlgym_ppo_b200.learner import Learner
    from rlgym_ppo_b200.ppo import ExperienceBuffer, PPOLearner
    from rlgym_ppo_b200.util import WelfordRunningStat
    wl = bench.WORKLOADS[args.workload]
    dev = "cuda:0"
    torch.manual_seed(123)
    with contextlib.redirect_stdout(io.StringIO()):
        ppo = PPOLearner(wl["obs"], wl["act"], 0, wl["layers"], wl["layers"], (0.1, 1.0), wl["batch"], wl["epochs"], 3e-4,
                         3e-4, 0.2, wl["ent"], wl["batch"], dev)
    ns = SimpleNamespace(ppo_learner=ppo, return_stats=WelfordRunningStat(1, device=dev), standardize_returns=True,
                         gae_gamma=0.99, gae_lambda=0.95, max_returns_per_stats_increment=150,
                         experience_buffer=ExperienceBuffer(wl["buffer"], 123, dev))
    rng = np.random.RandomState(0)
    n = wl["n_new"]
    states, rewards, next_states, dones, truncated = bench.synth_rollout(rng, n, wl["obs"])
    acts, logp = ppo.policy.get_action_device(torch.from_numpy(states).to(dev))
    host = [torch.from_numpy(a).pin_memory() for a in
            (states, acts.float().cpu().numpy(), logp.cpu().numpy(), rewards, next_states, dones, truncated)]
    exp_host = tuple(t.numpy() for t in host)
    exp_dev = tuple(t.to(dev) for t in host)
    exp = exp_host if args.host else exp_dev

    def step():
        Learner.add_new_experience(ns, exp)
        return ppo.learn(ns.experience_buffer)

    for _ in range(8):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(args.iters):
            step()
            torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    if not evs:
        print("no device events recorded")
        return
    # keep the LAST iteration: events after the largest gap
    starts = [e.time_range.start for e in evs]
    gaps = [(starts[i] - evs[i - 1].time_range.end, i) for i in range(1, len(evs))]
    cut = max(gaps)[1] if args.iters > 1 else 0
    evs = evs[cut:]
    t0 = evs[0].time_range.start
    busy, prev_end = 0.0, t0
    print(f"{'start_us':>10} {'dur_us':>9} {'gap_us':>8}  name")
    for e in evs:
        s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
        gap = e.time_range.start - prev_end
        print(f"{s:10.1f} {d:9.1f} {gap:8.1f}  {e.name[:90]}")
        busy += d
        prev_end = max(prev_end, e.time_range.end)
    span = prev_end - t0
    print(f"# events {len(evs)}  span {span:.1f} us  busy {busy:.1f} us  idle {span - busy:.1f} us")


if __name__ == "__main__":
    main()
