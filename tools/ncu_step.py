#!/usr/bin/env python
"""One steady-state learner iteration bracketed by cudaProfilerStart/Stop, for `ncu --profile-from-start off`.

    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_c2 \
        python tools/ncu_step.py [--workload c2|c3] [--gae-log2 26]

Profiled region: one `Learner.add_new_experience` + `PPOLearner.learn` (device-resident rollout, buffer full), and, with
--gae-log2 L, one `rlppo_gae_f32` launch over 2^L flat timesteps (the C5 sweep shape) so the scan is captured at a size
where it is bandwidth- rather than latency-bound.  Nothing here is a bench number (ncu replays every launch).
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
    ap.add_argument("--gae-log2", type=int, default=0)
    ap.add_argument("--no-step", action="store_true")
    args = ap.parse_args()
    import bench
    from rlgym_ppo_b200 import _lib, ops
    from rlgym_ppo_b200.learner import Learner
    from rlgym_ppo_b200.ppo import ExperienceBuffer, PPOLearner
    from rlgym_ppo_b200.util import WelfordRunningStat
    _lib.require_device()
    wl = bench.WORKLOADS[args.workload]
    dev = "cuda:0"
    torch.manual_seed(123)
    rt = torch.cuda.cudart()

    gae_args = None
    if args.gae_log2:
        n = 1 << args.gae_log2
        g = torch.Generator(device=dev)
        g.manual_seed(5)
        rew = torch.randn(n, device=dev, generator=g) * 0.1
        done = (torch.rand(n, device=dev, generator=g) < 1 / 300).float()
        tr = ((torch.rand(n, device=dev, generator=g) < 1 / 1500).float() * (1 - done))
        tr[-1] = 1 - done[-1]
        val = torch.randn(n + 1, device=dev, generator=g)
        out = tuple(torch.empty(n, device=dev) for _ in range(3))
        std = torch.tensor([0.7], device=dev)
        gae_args = (rew, done, tr.double(), val, 0.99, 0.95, std)
        for _ in range(2):
            ops.gae(*gae_args, out=out)

    if not args.no_step:
        with contextlib.redirect_stdout(io.StringIO()):
            ppo = PPOLearner(wl["obs"], wl["act"], 0, wl["layers"], wl["layers"], (0.1, 1.0), wl["batch"], wl["epochs"],
                             3e-4, 3e-4, 0.2, wl["ent"], wl["batch"], dev)
        ns = SimpleNamespace(ppo_learner=ppo, return_stats=WelfordRunningStat(1, device=dev), standardize_returns=True,
                             gae_gamma=0.99, gae_lambda=0.95, max_returns_per_stats_increment=150,
                             experience_buffer=ExperienceBuffer(wl["buffer"], 123, dev))
        rng = np.random.RandomState(0)
        states, rewards, next_states, dones, truncated = bench.synth_rollout(rng, wl["n_new"], wl["obs"])
        acts, logp = ppo.policy.get_action_device(torch.from_numpy(states).to(dev))
        exp = tuple(torch.from_numpy(a).to(dev) for a in
                    (states, acts.float().cpu().numpy(), logp.cpu().numpy(), rewards, next_states, dones, truncated))

        def step():
            Learner.add_new_experience(ns, exp)
            return ppo.learn(ns.experience_buffer)

        for _ in range(6):
            step()
    torch.cuda.synchronize()
    rt.cudaProfilerStart()
    if not args.no_step:
        step()
    if gae_args is not None:
        ops.gae(*gae_args, out=out)
    torch.cuda.synchronize()
    rt.cudaProfilerStop()


if __name__ == "__main__":
    main()
