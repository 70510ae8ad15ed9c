#!/usr/bin/env python
"""Where does the HOST time of a steady-state c2 iteration go?  (cProfile over 200 device-resident steps; the GPU idles
while the host prepares the next graph launch because PPOLearner.learn synchronises to return its report.)

    python tools/host_profile.py > gpurun_out/host_profile.txt
"""
import contextlib, cProfile, io, os, pstats, sys, time
from types import SimpleNamespace
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from rlgym_ppo_b200.learner import Learner
from rlgym_ppo_b200.ppo import ExperienceBuffer, PPOLearner
from rlgym_ppo_b200.util import WelfordRunningStat
wl = bench.WORKLOADS["c2"]; dev = "cuda:0"
torch.manual_seed(123)
with contextlib.redirect_stdout(io.StringIO()):
    ppo = PPOLearner(wl["obs"], wl["act"], 0, wl["layers"], wl["layers"], (0.1, 1.0), wl["batch"], wl["epochs"], 3e-4, 3e-4, 0.2, wl["ent"], wl["batch"], dev)
ns = SimpleNamespace(ppo_learner=ppo, return_stats=WelfordRunningStat(1, device=dev), standardize_returns=True, gae_gamma=0.99, gae_lambda=0.95,
                     max_returns_per_stats_increment=150, experience_buffer=ExperienceBuffer(wl["buffer"], 123, dev))
rng = np.random.RandomState(0)
states, rewards, next_states, dones, truncated = bench.synth_rollout(rng, wl["n_new"], wl["obs"])
acts, logp = ppo.policy.get_action_device(torch.from_numpy(states).to(dev))
exp = tuple(torch.from_numpy(a).to(dev) for a in (states, acts.float().cpu().numpy(), logp.cpu().numpy(), rewards, next_states, dones, truncated))
def step():
    Learner.add_new_experience(ns, exp)
    return ppo.learn(ns.experience_buffer)
for _ in range(10): step()
torch.cuda.synchronize()
N = 200
t0 = time.perf_counter()
ta = tl = 0.0
for _ in range(N):
    a = time.perf_counter(); Learner.add_new_experience(ns, exp); b = time.perf_counter(); ppo.learn(ns.experience_buffer); c = time.perf_counter()
    ta += b - a; tl += c - b
print(f"wall per step {(time.perf_counter() - t0) / N * 1e3:.3f} ms: add_new_experience (enqueue only) {ta / N * 1e3:.3f} ms, learn (enqueue + wait) {tl / N * 1e3:.3f} ms")
pr = cProfile.Profile(); pr.enable()
for _ in range(N): step()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue()[:6000])
