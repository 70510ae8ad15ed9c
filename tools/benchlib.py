"""Pieces of bench.py that are not the timed region itself: synthetic rollouts, the clock sampler, the reference arm
(the UNMODIFIED rlgym-ppo installed into baseline/_ref, else the oracle port), and the section-8(d) algorithmic work
table the roofline fractions are computed from."""
import json
import os
import subprocess
import sys
import threading
import time
import types
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def synth_rollout(rng, n, obs_dim):
    """the flat layout collect_timesteps produces (batched_agent_manager.py:159-168)."""
    states = rng.randn(n, obs_dim).astype(np.float32)
    next_states = np.roll(states, -1, axis=0).copy()
    rewards = (rng.randn(n) * 0.1).astype(np.float32)
    dones = (rng.rand(n) < 1 / 300).astype(np.float32)
    truncated = ((rng.rand(n) < 1 / 1500) * (1 - dones)).astype(np.float64)
    truncated[-1] = 1.0 - dones[-1]
    return states, rewards, next_states, dones, truncated


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], 0.0, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                continue
        # "under load" = the upper half of the samples (the sampler also sees the idle gaps between steps)
        sm.sort()
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_peaks():
    """MEASURED_PEAKS.json (driver-written) or the profiling recipe's fallback."""
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return {"hbm": float(p["hbm_gbs"]), "tc": float(p["bf16_tflops_sustained"]), "tc_burst": float(p["bf16_tflops"]),
                "src": "MEASURED_PEAKS.json"}
    except Exception:
        return {"hbm": 6650.0, "tc": 1400.0, "tc_burst": 1700.0, "src": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------------------------
# The reference itself (AechPro/rlgym-ppo v1.3.13, unmodified, pip-installed into baseline/_ref by build())
# ------------------------------------------------------------------------------------------------------------------
def import_reference():
    """The installed reference package, or None.  It needs `gym` at import time (rlgym_ppo/util/__init__.py:3 ->
    rlgym_v2_gym_wrapper.py:1) and gym is not in this image: a stub module stands in (no reference code is modified;
    nothing on the learner path touches gym)."""
    if not os.path.isdir(os.path.join(REF_DIR, "rlgym_ppo")):
        return None
    if "gym" not in sys.modules:
        try:
            import gym  # noqa: F401
        except ImportError:
            gym = types.ModuleType("gym")
            gym.Env = object
            gym.spaces = types.ModuleType("gym.spaces")
            sys.modules["gym"] = gym
            sys.modules["gym.spaces"] = gym.spaces
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import rlgym_ppo  # noqa: F401
    from rlgym_ppo.learner import Learner
    from rlgym_ppo.ppo import ExperienceBuffer, PPOLearner
    from rlgym_ppo.util import WelfordRunningStat
    assert os.path.realpath(rlgym_ppo.__file__).startswith(os.path.realpath(REF_DIR)), rlgym_ppo.__file__
    return SimpleNamespace(Learner=Learner, ExperienceBuffer=ExperienceBuffer, PPOLearner=PPOLearner,
                           WelfordRunningStat=WelfordRunningStat)


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(wl, steps, warmup, device="cpu", seed=0, fill=True):
    """One learner iteration = the reference's own Learner.add_new_experience (unbound, on a namespace carrying the
    attributes it reads: learner.py:330-385) + PPOLearner.learn (ppo_learner.py:92-238), stock code path, its own
    ExperienceBuffer / WelfordRunningStat / compute_gae, on `device` ("cpu": all host threads; "cuda:0": PyTorch eager
    on the B200 -- BASELINE.md section 3's same-box GPU comparator; its GAE is still the Python loop).
    Returns (timesteps/s, s/iteration, threads, kind)."""
    import contextlib
    import io

    import torch
    ref = import_reference()
    if ref is None:
        return run_port(wl, steps, warmup, seed) + ("port",)
    torch.set_num_threads(host_threads())           # torchrun exports OMP_NUM_THREADS=1 to its ranks: undo that here
    torch.manual_seed(123)
    with contextlib.redirect_stdout(io.StringIO()):
        ppo = ref.PPOLearner(wl["obs"], wl["act"], 0, wl["layers"], wl["layers"], (0.1, 1.0), wl["batch"], wl["epochs"],
                             3e-4, 3e-4, 0.2, wl["ent"], wl["batch"], device)
    ns = SimpleNamespace(ppo_learner=ppo, return_stats=ref.WelfordRunningStat(1), standardize_returns=True,
                         gae_gamma=0.99, gae_lambda=0.95, max_returns_per_stats_increment=150,
                         experience_buffer=ref.ExperienceBuffer(wl["buffer"], 123, "cpu"))     # learner.py:124-126
    rng = np.random.RandomState(seed)
    n = wl["n_new"]

    def rollout():
        states, rewards, next_states, dones, truncated = synth_rollout(rng, n, wl["obs"])
        with torch.no_grad():
            acts, logp = ppo.policy.get_action(states)          # discrete_policy.py:44-62, returns CPU tensors
        return (states, acts.numpy().astype(np.float32), logp.numpy().astype(np.float32), rewards, next_states, dones,
                truncated)

    def sync():
        if "cuda" in str(device):
            torch.cuda.synchronize()

    if fill:
        while ns.experience_buffer.rewards.shape[0] + n < wl["buffer"]:        # reach steady state without timing
            ref.Learner.add_new_experience(ns, rollout())
    times = []
    for it in range(warmup + steps):
        exp = rollout()
        sync()
        t0 = time.perf_counter()
        ref.Learner.add_new_experience(ns, exp)
        ppo.learn(ns.experience_buffer)
        sync()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    sec = float(np.mean(times))
    return n / sec, sec, torch.get_num_threads(), "reference"


def run_port(wl, steps, warmup, seed=0):
    """Fallback when baseline/_ref is absent: the oracle port of the same algorithm (oracle/ref_oracle.py)."""
    import torch
    import torch.nn as nn

    from oracle import ref_oracle as O
    torch.set_num_threads(host_threads())
    torch.manual_seed(123)

    def mk(out):
        dims = [wl["obs"], *wl["layers"], out]
        ps = []
        for i in range(len(dims) - 1):
            lin = nn.Linear(dims[i], dims[i + 1])
            ps += [lin.weight.detach().clone(), lin.bias.detach().clone()]
        return ps

    pol, val = mk(wl["act"]), mk(1)
    orc = O.PPOLearnerOracle(pol, val, wl["batch"], wl["epochs"], 3e-4, 3e-4, 0.2, wl["ent"], wl["batch"])
    buf = O.BufferOracle(wl["buffer"], 123)
    stats = O.WelfordOracle(1)
    rng = np.random.RandomState(seed)
    n = wl["n_new"]

    def rollout():
        states, rewards, next_states, dones, truncated = synth_rollout(rng, n, wl["obs"])
        with torch.no_grad():
            p = torch.clamp(O.policy_probs(orc.pol, torch.from_numpy(states)), 1e-11, 1.0)
            a = torch.multinomial(p, 1, True)
            lp = torch.log(p).gather(-1, a).flatten().numpy()
        return states, a.flatten().numpy().astype(np.float32), lp, rewards, next_states, dones, truncated

    while buf.f["rewards"].shape[0] + n < wl["buffer"]:
        O.add_new_experience(orc.val, buf, stats, rollout(), 0.99, 0.95)
    times = []
    for it in range(warmup + steps):
        exp = rollout()
        t0 = time.perf_counter()
        O.add_new_experience(orc.val, buf, stats, exp, 0.99, 0.95)
        orc.learn(buf)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return n / float(np.mean(times)), float(np.mean(times)), torch.get_num_threads()
