#!/usr/bin/env python
"""bench.py -- learner-side PPO throughput (GAE + normalisation + full update) of rlgym_ppo_b200 on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c3]

One "step" = one learner iteration at steady state on synthetic rollouts of the example shape:
Learner.add_new_experience on N_new fresh timesteps (value inference on N_new+1 states, GAE + reward
normalisation, Welford update, ring append) followed by PPOLearner.learn on the full buffer
(ppo_epochs x floor(buffer/B) optimiser steps: permutation gather, fwd/bwd, clip, Adam).

  value   device-timed (CUDA events per step, L2 flushed between steps), rollout arrays already resident in HBM
  e2e     the same iteration through the public API with HOST (pinned) rollout arrays: H2D of the 7 arrays and the
          D2H read of the report scalars inside the timed region (wall clock, synchronised both sides)
  roofline    per-kernel CUDA-event timing of one more step (rlgym_ppo_b200._lib.timing_begin), dominant kernel
  cpu_baseline  the CPU oracle (oracle/ref_oracle.py: the reference's algorithm restated, fp32 torch-CPU + the
          reference's Python GAE loop) on this box's host cores, bounded sample (rank 0, N=1 only)

`--impl reference` times that CPU path alone with all host threads and prints the same JSON line.
Multi-GPU (torchrun, one rank per GPU): weak scaling -- every rank contributes N_new timesteps and takes 1/R of
every batch; the global batch, rollout and buffer grow with R (config.global_*).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: example.py shape on 1xB200
    "c2": dict(name="example.py shape (configs[1]): 50k-step rollout, obs 89, 90 actions, 256x3 MLPs, buffer 150k, "
                    "batch 50k, 1 epoch", n_new=50000, obs=89, act=90, layers=(256, 256, 256), buffer=150000,
               batch=50000, epochs=1, ent=0.001),
    # BASELINE.json configs[2] (per-GPU share of it)
    "c3": dict(name="large nets (configs[2]): 2048-2048-1024-1024 policy/value, buffer 150k, batch 50k, 3 epochs",
               n_new=50000, obs=89, act=90, layers=(2048, 2048, 1024, 1024), buffer=150000, batch=50000, epochs=3,
               ent=0.001),
}
FLOP_PER_SAMPLE_UPDATE = {"c2": 1894912.0, "c3": 90097664.0}     # SURVEY.md 8(d), fwd+bwd both nets, un-padded
FLOP_PER_STATE_VALUE = {"c2": 308224.0, "c3": 15046656.0}


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


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm on the host cores (oracle = checker; here it is what is being timed)
# ------------------------------------------------------------------------------------------------------------------
def run_cpu_oracle(wl, steps, warmup, seed=0):
    import torch
    from oracle import ref_oracle as O
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1 to its ranks: undo that here)
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except (AttributeError, RuntimeError):
        pass
    torch.manual_seed(123)
    import torch.nn as nn

    def mk(out):
        dims = [wl["obs"], *wl["layers"], out]
        ps = []
        for i in range(len(dims) - 1):
            l = nn.Linear(dims[i], dims[i + 1])
            ps += [l.weight.detach().clone(), l.bias.detach().clone()]
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

    while buf.f["rewards"].shape[0] + n < wl["buffer"]:        # reach steady state without timing
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


def reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, min(args.warmup, 1))
    # bounded: the whole run must end within minutes (one c2 iteration is ~2 s on 8 cores, c3 far more)
    if args.workload == "c3":
        steps, warmup = 1, 0
    else:
        steps = min(steps, 5)
    tput, sec, threads = run_cpu_oracle(wl, steps, warmup)
    sample = f"{steps} steady-state iteration(s) of the full workload after {warmup} warm-up, oracle port of the reference"
    line = {"impl": "reference", "metric": "learner_timesteps_per_sec", "value": tput, "unit": "timesteps/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"]},
            "cpu_baseline": {"value": tput, "unit": "timesteps/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": tput, "unit": "timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))



# C-ABI entry point -> kernel name in the ncu reports (profiles/ncu_traffic.json is keyed by kernel name)
KERNEL_OF = {"rlppo_wgrad_multi": "wgrad_multi_kernel", "rlppo_policy_train_fused": "fused_mlp_kernel<1, 1>",
             "rlppo_value_train_fused": "fused_mlp_kernel<0, 1>", "rlppo_value_infer_fused": "fused_mlp_kernel<0, 0>",
             "rlppo_gather_batch": "gather_kernel", "rlppo_gae_f32": "gae_scan3_kernel<1, 1>",
             "rlppo_linear_wgrad": "wgrad_kernel<256>", "rlppo_linear_fwd": "rowgemm_kernel<256, 0>",
             "rlppo_linear_dgrad": "rowgemm_kernel<256, 1>", "rlppo_linear_dgrad_db": "rowgemm_kernel<256, 1>",
             "rlppo_clip_adam": "clip_adam_kernel", "rlppo_norm_clip_adam": "norm_clip_adam_kernel"}


def bound_of(v, hbm_peak, tc_peak):
    """Which roof bounds a call: the larger of flop / tensor peak and algorithmic bytes / HBM peak."""
    t_tc = v["flop"] / (tc_peak * 1e12) if v["flop"] else 0.0
    t_hbm = v["byte"] / (hbm_peak * 1e9) if v["byte"] else 0.0
    return "tensor" if t_tc > t_hbm else "hbm"


def roofline_of(name, v, hbm_peak, tc_peak, peak_src, workload):
    """roofline object for one C-ABI entry point: achieved = algorithmic work per launch / average CUDA-event time of
    a launch, against the roof that bounds it; traffic = dram bytes per launch from the committed ncu --set full
    capture (profiles/ncu_traffic.json), or null."""
    traffic = None
    try:
        tbl = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        traffic = tbl.get(workload, {}).get(KERNEL_OF.get(name, name))
    except Exception:
        pass
    bound = bound_of(v, hbm_peak, tc_peak)
    if bound == "tensor":
        ach, peak, unit, src = v["flop"] / v["ms"] / 1e9, tc_peak, "TFLOP/s", peak_src + " (sustained bf16)"
    else:
        ach, peak, unit, src = v["byte"] / v["ms"] / 1e6, hbm_peak, "GB/s", peak_src + " (copy bandwidth)"
    out = {"kernel": name, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
           "traffic": traffic, "peak_source": src, "launches": v["calls"], "avg_launch_ms": v["ms"] / v["calls"],
           "algorithmic_bytes_per_launch": v["byte"] / v["calls"], "algorithmic_flop_per_launch": v["flop"] / v["calls"]}
    if v["flop"] and v["byte"]:
        out["other_roof"] = ({"bound": "hbm", "achieved": v["byte"] / v["ms"] / 1e6, "unit": "GB/s",
                              "frac": v["byte"] / v["ms"] / 1e6 / hbm_peak} if bound == "tensor" else
                             {"bound": "tensor", "achieved": v["flop"] / v["ms"] / 1e9, "unit": "TFLOP/s",
                              "frac": v["flop"] / v["ms"] / 1e9 / tc_peak})
    return out


def gae_bandwidth(dev, hbm_peak, peak_src, log2n=26, reps=5):
    """The metric's second half ("GAE GB/s"): one flat 2^26-step rollout with random done masks (inputs 2 GB > L2),
    28 algorithmic bytes per step, CUDA events around the C-ABI call, median of `reps`."""
    import torch
    from rlgym_ppo_b200 import ops
    n = 1 << log2n
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    rew = torch.randn(n, device=dev, generator=g) * 0.1
    done = (torch.rand(n, device=dev, generator=g) < 1 / 300).float()
    tr = ((torch.rand(n, device=dev, generator=g) < 1 / 1500).float() * (1 - done))
    tr[-1] = 1 - done[-1]
    tr = tr.double()
    val = torch.randn(n + 1, device=dev, generator=g)
    std = torch.tensor([0.7], device=dev)
    out = tuple(torch.empty(n, device=dev) for _ in range(3))
    for _ in range(3):
        ops.gae(rew, done, tr, val, 0.99, 0.95, std, out=out)
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gae(rew, done, tr, val, 0.99, 0.95, std, out=out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    return {"timesteps": n, "truncated_dtype": "f64", "ms": ms, "algorithmic_GBps": 28 * n / ms / 1e6,
            "frac_of_hbm_peak": 28 * n / ms / 1e6 / hbm_peak, "actual_bytes_per_step": 32, "peak_source": peak_src,
            "timesteps_per_sec": n / ms * 1e3}

# ------------------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------------------
def b200_arm(args, wl):
    import torch
    from types import SimpleNamespace

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(dev))

    from rlgym_ppo_b200 import _lib
    from rlgym_ppo_b200.learner import Learner
    from rlgym_ppo_b200.ppo import ExperienceBuffer, PPOLearner
    from rlgym_ppo_b200.util import WelfordRunningStat
    _lib.require_device()

    R = world
    n_local = wl["n_new"]
    # Weak scaling, "sharded" data parallelism: every rank owns its rollout (as if fed by its own env workers), its
    # experience buffer and its shuffle; per-rank batch = the workload's batch, one optimiser step averages over R of them.
    n_glob, batch, cap = n_local * R, wl["batch"] * R, wl["buffer"] * R
    torch.manual_seed(123)        # identical initial weights on every rank
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        ppo = PPOLearner(wl["obs"], wl["act"], 0, wl["layers"], wl["layers"], (0.1, 1.0), wl["batch"], wl["epochs"], 3e-4,
                         3e-4, 0.2, wl["ent"], wl["batch"], dev, dp_mode="sharded")
    ns = SimpleNamespace(ppo_learner=ppo, return_stats=WelfordRunningStat(1, device=dev), standardize_returns=True,
                         gae_gamma=0.99, gae_lambda=0.95, max_returns_per_stats_increment=150,
                         experience_buffer=ExperienceBuffer(wl["buffer"], 123 + rank, dev))

    # ---- synthetic rollouts: a pool of distinct ones per rank, pinned on the host and resident on the device -------------
    rng = np.random.RandomState(rank)
    pool_host, pool_dev = [], []
    n_pool = 3
    for _ in range(n_pool):
        states, rewards, next_states, dones, truncated = synth_rollout(rng, n_local, wl["obs"])
        acts, logp = ppo.policy.get_action_device(torch.from_numpy(states).to(dev))
        host = [torch.from_numpy(a).pin_memory() for a in
                (states, acts.float().cpu().numpy(), logp.cpu().numpy(), rewards, next_states, dones, truncated)]
        pool_host.append((tuple(t.numpy() for t in host), host))   # numpy views of pinned memory (+ keep-alive)
        pool_dev.append(tuple(t.to(dev) for t in host))
    h2d_bytes = world * sum(t.numel() * t.element_size() for t in pool_host[0][1])     # all ranks
    d2h_bytes = world * ppo._tail_host.numel() * 4

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step(exp):
        Learner.add_new_experience(ns, exp)
        return ppo.learn(ns.experience_buffer)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # fill to steady state, then warm up
    it = 0
    while len(ns.experience_buffer) + n_local < wl["buffer"]:
        Learner.add_new_experience(ns, pool_dev[it % n_pool])
        it += 1
    n_warm = max(args.warmup, 3)
    barrier()                                     # ranks reach their first data-parallel step together
    for _ in range(max(n_warm, 2 * n_pool)):      # every pooled rollout at least twice: its CUDA graph exists before timing
        step(pool_dev[it % n_pool])
        it += 1

    # ---- (1) device-resident timing: K steps, one event pair per step, L2 flushed between steps --------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    calls0 = _lib.CALLS
    barrier()
    evs = []
    for _ in range(args.steps):
        flush_buf.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        report = step(pool_dev[it % n_pool])
        e1.record()
        evs.append((e0, e1))
        it += 1
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    calls = _lib.CALLS - calls0

    # ---- (2) end to end through the public API with host buffers ---------------------------------------------------
    for _ in range(n_warm):                       # the host-input path has its own staging slots and graph
        step(pool_host[it % n_pool][0])
        it += 1
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        report = step(pool_host[it % n_pool][0])
        _ = report["Policy Entropy"]
        it += 1
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if sampler else None

    if world > 1:
        t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        dev_ms, e2e_ms = float(t[0]), float(t[1])

    # ---- (3) per-kernel pass for the roofline object (rank 0; not part of the timed numbers) --------------------------
    roofline, kernels, gae_line = None, None, None
    if rank != 0 and world > 1:
        flush_buf.zero_()
        step(pool_dev[it % n_pool])      # the collectives of rank 0's instrumented step need their peers
        it += 1
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        tc_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        peak_src = "measured" if peaks else "fallback"
        flush_buf.zero_()
        torch.cuda.synchronize()
        _lib.timing_begin()
        step(pool_dev[it % n_pool])
        per = _lib.timing_end()
        it += 1
        total = sum(v["ms"] for v in per.values())
        kernels = {k.replace("rlppo_", ""): {"calls": v["calls"], "ms": round(v["ms"], 4),
                                             "share": round(v["ms"] / total, 4),
                                             **({"TFLOP/s": round(v["flop"] / v["ms"] / 1e9, 2)} if v["flop"] else {}),
                                             **({"GB/s": round(v["byte"] / v["ms"] / 1e6, 1)} if v["byte"] else {}),
                                             "bound": bound_of(v, hbm_peak, tc_peak)}
                   for k, v in sorted(per.items(), key=lambda kv: -kv[1]["ms"])}
        top_name, top = max(per.items(), key=lambda kv: kv[1]["ms"])
        roofline = roofline_of(top_name, top, hbm_peak, tc_peak, peak_src, args.workload)
        gae_line = gae_bandwidth(dev, hbm_peak, peak_src) if world == 1 else None

    # ---- (4) CPU baseline (rank 0, single-GPU run only) ------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        if args.workload == "c2":
            tput, sec, threads = run_cpu_oracle(wl, 3, 1)
            cpu = {"value": tput, "unit": "timesteps/s", "cores": threads, "kind": "port",
                   "sample": "3 steady-state iterations of the full workload after 1 warm-up (oracle/ref_oracle.py: "
                             "Python GAE loop + fp32 torch-CPU MLPs)", "sec_per_step": sec}
        else:
            small = dict(wl, n_new=5000, batch=5000, buffer=15000)
            tput, sec, threads = run_cpu_oracle(small, 1, 0)
            cpu = {"value": tput, "unit": "timesteps/s", "cores": threads, "kind": "port",
                   "sample": "1 iteration at 1/10 of the rows (5k new, batch 5k, buffer 15k), same nets and epochs",
                   "sec_per_step": sec}

    if rank == 0:
        K = args.steps
        n_updates = wl["epochs"] * (cap // batch)
        flop_step = n_glob * FLOP_PER_STATE_VALUE[args.workload] + n_updates * batch * FLOP_PER_SAMPLE_UPDATE[args.workload]
        line = {
            "metric": "learner_timesteps_per_sec", "value": n_glob * K / (dev_ms / 1e3), "unit": "timesteps/s",
            "n_gpus": world, "steps": K, "warmup": n_warm, "ms_per_step": dev_ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": wl["name"], "global_new_timesteps": n_glob, "global_batch": batch,
                       "global_buffer": cap, "optimizer_steps_per_step": n_updates,
                       "consumed_samples_per_step": n_updates * batch,
                       "parallelism": (f"dp{world} (per-rank experience shards; gradient exchange per optimiser step: "
                                       + ("summed inside the optimiser launch from the peers' arenas over NVLink)"
                                          if ppo.dp_collective == "p2p" else "NCCL allreduce of the flat gradient arena)"))
                       if world > 1 else "dp1",
                       "gradient_exchange": ppo.dp_collective,
                       "l2": "flushed between timed steps (256 MiB write); working set > L2",
                       "algorithmic_tflop_per_step": flop_step / 1e12,
                       "achieved_tflops": flop_step / (dev_ms / K / 1e3) / 1e12},
            "e2e": {"value": n_glob * K / (e2e_ms / 1e3), "unit": "timesteps/s", "ms_per_step": e2e_ms / K,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes},
            "gpu_launches": calls,
            "clocks": clocks,
            "roofline": roofline,
            "gae": gae_line,
            "cpu_baseline": cpu,
            "kernels": kernels,
            "report": {k: (float(v) if isinstance(v, (int, float, np.floating)) else v) for k, v in report.items()},
        }
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        reference_arm(args, wl)
    else:
        b200_arm(args, wl)


if __name__ == "__main__":
    main()
