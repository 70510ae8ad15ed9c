"""BatchedAgentManager: the learner side of timestep collection (replaces
rlgym_ppo/batched_agents/batched_agent_manager.py for the data path; process plumbing is deliberately simple).

What the reference does per environment step (batched_agent_manager.py:180-350): np.concatenate the ready
observations, one small torch forward + multinomial + .cpu(), per-message NumPy parsing, Python lists of per-agent
timesteps that are flattened with np.asarray at the end (:125-172).

Here every observation crosses PCIe exactly once and stays in HBM:
  * all workers tick in lock step; their observations land in one PINNED slab row [slots, obs];
  * one async H2D of that row, then on the device: optional standardisation (clip((x-mean)/std, +-5)) fused with
    the bf16 conversion, the policy MLP on tcgen05, softmax/clamp/categorical sample/log-prob in the head epilogue;
  * actions come back in one small D2H; observations, actions and log-probs stay in time-major device slabs;
  * a tick is ONE CUDA-graph launch (tick.TickInference) ending in the action D2H; the host waits on one event;
  * at the end of collect_timesteps the flat per-trajectory layout the learner expects (completed
    trajectories in completion order, then open ones by process id; each agent's steps in time order; last step of
    every run forced truncated unless done) is produced by ONE index gather on the device.
collect_timesteps therefore returns DEVICE tensors; Learner.add_new_experience consumes them without a copy.
"""
import multiprocessing as mp
import multiprocessing.sharedctypes
import os
import time

import numpy as np
import torch

from .. import _lib, ops
from ..util import WelfordRunningStat
from .batched_agent import batched_agent_process
from .env_worker import env_worker
from .wire import WireConn


class BatchedAgentManager(object):
    def __init__(self, policy, min_inference_size=8, seed=123, standardize_obs=True, steps_per_obs_stats_increment=5,
                 device=None):
        self.policy = policy
        self.seed = seed
        self.processes = []
        self.average_reward = None
        self.cumulative_timesteps = 0
        self.min_inference_size = min_inference_size   # kept for API parity; all slots are inferred together
        self.standardize_obs = standardize_obs
        self.steps_per_obs_stats_increment = steps_per_obs_stats_increment
        self.steps_since_obs_stats_update = 0
        self.obs_stats = None
        self.ep_rews = []
        self.n_procs = 0
        self.device = device
        self._current_obs = None     # pinned [slots, obs]: the observation every slot will act on next

    # ---- processes -------------------------------------------------------------------------------------------
    def init_processes(self, n_processes, build_env_fn, collect_metrics_fn=None, spawn_delay=None, render=False,
                       render_delay=None, shm_buffer_size=8192, transport=None):
        """transport: "pipe" (default; multiprocessing pipes, env_worker.py) or "wire" -- the reference's own worker
        protocol (UDP datagrams + a shared float32 slab of shm_buffer_size bytes per worker, batched_agent.py /
        comm_consts.py), so workers written for the reference plug in unchanged.  RLPPO_WORKER_TRANSPORT overrides."""
        _lib.require_device()
        if self.device is None:
            self.device = "cuda:%d" % torch.cuda.current_device()
        methods = mp.get_all_start_methods()
        ctx = mp.get_context("forkserver" if "forkserver" in methods else "spawn")
        self.n_procs = n_processes
        transport = os.environ.get("RLPPO_WORKER_TRANSPORT", transport or "pipe")
        assert transport in ("pipe", "wire"), transport
        self.transport = transport
        if transport == "wire":
            self.shm_size = shm_buffer_size // 4                         # floats per worker (batched_agent_manager.py:437)
            self.shm_buffer = multiprocessing.sharedctypes.RawArray("f", n_processes * self.shm_size)
        for proc_id in range(n_processes):
            if transport == "wire":
                parent = WireConn(self.shm_buffer, proc_id * self.shm_size * 4, self.shm_size)
                p = ctx.Process(target=batched_agent_process,
                                args=(proc_id, parent.endpoint, self.shm_buffer, proc_id * self.shm_size * 4, self.shm_size,
                                      self.seed + proc_id, render and proc_id == 0, render_delay), daemon=True)
                p.start()
            else:
                parent, child = ctx.Pipe(duplex=True)
                p = ctx.Process(target=env_worker, args=(child, proc_id, self.seed + proc_id, render and proc_id == 0,
                                                         render_delay), daemon=True)
                p.start()
                child.close()
            self.processes.append((p, parent))
        if transport == "wire":
            for _, conn in self.processes:
                conn.accept()
        for _, conn in self.processes:
            if spawn_delay is not None:
                time.sleep(spawn_delay)
            conn.send(("init", build_env_fn, collect_metrics_fn))
        first = [self._recv(conn, "reset")[1] for _, conn in self.processes]
        self._agents = [o.shape[0] for o in first]
        self._slot0 = np.concatenate([[0], np.cumsum(self._agents)]).astype(np.int64)
        self.n_slots = int(self._slot0[-1])
        self.obs_dim = int(first[0].shape[1])
        self.ep_rews = [np.zeros(a, np.float64) for a in self._agents]
        self._current_obs = torch.empty((self.n_slots, self.obs_dim), dtype=torch.float32).pin_memory()
        cur = self._current_obs.numpy()
        for p, o in enumerate(first):
            cur[self._slot0[p]:self._slot0[p + 1]] = o
        if self.standardize_obs:
            # NB the reference builds WelfordRunningStat(shape=obs) and seeds it with the reset observations (:377-380)
            self.obs_stats = WelfordRunningStat(self.obs_dim, device=self.device)
            for o in first:
                self.obs_stats.increment(o, o.shape[0])
        self.processes[0][1].send(("shapes",))
        _, obs_size, n_acts, space_type = self._recv(self.processes[0][1], "shapes")
        return obs_size, n_acts, space_type

    @staticmethod
    def _recv(conn, expect):
        msg = conn.recv()
        if msg[0] == "error":
            raise RuntimeError("environment worker failed:\n" + msg[1])
        assert msg[0] == expect, f"expected {expect!r} from worker, got {msg[0]!r}"
        return msg

    # ---- collection ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def collect_timesteps(self, n):
        """Collect at least n timesteps.  Returns ((states, actions, log_probs, rewards, next_states, dones,
        truncated) as DEVICE tensors in the reference's flat layout, collected_metrics, n_collected, elapsed)."""
        from .tick import TickInference
        t1 = time.perf_counter()
        S, D, dev = self.n_slots, self.obs_dim, torch.device(self.device)
        T = (int(n) + S - 1) // S
        tick = getattr(self, "_tick", None)
        if tick is None or tick.policy is not self.policy:
            tick = self._tick = TickInference(self.policy, S, D, standardize=self.standardize_obs)
            tick.obs_host = self._current_obs          # the workers' observations land in the tick's pinned row
        act_w = tick.act_w
        obs_slab = torch.empty((T + 1, S, D), dtype=torch.float32, device=dev)     # what the policy saw (standardised)
        act_slab = torch.empty((T, S) if act_w == 1 else (T, S, act_w), dtype=torch.float32, device=dev)
        logp_slab = torch.empty((T, S), dtype=torch.float32, device=dev)
        rew = np.zeros((T, S), np.float32)
        done = np.zeros((T, S), np.float32)
        trunc = np.zeros((T, S), np.float32)
        collected_metrics = []
        cur = self._current_obs.numpy()
        env_done = np.zeros((T, self.n_procs), bool)

        def push_stats():
            # the reference standardises EVERY feature with the statistics of feature 0
            # (`self.obs_stats.mean[0]`, `.std[0]`, batched_agent_manager.py:233-235): reproduced as is
            if self.standardize_obs:
                tick.set_obs_stats(self.obs_stats.device_mean()[0:1].expand(D), self.obs_stats.device_std()[0:1].expand(D))

        push_stats()
        for t in range(T):
            # one CUDA-graph launch: H2D obs -> standardise/stage -> fused MLP + sample -> D2H actions; then the slab rows
            # (device-to-device, behind the event the host waits on: they never delay the workers)
            tick.enqueue()
            obs_slab[t].copy_(tick.obs_seen, non_blocking=True)
            act_slab[t].copy_(tick.act_dev, non_blocking=True)
            logp_slab[t].copy_(tick.logp_dev, non_blocking=True)
            tick.done_ev.synchronize()
            a = tick.act_host.numpy()
            for p, (_, conn) in enumerate(self.processes):
                conn.send(("act", a[self._slot0[p]:self._slot0[p + 1]].copy()))
            tick_obs = []
            for p, (_, conn) in enumerate(self.processes):
                _, o, r, d, tr, metrics = self._recv(conn, "step")
                lo, hi = self._slot0[p], self._slot0[p + 1]
                if o.shape[0] != hi - lo:
                    raise RuntimeError("environments whose agent count changes on reset are not supported")
                cur[lo:hi] = o
                rew[t, lo:hi] = r
                done[t, lo:hi] = float(d)
                trunc[t, lo:hi] = float(tr)
                env_done[t, p] = d
                tick_obs.append(o)
                if metrics is not None:
                    collected_metrics.append(metrics)
                self.ep_rews[p] += r
                if d or tr:                                   # :327-335
                    if self.average_reward is None:
                        self.average_reward = float(self.ep_rews[p][0])
                    else:
                        for ep_rew in self.ep_rews[p]:
                            self.average_reward = self.average_reward * 0.9 + float(ep_rew) * 0.1
                    self.ep_rews[p][:] = 0
            if self.standardize_obs:                          # :303-315 (stats move between ticks)
                self.steps_since_obs_stats_update += self.n_procs
                if self.steps_since_obs_stats_update > self.steps_per_obs_stats_increment:
                    o = tick_obs[t % len(tick_obs)]
                    self.obs_stats.increment(o, o.shape[0])
                    self.steps_since_obs_stats_update = 0
                    push_stats()
        # the observation after the last step (next_states of tick T-1): staged like a tick's, without inference
        self._stage_only(tick, obs_slab[T])
        out = self.flatten(obs_slab, act_slab, logp_slab, rew, done, trunc, env_done, self._slot0)
        n_collected = int(out[3].shape[0])
        self.cumulative_timesteps += n_collected
        return out, collected_metrics, n_collected, time.perf_counter() - t1

    def _stage_only(self, tick, dst):
        tick.obs_raw.copy_(tick.obs_host, non_blocking=True)
        if self.standardize_obs:
            ops.rows_to_bf16(tick.obs_raw, tick._scratch(), tick.mean_dev, tick.std_dev, tick.clip, dst_f32=dst)
        else:
            dst.copy_(tick.obs_raw, non_blocking=True)
        torch.cuda.current_stream().synchronize()      # obs_host may be rewritten by the next collect_timesteps

    @staticmethod
    def flatten(obs_slab, act_slab, logp_slab, rew, done, trunc, env_done, slot0):
        """Time-major device slabs -> the reference's flat rollout (batched_agent_manager.py:125-172 +
        BatchedTrajectory.get_all, batched_trajectory.py:58-105): completed trajectories in completion
        order (within a tick by process id), then the still-open ones by process id; inside a trajectory agent after agent,
        each agent's steps in time order; the last step of every run is truncated iff it is not done (:145).  The index list
        is built on the host from the done flags (O(T x procs)); the data moves in two device gathers.
          obs_slab [T+1, S, D], act_slab [T, S] or [T, S, A], logp_slab [T, S] device f32;
          rew / done / trunc [T, S] host f32; env_done [T, procs] bool; slot0 [procs+1] first slot of every process."""
        dev = obs_slab.device
        T, S, D = act_slab.shape[0], obs_slab.shape[1], obs_slab.shape[2]
        n_procs = env_done.shape[1]
        runs = []           # (completion tick, proc, t0, t1) with t1 inclusive
        for p in range(n_procs):
            t0 = 0
            for t in np.flatnonzero(env_done[:, p]):
                runs.append((int(t), p, t0, int(t)))
                t0 = int(t) + 1
            if t0 < T:
                runs.append((T + p, p, t0, T - 1))      # still open: after the completed ones, by process id
        runs.sort()
        flat, last_rows, pos = [], [], 0
        for _, p, t0, t1e in runs:
            ts = np.arange(t0, t1e + 1, dtype=np.int64)
            for s in range(int(slot0[p]), int(slot0[p + 1])):
                flat.append(ts * S + s)
                pos += len(ts)
                last_rows.append(pos - 1)
        flat = np.concatenate(flat) if flat else np.zeros(0, np.int64)
        n_collected = int(flat.shape[0])
        # f32 flags: the reference's list mixes np.float32 and Python ints and so comes out float64; the values are the same
        # 0/1 and the scan reads 4 instead of 8 bytes per step
        tr_flat = trunc.reshape(-1)[flat].astype(np.float32)
        dn_flat = done.reshape(-1)[flat]
        last_rows = np.asarray(last_rows, np.int64)
        tr_flat[last_rows] = np.where(dn_flat[last_rows] == 0, 1.0, 0.0)       # :145

        idx = torch.from_numpy(flat).to(dev)
        f32 = lambda: torch.empty(n_collected, dtype=torch.float32, device=dev)  # noqa: E731
        states = torch.empty((n_collected, D), dtype=torch.float32, device=dev)
        next_states = torch.empty_like(states)
        log_probs, rewards, dones = f32(), f32(), f32()
        rew_d, done_d = torch.from_numpy(rew.reshape(-1)).to(dev), torch.from_numpy(done.reshape(-1)).to(dev)
        wide = act_slab.dim() == 3
        actions = torch.empty((n_collected, act_slab.shape[2]), dtype=torch.float32, device=dev) if wide else f32()
        src = _Slab(obs_slab.view(-1, D), None if wide else act_slab.view(-1), logp_slab.view(-1), rew_d, done_d)
        ops.gather_batch(src, idx, out_actions=None if wide else actions, out_logp=log_probs, out_values=rewards,
                         out_adv=dones, out_states=states)
        if wide:
            ops.gather_batch(_Slab(act_slab.view(T * S, -1), None, None, None, None), idx, out_states=actions)
        nxt = _Slab(obs_slab.view(-1, D)[S:], None, None, None, None)           # row (t+1, s)
        ops.gather_batch(nxt, idx, out_states=next_states)
        truncated = torch.from_numpy(tr_flat).to(dev)
        return states, actions, log_probs, rewards, next_states, dones, truncated

    def cleanup(self):
        import traceback
        for p, conn in self.processes:
            try:
                conn.send(("stop",))
            except Exception:
                pass
        for p, conn in self.processes:
            try:
                p.join(timeout=5)
                if p.is_alive():
                    p.terminate()
            except Exception:
                print("Unable to join process")
                traceback.print_exc()
            try:
                conn.close()
            except Exception:
                pass
        self.processes = []


class _Slab:
    """Adapter so ops.gather_batch (written for the experience ring) reads the time-major rollout slabs."""

    def __init__(self, states, actions, log_probs, values, advantages):
        self.states, self.actions, self.log_probs = states, actions, log_probs
        self.values, self.advantages = values, advantages
        self.states_bf16 = None
        self.obs_dim = states.shape[1]
        self.capacity = states.shape[0]
        self.start = 0
