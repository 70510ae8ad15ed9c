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
  * at the end of collect_timesteps the flat per-trajectory layout the learner expects (completed
    trajectories in completion order, then open ones by process id; each agent's steps in time order; last step of
    every run forced truncated unless done) is produced by ONE index gather on the device.
collect_timesteps therefore returns DEVICE tensors; Learner.add_new_experience consumes them without a copy.
"""
import multiprocessing as mp
import time

import numpy as np
import torch

from .. import _lib, ops
from ..util import WelfordRunningStat
from .env_worker import env_worker


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
                       render_delay=None, shm_buffer_size=8192):
        _lib.require_device()
        if self.device is None:
            self.device = "cuda:%d" % torch.cuda.current_device()
        methods = mp.get_all_start_methods()
        ctx = mp.get_context("forkserver" if "forkserver" in methods else "spawn")
        self.n_procs = n_processes
        for proc_id in range(n_processes):
            parent, child = ctx.Pipe(duplex=True)
            p = ctx.Process(target=env_worker, args=(child, proc_id, self.seed + proc_id, render and proc_id == 0,
                                                     render_delay), daemon=True)
            p.start()
            child.close()
            self.processes.append((p, parent))
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
    def _bf16_scratch(self, rows, width):
        sc = getattr(self, "_bf16_sc", None)
        if sc is None or sc.shape[0] < rows or sc.shape[1] < ops.pad8(width):
            sc = self._bf16_sc = torch.empty((rows, ops.pad8(width)), dtype=torch.bfloat16, device=self.device)
        return sc

    @torch.no_grad()
    def collect_timesteps(self, n):
        """Collect at least n timesteps.  Returns ((states, actions, log_probs, rewards, next_states, dones,
        truncated) as DEVICE tensors in the reference's flat layout, collected_metrics, n_collected, elapsed)."""
        t1 = time.perf_counter()
        S, D, dev = self.n_slots, self.obs_dim, torch.device(self.device)
        T = (int(n) + S - 1) // S
        obs_slab = torch.empty((T + 1, S, D), dtype=torch.float32, device=dev)     # what the policy saw (standardised)
        act_slab = torch.empty((T, S), dtype=torch.float32, device=dev)
        logp_slab = torch.empty((T, S), dtype=torch.float32, device=dev)
        act_host = torch.empty(S, dtype=torch.float32).pin_memory()
        rew = np.zeros((T, S), np.float32)
        done = np.zeros((T, S), np.float32)
        trunc = np.zeros((T, S), np.float32)
        raw = torch.empty((S, D), dtype=torch.float32, device=dev)
        st = self.policy._stack
        collected_metrics = []
        cur = self._current_obs.numpy()
        env_done = np.zeros((T, self.n_procs), bool)

        def stage(t):
            """pinned row -> HBM -> (standardised) f32 slab row + bf16 GEMM operand, all on the device."""
            ws = st.workspace(S)
            if self.standardize_obs:
                raw.copy_(self._current_obs, non_blocking=True)
                # the reference standardises EVERY feature with the statistics of feature 0
                # (`self.obs_stats.mean[0]`, `.std[0]`, batched_agent_manager.py:233-235): reproduced as is
                mean = self.obs_stats.device_mean()[0:1].expand(D).contiguous()
                std = self.obs_stats.device_std()[0:1].expand(D).contiguous()
                if st.exact:
                    # "fp32" mode: standardised f32 rows first (what the trajectory stores), then the split operand
                    ops.rows_to_bf16(raw, self._bf16_scratch(S, D), mean, std, 5.0, dst_f32=obs_slab[t])
                    st.stage_rows(obs_slab[t], ws["x"])
                else:
                    ops.rows_to_bf16(raw, ws["x"], mean, std, 5.0, dst_f32=obs_slab[t])
            else:
                obs_slab[t].copy_(self._current_obs, non_blocking=True)
                st.stage_rows(obs_slab[t], ws["x"])
            return ws

        for t in range(T):
            ws = stage(t)
            st.refresh_operands()
            h = st.forward_hidden(ws["x"], S, ws)
            st.policy_head_sample(h, S, self.policy.n_actions, seed=self.policy._seed, offset=self.policy._offset,
                                  actions_out=act_slab[t], logp_out=logp_slab[t])
            self.policy._offset += S
            act_host.copy_(act_slab[t], non_blocking=True)
            torch.cuda.current_stream().synchronize()
            a = act_host.numpy()
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
        stage(T)      # the observation after the last step: next_states of tick T-1

        # ---- flat per-trajectory order, as one index gather on the device -----------------------------
        runs = []           # (completion tick, proc, t0, t1) with t1 inclusive
        for p in range(self.n_procs):
            t0 = 0
            for t in np.flatnonzero(env_done[:, p]):
                runs.append((int(t), p, t0, int(t)))
                t0 = int(t) + 1
            if t0 < T:
                runs.append((T + p, p, t0, T - 1))      # still open: after the completed ones, by process id
        runs.sort()
        flat = []
        last_rows = []
        pos = 0
        for _, p, t0, t1e in runs:
            ts = np.arange(t0, t1e + 1, dtype=np.int64)
            for s in range(self._slot0[p], self._slot0[p + 1]):
                flat.append(ts * S + s)
                pos += len(ts)
                last_rows.append(pos - 1)
        flat = np.concatenate(flat)
        n_collected = int(flat.shape[0])
        tr_flat = trunc.reshape(-1)[flat].astype(np.float32)
        dn_flat = done.reshape(-1)[flat]
        last_rows = np.asarray(last_rows, np.int64)
        tr_flat[last_rows] = np.where(dn_flat[last_rows] == 0, 1.0, 0.0)       # :145

        idx = torch.from_numpy(flat).to(dev)
        f32 = lambda: torch.empty(n_collected, dtype=torch.float32, device=dev)  # noqa: E731
        states = torch.empty((n_collected, D), dtype=torch.float32, device=dev)
        next_states = torch.empty_like(states)
        actions, log_probs = f32(), f32()
        rew_d, done_d = torch.from_numpy(rew.reshape(-1)).to(dev), torch.from_numpy(done.reshape(-1)).to(dev)
        rewards, dones = f32(), f32()
        src = _Slab(obs_slab.view(-1, D), act_slab.view(-1), logp_slab.view(-1), rew_d, done_d)
        ops.gather_batch(src, idx, out_actions=actions, out_logp=log_probs, out_values=rewards, out_adv=dones,
                         out_states=states)
        nxt = _Slab(obs_slab.view(-1, D)[S:], None, None, None, None)           # row (t+1, s)
        ops.gather_batch(nxt, idx, out_states=next_states)
        truncated = torch.from_numpy(tr_flat).to(dev)

        self.cumulative_timesteps += n_collected
        t2 = time.perf_counter()
        return (states, actions, log_probs, rewards, next_states, dones, truncated), collected_metrics, n_collected, \
            t2 - t1

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
