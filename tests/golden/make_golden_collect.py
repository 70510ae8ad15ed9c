"""
Golden fixture for SURVEY.md 8(a) row a-4: the flat rollout `BatchedAgentManager.collect_timesteps` returns
(batched_agent_manager.py:125-172 + BatchedTrajectory.get_all, batched_trajectory.py:58-105), produced by the UNMODIFIED
reference in the authoring container.

    python tests/golden/make_golden_collect.py        ->  collect.npz

The reference's manager talks to env processes over sockets + shared memory; no env is installable here.  The two
methods that touch the transport (`_send_actions`, `_collect_responses`) are replaced by scripted stand-ins that write
exactly the trajectory fields the originals write (`.state/.action/.log_prob` at :213-215, `.reward/.next_state/.done/
.truncated` at :338-341) from a fixed script, every process answering every pass; everything else -- the collection
loop, `_sync_trajectories`, BatchedTrajectory.update/get_all, the completed-then-open ordering, the forced truncation of
each run's last step (:145), np.asarray of the seven lists -- is the reference's own code, called unmodified.
Cases: (a) 3 processes with 2/1/3 agents, scripted dones and env-level truncations, 37 ticks; (b) 2 processes, no done at
all (only open trajectories); (c) one process, done on every tick; (d) multi-discrete action rows (8 per agent).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, import_reference  # noqa: E402


def script(rng, T, agents, obs_dim, p_done, p_trunc, act_w=1):
    """Per tick and process: observations for tick t (obs[t]) and t+1, actions, log-probs, rewards, done, truncated."""
    P = len(agents)
    S = sum(agents)
    obs = rng.randn(T + 1, S, obs_dim).astype(np.float32)
    acts = rng.randint(0, 3, (T, S) if act_w == 1 else (T, S, act_w)).astype(np.float32)
    logp = (-np.abs(rng.randn(T, S))).astype(np.float32)
    rew = rng.randn(T, S).astype(np.float32)
    done = rng.rand(T, P) < p_done
    trunc = (rng.rand(T, P) < p_trunc) & ~done
    return obs, acts, logp, rew, done, trunc


def run_reference_collect(ref_mod, agents, T, obs, acts, logp, rew, done, trunc):
    from rlgym_ppo.batched_agents import BatchedTrajectory
    from rlgym_ppo.batched_agents.batched_agent_manager import BatchedAgentManager
    P = len(agents)
    slot0 = np.concatenate([[0], np.cumsum(agents)])
    mgr = BatchedAgentManager(None, min_inference_size=P, seed=0, standardize_obs=False)
    mgr.processes = [None] * P
    mgr.trajectory_map = [BatchedTrajectory() for _ in range(P)]
    mgr.completed_trajectories = []
    mgr.current_pids = list(range(P))
    mgr.current_obs = [obs[0, slot0[p]:slot0[p + 1]] for p in range(P)]
    mgr.next_obs = [None] * P
    tick = {"t": 0}

    def send_actions():                                  # fields written at batched_agent_manager.py:213-215
        t = tick["t"]
        for p in range(P):
            lo, hi = slot0[p], slot0[p + 1]
            mgr.trajectory_map[p].action = acts[t, lo:hi]
            mgr.trajectory_map[p].log_prob = torch.from_numpy(logp[t, lo:hi])      # 0-d torch tensors per agent, as there
            mgr.trajectory_map[p].state = mgr.current_obs[p]
        mgr.current_pids = []

    def collect_responses(n_obs_per_inference):          # fields written at :336-341
        t = tick["t"]
        n = 0
        mgr.current_pids = []
        for p in range(P):
            lo, hi = slot0[p], slot0[p + 1]
            mgr.current_pids.append(p)
            mgr.next_obs[p] = obs[t + 1, lo:hi]
            mgr.trajectory_map[p].reward = [x for x in rew[t, lo:hi]]
            mgr.trajectory_map[p].next_state = obs[t + 1, lo:hi]
            mgr.trajectory_map[p].done = np.float32(done[t, p])
            mgr.trajectory_map[p].truncated = np.float32(trunc[t, p])
            n += hi - lo
        tick["t"] += 1
        return [], n

    mgr._send_actions = send_actions
    mgr._collect_responses = collect_responses
    (states, actions, log_probs, rewards, next_states, dones, truncated), _, n_collected, _ = \
        mgr.collect_timesteps(T * sum(agents))
    assert tick["t"] == T
    return states, actions, log_probs, rewards, next_states, dones, truncated, n_collected


def main():
    ref = import_reference()
    print("reference imported from", REF)
    out = {}
    cases = [("mixed", (2, 1, 3), 37, 5, 0.08, 0.05, 1), ("open", (1, 2), 11, 4, 0.0, 0.1, 1),
             ("alldone", (2,), 9, 3, 1.0, 0.0, 1), ("rows", (2, 2), 23, 6, 0.1, 0.0, 8)]
    out["cases"] = np.asarray([c[0] for c in cases])
    for name, agents, T, obs_dim, pd, pt, act_w in cases:
        rng = np.random.RandomState(len(name) * 7 + T)
        obs, acts, logp, rew, done, trunc = script(rng, T, agents, obs_dim, pd, pt, act_w)
        res = run_reference_collect(ref, agents, T, obs, acts, logp, rew, done, trunc)
        out[f"{name}.agents"] = np.asarray(agents)
        for k, v in zip(("obs", "acts", "logp", "rew", "done", "trunc"), (obs, acts, logp, rew, done, trunc)):
            out[f"{name}.in.{k}"] = v
        for k, v in zip(("states", "actions", "log_probs", "rewards", "next_states", "dones", "truncated"), res[:7]):
            v = np.asarray(v)
            out[f"{name}.out.{k}"] = v
        out[f"{name}.n"] = np.asarray([res[7]])
        print(name, "n =", res[7], "| truncated dtype", np.asarray(res[6]).dtype, "| actions", np.asarray(res[1]).shape)
    np.savez_compressed(os.path.join(HERE, "collect.npz"), **out)
    print("collect.npz", os.path.getsize(os.path.join(HERE, "collect.npz")), "B")


if __name__ == "__main__":
    main()
