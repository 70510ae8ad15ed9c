"""Records tests/golden/wire.npz: the worker wire protocol exercised with the UNMODIFIED reference on each end in turn.

    python tests/golden/make_golden_wire.py        (needs /root/reference; `gym` is replaced by tests/fake_gym)

(a) refworker_*  : the reference worker (rlgym_ppo.batched_agents.batched_agent.batched_agent_process) in a child process,
                   driven by THIS package's manager-side endpoint (batched_agents/wire.py WireConn) with a scripted
                   environment and a scripted action sequence: everything the worker sent back.
(b) refmanager_* : the reference BatchedAgentManager (init_processes + collect_timesteps, its own socket / slab parser /
                   trajectory flattening) with THIS package's worker process substituted for its own, a scripted policy;
                   the flat experience arrays it returned, next to expected_* rebuilt from the environment stepped in-process.
hdr_*            : the header constants of rlgym_ppo/batched_agents/comm_consts.py.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, "/root/reference", os.path.join(ROOT, "tests", "fake_gym")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

from tests import wire_env  # noqa: E402
from tests.test_wire_cpu import N_STEPS, drive, in_process, spawn_wire_worker  # noqa: E402


class ScriptedPolicy:
    """get_action(obs) -> (actions [n, 1] tensor, log_probs [n] tensor): the test's action script, one call per env step."""

    def __init__(self):
        self.step = 0

    def get_action(self, obs, deterministic=False):
        a = torch.from_numpy(wire_env.action_script(self.step, n_agents=obs.shape[0]))
        self.step += 1
        return a, torch.full((obs.shape[0],), -1.25)


def main():
    out = {}
    # ---- header constants -------------------------------------------------------------------------------------------
    from rlgym_ppo.batched_agents import comm_consts as ref_cc
    for name in ("ENV_SHAPES_HEADER", "ENV_RESET_STATE_HEADER", "ENV_STEP_DATA_HEADER", "POLICY_ACTIONS_HEADER",
                 "PROC_MESSAGE_SHAPES_HEADER", "STOP_MESSAGE_HEADER"):
        out["hdr_" + name] = np.asarray(getattr(ref_cc, name), np.float64)
    out["hdr_HEADER_LEN"] = np.asarray(ref_cc.HEADER_LEN)

    # ---- (a) reference worker <- our WireConn -----------------------------------------------------------------------
    from rlgym_ppo.batched_agents.batched_agent import batched_agent_process as ref_worker
    proc, conn = spawn_wire_worker(ref_worker)
    rec = drive(conn)
    conn.send(("stop",))
    proc.join(timeout=20)
    conn.close()
    direct = in_process()
    for k in ("reset", "obs", "rew", "flags", "metrics"):
        assert np.array_equal(rec[k], direct[k]), k
        out["refworker_" + k] = rec[k]
    out["refworker_shapes"] = rec["shapes"]
    print("(a) reference worker driven by WireConn:", N_STEPS, "steps, streams equal the in-process environment")

    # ---- (b) reference manager -> our worker ------------------------------------------------------------------------------
    import rlgym_ppo.batched_agents.batched_agent_manager as ref_mgr
    from rlgym_ppo_b200.batched_agents.batched_agent import batched_agent_process as our_worker
    ref_mgr.batched_agent_process = our_worker          # the only substitution: which function the child process runs
    mgr = ref_mgr.BatchedAgentManager(ScriptedPolicy(), min_inference_size=1, seed=123, standardize_obs=False)
    shapes = mgr.init_processes(1, wire_env.build_env, collect_metrics_fn=wire_env.metrics_fn)
    assert tuple(shapes) == (wire_env.ScriptedEnv.OBS, wire_env.ScriptedEnv.N_ACT, 0), shapes
    n_env_steps = 2 * wire_env.ScriptedEnv.EP_LEN
    exp, metrics, n_collected, _ = mgr.collect_timesteps(n_env_steps * wire_env.ScriptedEnv.N_AGENTS)
    mgr.cleanup()
    states, actions, log_probs, rewards, next_states, dones, truncated = (np.asarray(x) for x in exp)
    # expected layout: completed trajectories in completion order, agent after agent, time order
    env = wire_env.build_env()
    o = np.asarray(env.reset(), np.float32)
    ep, eps = [], []
    t = 0
    while sum(len(e) for e in eps) * wire_env.ScriptedEnv.N_AGENTS < n_collected:
        a = wire_env.action_script(t)
        o2, r, d, tr, _ = env.step(a)
        ep.append((o, a, np.asarray(r, np.float32), np.asarray(o2, np.float32), float(d), float(tr)))
        o = np.asarray(env.reset(), np.float32) if (d or tr) else np.asarray(o2, np.float32)
        t += 1
        if d or tr:
            eps.append(ep)
            ep = []
    e_states, e_rewards, e_dones, e_actions = [], [], [], []
    for e in eps:
        for ag in range(wire_env.ScriptedEnv.N_AGENTS):
            for (s, a, r, s2, d, tr) in e:
                e_states.append(s[ag]); e_rewards.append(r[ag]); e_dones.append(d); e_actions.append(a[ag])
    out.update(refmanager_states=states.astype(np.float32), refmanager_rewards=rewards.astype(np.float32),
               refmanager_dones=dones.astype(np.float32), refmanager_actions=actions.astype(np.float32).reshape(len(rewards), -1),
               refmanager_truncated=truncated.astype(np.float32),
               expected_states=np.asarray(e_states, np.float32), expected_rewards=np.asarray(e_rewards, np.float32),
               expected_dones=np.asarray(e_dones, np.float32), expected_actions=np.asarray(e_actions, np.float32))
    assert np.array_equal(out["refmanager_states"], out["expected_states"])
    assert np.array_equal(out["refmanager_rewards"], out["expected_rewards"])
    assert np.array_equal(out["refmanager_dones"], out["expected_dones"])
    assert np.array_equal(out["refmanager_actions"], out["expected_actions"])
    assert len(metrics) == n_env_steps
    print("(b) reference manager driving our worker:", n_collected, "timesteps, experience equals the in-process environment")
    np.savez_compressed(os.path.join(HERE, "wire.npz"), **out)
    print("wrote", os.path.join(HERE, "wire.npz"))


if __name__ == "__main__":
    main()
