"""Integration: `Learner(env_fn).learn()` end to end on a B200 with a NumPy fake environment -- the drop-in surface
(constructor keywords, report keys, checkpoint layout, resume), SURVEY.md section 4 'Integration' tier."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_learner_runs_checkpoints_and_resumes(tmp_path, capsys):
    from tests.fake_env import FakeEnv, make_env
    from rlgym_ppo_b200 import Learner
    folder = str(tmp_path / "ckpt")
    kw = dict(n_proc=3, min_inference_size=2, ts_per_iteration=600, exp_buffer_size=1800, ppo_batch_size=600,
              ppo_epochs=2, policy_layer_sizes=(64, 64), critic_layer_sizes=(64, 64), standardize_obs=True,
              standardize_returns=True, checkpoints_save_folder=folder, add_unix_timestamp=False, save_every_ts=1200,
              log_to_wandb=False, ppo_ent_coef=0.01)
    learner = Learner(make_env, timestep_limit=2400, **kw)
    assert learner.ppo_learner.policy.n_actions == FakeEnv.ACT
    learner.learn()
    out = capsys.readouterr().out
    assert out.count("BEGIN ITERATION REPORT") == 4 and "LEARNING LOOP ENCOUNTERED AN ERROR" not in out
    for key in ("Policy Reward", "Policy Entropy", "Value Function Loss", "Mean KL Divergence", "SB3 Clip Fraction",
                "Policy Update Magnitude", "Value Function Update Magnitude", "Collected Steps per Second",
                "Overall Steps per Second", "Timestep Collection Time", "Timestep Consumption Time",
                "PPO Batch Consumption Time", "Total Iteration Time", "Cumulative Model Updates",
                "Cumulative Timesteps", "Timesteps Collected"):
        assert key + ":" in out, key
    ckpts = sorted(int(d) for d in os.listdir(folder))
    assert ckpts and ckpts[-1] >= 2400
    last = os.path.join(folder, str(ckpts[-1]))
    for f in ("PPO_POLICY.pt", "PPO_VALUE_NET.pt", "PPO_POLICY_OPTIMIZER.pt", "PPO_VALUE_NET_OPTIMIZER.pt",
              "BOOK_KEEPING_VARS.json"):
        assert os.path.exists(os.path.join(last, f)), f
    book = json.load(open(os.path.join(last, "BOOK_KEEPING_VARS.json")))
    assert {"cumulative_timesteps", "cumulative_model_updates", "policy_average_reward", "epoch",
            "ts_since_last_save", "reward_running_stats", "obs_running_stats"} <= set(book)
    assert book["reward_running_stats"]["count"] == 4 * 150 and book["cumulative_model_updates"] > 0

    # resume ("latest") picks the newest checkpoint up: weights, Adam state, running stats, counters
    learner2 = Learner(make_env, timestep_limit=3000, **kw)
    assert learner2.agent.cumulative_timesteps == book["cumulative_timesteps"]
    assert learner2.ppo_learner.cumulative_model_updates == book["cumulative_model_updates"]
    assert learner2.return_stats.count == book["reward_running_stats"]["count"]
    sd = torch.load(os.path.join(last, "PPO_POLICY.pt"))
    for k, v in learner2.ppo_learner.policy.state_dict().items():
        assert torch.equal(v.cpu(), sd[k])
    assert int(learner2.ppo_learner._steps[0]) > 0
    learner2.learn()
    assert learner2.agent.cumulative_timesteps >= 3000


@pytest.mark.parametrize("transport", ["pipe", "wire"])
def test_collect_timesteps_layout(transport):
    """collect_timesteps yields the reference's flat layout: every run ends done or truncated, next_states are the
    following observation of the same agent inside a run, rewards match the fake env's rule for the sampled actions.
    transport="wire": the workers are batched_agent_process speaking the reference's UDP + shared-slab protocol
    (tests/test_wire_cpu.py pins it against the reference itself)."""
    from tests.fake_env import FakeEnv, make_env
    from rlgym_ppo_b200.batched_agents import BatchedAgentManager
    from rlgym_ppo_b200.ppo import DiscreteFF
    torch.manual_seed(0)
    mgr = BatchedAgentManager(None, seed=1, standardize_obs=False, device="cuda:0")
    try:
        obs_size, n_act, kind = mgr.init_processes(2, make_env, transport=transport)
        assert mgr.transport == transport
        assert (obs_size, n_act, kind) == (FakeEnv.OBS, FakeEnv.ACT, 0)
        mgr.policy = DiscreteFF(obs_size, n_act, (32,), "cuda:0")
        (states, actions, log_probs, rewards, next_states, dones, truncated), _, n, _ = mgr.collect_timesteps(200)
        assert n >= 200 and states.shape == (n, FakeEnv.OBS) and truncated.shape == (n,)
        s, a, r = states.cpu().numpy(), actions.cpu().numpy(), rewards.cpu().numpy()
        d, tr, ns = dones.cpu().numpy(), truncated.cpu().numpy(), next_states.cpu().numpy()
        assert np.array_equal(r, (a == (s.argmax(-1) % FakeEnv.ACT)).astype(np.float32))
        ends = (d + tr) > 0
        assert ends[-1] and np.all(ns[:-1][~ends[:-1]] == s[1:][~ends[:-1]])
        assert np.all(log_probs.cpu().numpy() <= 0) and mgr.cumulative_timesteps == n
        run_lengths = np.diff(np.concatenate([[-1], np.flatnonzero(ends)]))
        assert run_lengths.max() <= FakeEnv.LEN
    finally:
        mgr.cleanup()


def test_collect_timesteps_asynchronous_batching():
    """min_inference_size < n_procs with environments of uneven speed (batched_agent_manager.py:98-123): inference runs on
    whatever has arrived, fast processes contribute more steps, and the flat layout still holds -- every run ends done or
    truncated, next_states chain inside a run, rewards follow the environment's rule for the recorded (state, action)."""
    from tests.fake_env import FakeEnv, make_slow_env
    from rlgym_ppo_b200.batched_agents import BatchedAgentManager
    from rlgym_ppo_b200.ppo import DiscreteFF
    torch.manual_seed(0)
    mgr = BatchedAgentManager(None, min_inference_size=1, seed=1, standardize_obs=True, device="cuda:0")
    try:
        obs_size, n_act, _ = mgr.init_processes(4, make_slow_env)
        mgr.policy = DiscreteFF(obs_size, n_act, (32,), "cuda:0")
        total = 0
        for _ in range(2):                                    # a second call continues the same environments
            (states, actions, log_probs, rewards, next_states, dones, truncated), _, n, _ = mgr.collect_timesteps(300)
            total += n
            assert n >= 300 and states.shape == (n, FakeEnv.OBS)
            d, tr = dones.cpu().numpy(), truncated.cpu().numpy()
            s, ns = states.cpu().numpy(), next_states.cpu().numpy()
            ends = (d + tr) > 0
            assert ends[-1] and np.all(ns[:-1][~ends[:-1]] == s[1:][~ends[:-1]])
            assert np.all(log_probs.cpu().numpy() <= 0)
            run_lengths = np.diff(np.concatenate([[-1], np.flatnonzero(ends)]))
            assert run_lengths.max() <= FakeEnv.LEN
            a, r = actions.cpu().numpy(), rewards.cpu().numpy()
            assert set(np.unique(r)) <= {0.0, 1.0} and 0 <= a.min() and a.max() < FakeEnv.ACT
        assert mgr.cumulative_timesteps == total
    finally:
        mgr.cleanup()
