"""Collection path on the GPU: the device-side flattening against the fixture the reference's
own collect_timesteps produced (tests/golden/collect.npz), and the per-tick inference graph."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_flatten_matches_reference_golden(golden):
    """BatchedAgentManager.flatten (two device gathers over time-major slabs) == the reference's 7-list flattening
    (batched_agent_manager.py:125-172, batched_trajectory.py:58-105), value for value, including multi-agent processes,
    env-level truncation in mid run, open trajectories, done on every tick and action ROWS."""
    from rlgym_ppo_b200.batched_agents import BatchedAgentManager
    g = golden("collect")
    for case in g["cases"]:
        obs, acts, logp, rew, done, trunc = (g[f"{case}.in.{k}"] for k in ("obs", "acts", "logp", "rew", "done", "trunc"))
        agents = [int(a) for a in g[f"{case}.agents"]]
        slot0 = np.concatenate([[0], np.cumsum(agents)]).astype(np.int64)
        T, P = done.shape
        S = int(slot0[-1])
        # per-slot host flags as collect_timesteps keeps them
        done_s, trunc_s = np.zeros((T, S), np.float32), np.zeros((T, S), np.float32)
        for p in range(P):
            done_s[:, slot0[p]:slot0[p + 1]] = done[:, p:p + 1]
            trunc_s[:, slot0[p]:slot0[p + 1]] = trunc[:, p:p + 1]
        out = BatchedAgentManager.flatten(torch.from_numpy(obs).to(DEV), torch.from_numpy(acts).to(DEV),
                                          torch.from_numpy(logp).to(DEV), rew, done_s, trunc_s, done.astype(bool), slot0)
        assert out[0].shape[0] == int(g[f"{case}.n"][0])
        for arr, k in zip(out, ("states", "actions", "log_probs", "rewards", "next_states", "dones", "truncated")):
            want = g[f"{case}.out.{k}"]
            got = arr.cpu().numpy()
            assert got.shape == want.shape, (case, k, got.shape, want.shape)
            assert np.array_equal(got.astype(np.float64), want.astype(np.float64)), (case, k)


@pytest.mark.parametrize("standardize", [False, True])
def test_tick_inference_graph(standardize):
    """A tick is one CUDA-graph replay: fresh Philox numbers on every replay (device-side offset), actions / log-probs equal
    to the eager per-layer path on the same observations and uniforms, standardisation applied from the fixed buffers."""
    from rlgym_ppo_b200 import ops
    from rlgym_ppo_b200.batched_agents.tick import TickInference
    from rlgym_ppo_b200.ppo import DiscreteFF
    torch.manual_seed(3)
    S, D, A = 300, 89, 90
    pol = DiscreteFF(D, A, (256, 256, 256), DEV)
    tick = TickInference(pol, S, D, standardize=standardize)
    rng = np.random.RandomState(0)
    if standardize:
        tick.set_obs_stats(torch.full((D,), 0.25, device=DEV), torch.full((D,), 2.0, device=DEV))
    seen = []
    for t in range(6):
        tick.obs_host.copy_(torch.from_numpy(rng.randn(S, D).astype(np.float32) * 3))
        a = tick.run().clone()
        seen.append((tick.obs_host.clone(), a, tick.logp_dev.cpu().clone(), tick.obs_seen.cpu().clone()))
    assert tick._graph is not None and tick.ticks == 6
    assert int(tick.offset_dev.item()) == 6 * S
    for t, (obs, a, lp, obs_seen) in enumerate(seen):
        want_obs = ((obs - 0.25) / 2.0).clamp(-5, 5) if standardize else obs
        assert torch.allclose(obs_seen, want_obs, atol=1e-6)
        assert float(a.min()) >= 0 and float(a.max()) < A and float(lp.max()) <= 0
        # the same rows through the layer-wise kernels with the same Philox counter: identical actions (up to CDF ties)
        st = pol._stack
        ws = st.workspace(S)
        st.stage_rows(want_obs.to(DEV).contiguous(), ws["x"])
        h = st.forward_hidden(ws["x"], S, ws)
        acts2 = torch.empty(S, device=DEV)
        lp2 = torch.empty(S, device=DEV)
        st.policy_head_sample(h, S, A, seed=pol._seed, offset=t * S, actions_out=acts2, logp_out=lp2)
        same = (acts2.cpu() == a)
        assert same.float().mean() > 0.97, float(same.float().mean())
        assert torch.allclose(lp2.cpu()[same], lp[same], atol=2e-2)
    # consecutive ticks on IDENTICAL observations still draw different actions
    tick.obs_host.copy_(seen[0][0])
    a1 = tick.run().clone()
    a2 = tick.run().clone()
    assert (a1 != a2).float().mean() > 0.3
