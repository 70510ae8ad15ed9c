"""Multi-GPU data-parallel check, run under torchrun (one rank per GPU, NCCL):
  replicated mode: R ranks on the SAME buffer == one rank alone on that buffer (weights, Adam state, report);
  sharded mode:    R ranks on their OWN buffers stay bit-identical replicas, return statistics follow rank 0."""
import contextlib
import io
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def make_buffer(seed, n, dev, obs=89, act=90):
    from rlgym_ppo_b200.ppo import ExperienceBuffer
    r = np.random.RandomState(seed)
    b = ExperienceBuffer(n, 11, dev)
    b.submit_experience(r.randn(n, obs).astype(np.float32), r.randint(0, act, n).astype(np.float32),
                        (-4.5 + 0.3 * r.randn(n)).astype(np.float32), r.randn(n).astype(np.float32),
                        r.randn(n, obs).astype(np.float32), np.zeros(n, np.float32), np.zeros(n),
                        r.randn(n).astype(np.float32), r.randn(n).astype(np.float32))
    return b


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    solo = None
    for r in range(world):                       # every rank must take part in every new_group call
        g = dist.new_group([r])
        if r == rank:
            solo = g
    from rlgym_ppo_b200.ppo import PPOLearner
    from rlgym_ppo_b200.util import WelfordRunningStat
    B, n = 2048, 3 * 2048

    def learner(group, mode):
        torch.manual_seed(5)
        with contextlib.redirect_stdout(io.StringIO()):
            return PPOLearner(89, 90, 0, (256, 256), (256, 256), (0.1, 1.0), B, 2, 3e-4, 3e-4, 0.2, 0.01, B, dev,
                              process_group=group, dp_mode=mode)

    # ---- replicated: same buffer everywhere; compare with a one-rank learner on the same buffer ----
    dp, alone = learner(None, "replicated"), learner(solo, "replicated")
    assert dp.world_size == world and alone.world_size == 1
    p_init = dp._params.clone()
    for it in range(3):                                  # eager, captured, replayed
        rep_dp = dp.learn(make_buffer(100 + it, n, dev))
        rep_1 = alone.learn(make_buffer(100 + it, n, dev))
    # Same per-row math, different order of the fp32 sums (tests/test_learner_gpu.py::test_row_partition_invariance shows
    # the single-GPU version of this statement).  Adam amplifies a 1e-10 wobble on gradient elements of magnitude <= eps,
    # so compare the update: 1e-3 rel-L2, and no element off by as much as one step (lr) after 18 steps.
    upd_dp, upd_1 = dp._params - p_init, alone._params - p_init
    err = float((upd_dp - upd_1).norm() / upd_1.norm())
    assert err < 1e-3, f"replicated DP differs from the single-rank run: rel-L2 of the update {err}"
    assert float((upd_dp - upd_1).abs().max()) < 3e-4
    assert float((dp._m - alone._m).abs().max()) < 1e-6
    for k in ("Policy Entropy", "Mean KL Divergence", "Value Function Loss", "SB3 Clip Fraction"):
        assert abs(rep_dp[k] - rep_1[k]) < 1e-5 * max(1.0, abs(rep_1[k])), (k, rep_dp[k], rep_1[k])
    assert rep_dp["Cumulative Model Updates"] == rep_1["Cumulative Model Updates"] == 3 * 2 * 3

    # ---- sharded: own buffer per rank; replicas must stay identical ----
    sh = learner(None, "sharded")
    for it in range(3):
        rep = sh.learn(make_buffer(1000 * (rank + 1) + it, n, dev))
    gathered = [torch.empty_like(sh._params) for _ in range(world)]
    dist.all_gather(gathered, sh._params)
    for g in gathered[1:]:
        assert torch.equal(g, gathered[0]), "sharded replicas diverged"
    assert np.isfinite(rep["Policy Entropy"]) and rep["Cumulative Model Updates"] == 18

    # ---- return statistics follow rank 0 ----
    st = WelfordRunningStat(1, device=dev)
    st.increment((np.arange(150, dtype=np.float64) + 10.0 * rank), 150)
    st.broadcast_(src=0)
    want = WelfordRunningStat(1, device=dev)
    want.increment(np.arange(150, dtype=np.float64), 150)
    assert st.count == 150 and np.array_equal(st.running_mean, want.running_mean) and np.array_equal(st.std, want.std)
    dist.barrier()
    if rank == 0:
        print(f"dp_check OK world={world} replicated_err={err:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
