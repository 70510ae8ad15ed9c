"""Host-side data-parallel logic on CPU: world_size 2 over gloo (spawned processes, 127.0.0.1 rendezvous)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rlgym_ppo_b200 import parallel
        assert parallel.world() == (world, rank)
        B, total = 12, 40
        perm = np.random.RandomState(7).permutation(total)          # same stream on every rank (replicated mode)
        # each rank "computes" a partial gradient = sum of one-hot rows it owns, weighted 1/B like the kernels do
        grads = torch.zeros(total, dtype=torch.float64)
        sums = torch.zeros(8, dtype=torch.float64)
        taken = []
        for k in range(total // B):
            first, count = parallel.rank_rows(k, B, rank, world, "replicated")
            rows = perm[first:first + count]
            taken.append(rows)
            g = torch.zeros(total, dtype=torch.float64)
            g[torch.from_numpy(rows)] += 1.0 / parallel.samples_per_step(B, world, "replicated")
            parallel.allreduce_sum_(g)
            grads += g
            sums[0] += float(rows.sum())
            sums[4] += count
            sums[6] += count
        parallel.allreduce_sum_(sums)
        np.save(os.path.join(out_dir, f"taken{rank}.npy"), np.concatenate(taken))
        torch.save({"grads": grads, "sums": sums, "perm": perm}, os.path.join(out_dir, f"r{rank}.pt"))
        # sharded mode: whole per-rank batches, global weight 1 / (B * world)
        assert parallel.rank_rows(2, B, rank, world, "sharded") == (2 * B, B)
        assert parallel.samples_per_step(B, world, "sharded") == B * world
    finally:
        dist.destroy_process_group()


def test_replicated_slices_partition_every_batch_and_reduce(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(tmp_path / f"r{i}.pt", weights_only=False) for i in range(world)]
    perm, B = r[0]["perm"], 12
    n_batches = len(perm) // B
    # the ranks' slices, interleaved per batch, are exactly the reference's consecutive minibatch slices of the batch
    t = [np.load(tmp_path / f"taken{i}.npy").reshape(n_batches, B // world) for i in range(world)]
    for k in range(n_batches):
        assert np.array_equal(np.concatenate([t[i][k] for i in range(world)]), perm[k * B:(k + 1) * B])
    # after the allreduce every rank holds the single-process gradient: 1/B on every consumed row
    want = torch.zeros(len(perm), dtype=torch.float64)
    want[torch.from_numpy(perm[:n_batches * B])] = 1.0 / B
    for i in range(world):
        assert torch.allclose(r[i]["grads"], want) and torch.equal(r[i]["sums"], r[0]["sums"])
    assert r[0]["sums"][4] == n_batches * B and r[0]["sums"][0] == perm[:n_batches * B].sum()


def test_report_from_sums_and_single_process_defaults():
    from rlgym_ppo_b200 import parallel
    assert parallel.world() == (1, 0)
    assert parallel.rank_rows(3, 10, 0, 1, "replicated") == (30, 10)
    t = torch.arange(4.0)
    assert parallel.allreduce_sum_(t) is t
    rep = parallel.report_from_sums([9.0, 0.5, 2.0, 1.0, 4.0, 8.0, 2.0, 0.0])
    assert rep == {"Policy Entropy": 2.25, "Mean KL Divergence": 0.125, "SB3 Clip Fraction": 0.5,
                   "Value Function Loss": 4.0}
    assert parallel.report_from_sums([0.0] * 8)["Policy Entropy"] == 0.0      # no minibatches: zeros, as the reference


def test_choose_collective():
    from rlgym_ppo_b200 import parallel
    small, big = 332_635, 15_150_171                      # parameter counts of the 256x3 and the 2048-2048-1024-1024 nets
    assert parallel.choose_collective(1, small) == "none"
    assert [parallel.choose_collective(r, small) for r in (2, 4, 8)] == ["p2p"] * 3      # 9.3 MB of peer reads at 8 ranks
    assert [parallel.choose_collective(r, big) for r in (2, 4, 8)] == ["p2p2"] * 3       # 60 MB arena: two-shot peer exchange
    assert parallel.choose_collective(8, big, requested="p2p") == "p2p"                  # explicit request wins
    assert parallel.choose_collective(2, small, env="nccl") == "nccl"
    assert parallel.choose_collective(2, small, requested="p2p", env="nccl") == "p2p"
    assert parallel.choose_collective(16, small, requested="p2p") == "nccl"              # peer mappings stop at the box
    assert parallel.choose_collective(8, big, requested="nccl") == "nccl"
    assert parallel.choose_collective(16, big) == "nccl"                                 # no peer mappings across boxes
    import pytest
    with pytest.raises(ValueError):
        parallel.choose_collective(2, small, requested="ring")


def _gae_worker(rank, world, port, out_dir):
    """The sharded-GAE protocol of Learner.add_new_experience (replicated mode) with the fp64 oracle standing in for the
    kernels: chunk bounds from parallel.gae_chunk, 4-double summaries all-gathered, carry composed rightmost first, chunk
    scanned with its carry, padded chunks all-gathered."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rlgym_ppo_b200 import parallel
        n, g, lam = 1000, 0.99, 0.95
        rng = np.random.RandomState(3)                       # the same rollout on every rank
        rew, val = rng.randn(n), rng.randn(n + 1)
        done = (rng.rand(n) < 0.02).astype(np.float64)
        trunc = ((rng.rand(n) < 0.01) * (1 - done)).astype(np.float64)
        lo, hi, m = parallel.gae_chunk(n, rank, world, align=64)
        assert m % 64 == 0 and lo == min(n, rank * m) and hi == min(n, lo + m)

        def scan(lo, hi, cA, cR):
            adv, ret = np.zeros(hi - lo), np.zeros(hi - lo)
            aA = aR = 1.0
            for t in range(hi - 1, lo - 1, -1):
                nd, nt = 1 - done[t], 1 - trunc[t]
                delta = rew[t] + g * val[t + 1] * nd - val[t]
                cA = delta + g * lam * nd * nt * cA
                cR = rew[t] + g * nd * nt * cR
                aA *= g * lam * nd * nt
                aR *= g * nd * nt
                adv[t - lo], ret[t - lo] = cA, cR
            return adv, ret, aA, aR

        adv0, ret0, aA, aR = scan(lo, hi, 0.0, 0.0)           # summary: x -> b + a x with b = the carry-0 result
        summ = torch.tensor([aA, adv0[0] if hi > lo else 0.0, aR, ret0[0] if hi > lo else 0.0], dtype=torch.float64)
        if hi == lo:
            summ = torch.tensor([1.0, 0.0, 1.0, 0.0], dtype=torch.float64)
        allsum = torch.zeros(4 * world, dtype=torch.float64)
        dist.all_gather_into_tensor(allsum, summ)
        A = R = 0.0
        for k in range(world - 1, rank, -1):                  # rlppo_gae_compose_carry
            A = float(allsum[4 * k + 1] + allsum[4 * k] * A)
            R = float(allsum[4 * k + 3] + allsum[4 * k + 2] * R)
        adv, ret, _, _ = scan(lo, hi, A, R)
        loc = torch.zeros(m, dtype=torch.float64)
        loc[:hi - lo] = torch.from_numpy(adv)
        full = torch.zeros(world * m, dtype=torch.float64)
        dist.all_gather_into_tensor(full, loc)
        want, _, _, _ = scan(0, n, 0.0, 0.0)
        np.save(os.path.join(out_dir, f"gae{rank}.npy"), np.abs(full[:n].numpy() - want).max())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_gae_protocol(tmp_path, world):
    mp.spawn(_gae_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert float(np.load(tmp_path / f"gae{r}.npy")) < 1e-12


def test_gae_chunk_bounds():
    from rlgym_ppo_b200 import parallel
    for n in (1, 63, 64, 1000, 50000, 1003520):
        for world in (1, 2, 3, 8):
            spans = [parallel.gae_chunk(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))        # contiguous, ordered, disjoint
            assert all(s[2] == spans[0][2] and s[2] % 64 == 0 for s in spans)
            assert all(s[0] % 64 == 0 or s[0] == n for s in spans)            # chunk starts are 16-byte aligned
