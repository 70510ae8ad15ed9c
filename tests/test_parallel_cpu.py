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
    assert [parallel.choose_collective(r, big) for r in (2, 4, 8)] == ["nccl"] * 3       # 60 MB arena: ring all-reduce
    assert parallel.choose_collective(8, big, requested="p2p") == "p2p"                  # explicit request wins
    assert parallel.choose_collective(2, small, env="nccl") == "nccl"
    assert parallel.choose_collective(2, small, requested="p2p", env="nccl") == "p2p"
    assert parallel.choose_collective(16, small, requested="p2p") == "nccl"              # peer mappings stop at the box
    assert parallel.choose_collective(8, big, requested="p2p2") == "p2p2"                # experimental two-shot: opt-in only
    assert "p2p2" not in {parallel.choose_collective(r, n) for r in (2, 4, 8) for n in (small, big)}
    import pytest
    with pytest.raises(ValueError):
        parallel.choose_collective(2, small, requested="ring")
