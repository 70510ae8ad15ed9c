"""Host-side data-parallel bookkeeping (one process per GPU; torch.distributed carries the collectives: NCCL on the
GPUs, gloo in the CPU tests).  No tensor math here beyond the collectives themselves.

The reference has no distributed mode; the seam is its gradient accumulation over minibatch slices
(ppo_learner.py:134-193), which sums (mb/B)-scaled minibatch gradients before ONE clip + Adam step.
"""
import torch
import torch.distributed as dist


def world(group=None):
    """(world_size, rank) of the initialised process group, (1, 0) without one."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def rank_rows(k, batch_size, rank, world_size, mode):
    """Rows of the epoch's permutation that `rank` consumes for optimiser step k: (first, count).

    replicated: every rank holds the same buffer and the same permutation; batch k = perm[k*B : (k+1)*B] is cut into
                world_size consecutive slices -- the reference's minibatch slices with mini_batch_size = B / world_size.
    sharded:    every rank holds its own buffer and permutation and consumes a whole per-rank batch."""
    if mode == "sharded" or world_size == 1:
        return k * batch_size, batch_size
    assert batch_size % world_size == 0, "batch_size must be a multiple of the number of ranks"
    local = batch_size // world_size
    return k * batch_size + rank * local, local


def gae_chunk(n, rank, world_size, align=64):
    """Contiguous chunk [lo, hi) of an n-step flat rollout that `rank` scans when GAE is sharded (the
    recurrence is an associative affine scan, so the step axis can be cut anywhere).  Chunks are `m` steps long (a
    multiple of `align`, so every chunk starts 16-byte aligned in f32 and f64 arrays), the last ones may be short or
    empty.  Returns (lo, hi, m)."""
    m = -(-int(n) // int(world_size))
    m = -(-m // align) * align
    lo = min(int(n), rank * m)
    hi = min(int(n), lo + m)
    return lo, hi, m


def samples_per_step(batch_size, world_size, mode):
    """Number of samples one optimiser step averages over (the B of the 1/B gradient weight)."""
    return batch_size * world_size if (mode == "sharded" and world_size > 1) else batch_size


ONE_SHOT_BYTES = 16 << 20


def choose_collective(world_size, n_params, requested=None, env=None):
    """How data-parallel ranks exchange gradients: "none" (one rank), "p2p" (summed inside the optimiser launch from the
    peers' symmetric-memory arenas: rlppo_norm_clip_adam_peers) or "nccl" (all_reduce on the flat arena).

    The one-shot peer exchange makes every rank read all (R-1) peer arenas, (R-1) * 4 * n bytes over NVLink per step:
    latency-optimal for the example-size nets.  Big arenas take the two-shot form "p2p2" (rlppo_norm_clip_adam_peers2:
    every rank reduces 1/R of the arena from all peers, then reads the reduced slices -- 2 * 4 * n bytes however many ranks
    there are), still inside the optimiser launch, so a whole learn() stays one CUDA graph.  Round 2 on 8 x B200 with the
    60 MB arena of the 2048-2048-1024-1024 nets: 52.4 ms/step against 54.1 ms with the NCCL all-reduce between two graphs;
    validated by tests/dp_check.py (sum equal to NCCL's to rounding, same bits on every rank, 1500 launches under random
    per-rank skew).  Peer mappings exist inside one box only (at most 8 ranks): beyond that, and whenever the mappings
    cannot be set up, NCCL.  `requested` (constructor argument) wins over `env` (RLPPO_DP_COLLECTIVE) which wins over the
    size rule."""
    if world_size <= 1:
        return "none"
    auto = "p2p" if (world_size - 1) * 4 * int(n_params) <= ONE_SHOT_BYTES else "p2p2"
    choice = requested or env or auto
    if choice not in ("p2p", "nccl", "p2p2"):
        raise ValueError(f"dp_collective must be 'p2p', 'nccl' or 'p2p2', got {choice!r}")
    if choice in ("p2p", "p2p2") and world_size > 8:
        choice = "nccl"
    return choice


def allreduce_sum_(t, group=None):
    """In-place sum over ranks (flat gradient arena, metric sums); identity without a process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def report_from_sums(sums):
    """sums: the 8 metric accumulators of rlppo_policy_head_train / rlppo_value_head summed over ranks and minibatches
    -> the four averaged report entries of ppo_learner.py:204-210 (equal-size minibatches: mean of means = total mean)."""
    s = [float(x) for x in sums]
    rows_p = s[4] if s[4] > 0 else 1.0
    rows_v = s[6] if s[6] > 0 else 1.0
    return {"Policy Entropy": s[0] / rows_p, "Mean KL Divergence": s[1] / rows_p, "SB3 Clip Fraction": s[2] / rows_p,
            "Value Function Loss": s[5] / rows_v}
