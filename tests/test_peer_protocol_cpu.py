"""Model check of the flag protocol of rlppo_norm_clip_adam_peers (csrc/optim.cu, phases 0 / 1 / exit) with host threads
standing in for the ranks: monotonic epochs, "gradients complete" flags before any peer load, "done reading" flags before
a launch may end (= before the owner overwrites its arena).  The CUDA kernel itself is checked on 2+ GPUs by
tests/dp_check.py; this pins the ORDERING argument, including that dropping the exit wait is caught."""
import random
import threading
import time

import pytest

DONE, EPOCH, REDUCED = 32, 64, 96          # flag block layout (uint32 index), as in optim.cu


def _run(world, steps, exit_wait=True, seed=0):
    arenas = [[0] for _ in range(world)]                 # one "gradient" word per rank: the step it belongs to
    flags = [[0] * 128 for _ in range(world)]
    errors, lock = [], threading.Lock()

    def wait(block, idx, epoch, deadline):
        while flags[block][idx] - epoch < 0:
            if time.time() > deadline:
                raise TimeoutError((block, idx, epoch))
            time.sleep(0)

    def rank_main(rank):
        rng = random.Random(seed * 100 + rank)
        for step in range(1, steps + 1):
            time.sleep(rng.random() * 2e-4)
            arenas[rank][0] = step                                           # backward kernels write this rank's gradients
            # ---- the launch ----
            deadline = time.time() + 20
            epoch = flags[rank][EPOCH] + 1
            for r in range(world):
                flags[r][rank] = epoch                                       # phase 0: tell every peer
            for r in range(world):
                wait(rank, r, epoch, deadline)                               # ... and wait for all of them (local polls)
            time.sleep(rng.random() * 2e-4)
            seen = [arenas[r][0] for r in range(world)]                      # phase 1: peer loads
            if seen != [step] * world:
                with lock:
                    errors.append((rank, step, seen))
            for r in range(world):
                flags[r][DONE + rank] = epoch                                # after the grid barrier: done reading
            time.sleep(rng.random() * 2e-4)                                  # phase 2: clip + Adam
            if exit_wait:
                for r in range(world):
                    wait(rank, DONE + r, epoch, deadline)                    # exit: nobody still reads my arena
            flags[rank][EPOCH] = epoch

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    return errors


@pytest.mark.parametrize("world", [2, 4, 8])
def test_protocol_orders_loads_between_writes(world):
    assert _run(world, steps=60, seed=world) == []


def test_dropping_the_exit_wait_is_caught():
    # without the exit wait a fast rank overwrites its arena while a slow peer still reads the previous step
    assert any(_run(4, steps=200, exit_wait=False, seed=s) for s in range(5))


def _run_two_shot(world, steps, reduced_wait=True, seed=0):
    """The experimental two-shot form (rlppo_norm_clip_adam_peers2): (a) each rank sums ITS slice of all arenas into its
    reduced buffer, "slice reduced" flags, (c) every rank reads every owner's slice.  One word per (rank, slice)."""
    arenas = [[0] * world for _ in range(world)]         # arenas[rank][slice] = step the gradient belongs to
    red = [[0] * world for _ in range(world)]            # red[rank][slice]: only slice == rank is read by the peers
    flags = [[0] * 128 for _ in range(world)]
    errors, lock = [], threading.Lock()

    def wait(block, idx, epoch, deadline):
        while flags[block][idx] - epoch < 0:
            if time.time() > deadline:
                raise TimeoutError((block, idx, epoch))
            time.sleep(0)

    def rank_main(rank):
        rng = random.Random(seed * 100 + rank)
        for step in range(1, steps + 1):
            time.sleep(rng.random() * 2e-4)
            arenas[rank][:] = [step] * world
            deadline = time.time() + 20
            epoch = flags[rank][EPOCH] + 1
            for r in range(world):
                flags[r][rank] = epoch
            for r in range(world):
                wait(rank, r, epoch, deadline)
            time.sleep(rng.random() * 2e-4)
            mine = [arenas[r][rank] for r in range(world)]                   # (a) my slice of every arena
            red[rank][rank] = step if mine == [step] * world else -1
            for r in range(world):
                flags[r][REDUCED + rank] = epoch                             # my slice is reduced
            if reduced_wait:
                for r in range(world):
                    wait(rank, REDUCED + r, epoch, deadline)
            time.sleep(rng.random() * 2e-4)
            seen = [red[o][o] for o in range(world)]                         # (c) every owner's slice
            if seen != [step] * world:
                with lock:
                    errors.append((rank, step, seen))
            for r in range(world):
                flags[r][DONE + rank] = epoch
            time.sleep(rng.random() * 2e-4)
            for r in range(world):
                wait(rank, DONE + r, epoch, deadline)
            flags[rank][EPOCH] = epoch

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    return errors


@pytest.mark.parametrize("world", [2, 8])
def test_two_shot_protocol(world):
    assert _run_two_shot(world, steps=60, seed=world) == []


def test_two_shot_without_the_reduced_wait_is_caught():
    assert any(_run_two_shot(4, steps=200, reduced_wait=False, seed=s) for s in range(5))
