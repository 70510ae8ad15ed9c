"""Worker wire protocol: this package's worker and manager-side parser speak the reference's
UDP + shared-memory-slab protocol (rlgym_ppo/batched_agents/comm_consts.py:3-15, batched_agent.py:60-167,
batched_agent_manager.py:254-300).  tests/golden/wire.npz was recorded by tests/golden/make_golden_wire.py with the
UNMODIFIED reference on both ends in turn: (a) the reference worker driven by this package's WireConn, (b) the reference
manager driving this package's worker.  Here, without the reference: our worker + our parser must reproduce stream (a)
byte for byte, and the in-process environment must agree with both."""
import multiprocessing as mp
import multiprocessing.sharedctypes
import os

import numpy as np
import pytest

from tests import wire_env

GOLD = os.path.join(os.path.dirname(__file__), "golden", "wire.npz")
N_STEPS = 20


def drive(conn, with_metrics=True):
    """init -> reset -> shapes -> N scripted steps; returns the flat record of everything the worker said."""
    conn.send(("init", wire_env.build_env, wire_env.metrics_fn if with_metrics else None))
    rec = {"reset": conn.recv()[1]}
    conn.send(("shapes",))
    rec["shapes"] = np.asarray(conn.recv()[1:], np.int64)
    obs, rew, flags, metrics = [], [], [], []
    for t in range(N_STEPS):
        conn.send(("act", wire_env.action_script(t)))
        _, o, r, d, tr, m = conn.recv()
        obs.append(o)
        rew.append(r)
        flags.append([float(d), float(tr)])
        metrics.append(m if m is not None else np.zeros((0,), np.float32))
    rec.update(obs=np.stack(obs), rew=np.stack(rew), flags=np.asarray(flags, np.float32), metrics=np.stack(metrics))
    return rec


def in_process():
    env = wire_env.build_env()
    rec = {"reset": np.asarray(env.reset(), np.float32)}
    obs, rew, flags, metrics = [], [], [], []
    for t in range(N_STEPS):
        o, r, d, tr, info = env.step(wire_env.action_script(t))
        metrics.append(wire_env.metrics_fn(info["state"]))
        if d or tr:
            o = env.reset()
        obs.append(np.asarray(o, np.float32))
        rew.append(np.asarray(r, np.float32))
        flags.append([float(d), float(tr)])
    rec.update(obs=np.stack(obs), rew=np.stack(rew), flags=np.asarray(flags, np.float32), metrics=np.stack(metrics))
    return rec


def spawn_wire_worker(target, shm_floats=2048):
    from rlgym_ppo_b200.batched_agents.wire import WireConn
    ctx = mp.get_context("forkserver" if "forkserver" in mp.get_all_start_methods() else "spawn")
    shm = multiprocessing.sharedctypes.RawArray("f", shm_floats)
    conn = WireConn(shm, 0, shm_floats, timeout=60.0)
    proc = ctx.Process(target=target, args=(0, conn.endpoint, shm, 0, shm_floats, 123, False, None), daemon=True)
    proc.start()
    conn.accept()
    return proc, conn


def test_comm_consts_are_the_protocol():
    from rlgym_ppo_b200.batched_agents import comm_consts as cc
    g = np.load(GOLD)
    for name in ("ENV_SHAPES_HEADER", "ENV_RESET_STATE_HEADER", "ENV_STEP_DATA_HEADER", "POLICY_ACTIONS_HEADER",
                 "PROC_MESSAGE_SHAPES_HEADER", "STOP_MESSAGE_HEADER"):
        assert list(g["hdr_" + name]) == getattr(cc, name), name
    assert cc.HEADER_LEN == int(g["hdr_HEADER_LEN"])
    assert cc.unpack_message(cc.pack_message([1.5, -2.0, 83775.0])) == [1.5, -2.0, 83775.0]


def test_wire_worker_matches_reference_worker_and_in_process_env():
    from rlgym_ppo_b200.batched_agents.batched_agent import batched_agent_process
    proc, conn = spawn_wire_worker(batched_agent_process)
    try:
        ours = drive(conn)
        conn.send(("stop",))
        proc.join(timeout=20)
        assert not proc.is_alive()
    finally:
        conn.close()
        if proc.is_alive():
            proc.terminate()
    direct = in_process()
    g = np.load(GOLD)
    assert list(ours["shapes"]) == [wire_env.ScriptedEnv.OBS, wire_env.ScriptedEnv.N_ACT, 0]
    for k in ("reset", "obs", "rew", "flags", "metrics"):
        assert ours[k].dtype == np.float32
        assert np.array_equal(ours[k], direct[k]), k                       # the protocol loses nothing
        assert np.array_equal(ours[k], g["refworker_" + k]), k             # == what the reference worker sent (a)
    assert np.array_equal(ours["shapes"], g["refworker_shapes"])
    # (b) the reference manager, driving OUR worker, collected exactly the environment's stream
    assert np.array_equal(g["refmanager_rewards"], g["expected_rewards"])
    assert np.array_equal(g["refmanager_dones"], g["expected_dones"])
    assert np.array_equal(g["refmanager_states"], g["expected_states"])


def test_pipe_and_wire_transports_agree():
    """The default pipe worker and the wire worker are interchangeable behind the manager's send/recv interface."""
    from rlgym_ppo_b200.batched_agents.batched_agent import batched_agent_process
    from rlgym_ppo_b200.batched_agents.env_worker import env_worker
    ctx = mp.get_context("forkserver" if "forkserver" in mp.get_all_start_methods() else "spawn")
    parent, child = ctx.Pipe(duplex=True)
    p1 = ctx.Process(target=env_worker, args=(child, 0, 123, False, None), daemon=True)
    p1.start()
    child.close()
    p2, wire = spawn_wire_worker(batched_agent_process)
    try:
        a, b = drive(parent), drive(wire)
        for k in a:
            assert np.array_equal(np.asarray(a[k]).reshape(-1), np.asarray(b[k]).reshape(-1)), k
    finally:
        for c in (parent, wire):
            try:
                c.send(("stop",))
            except Exception:
                pass
        p1.join(timeout=10)
        p2.join(timeout=10)
        for p in (p1, p2):
            if p.is_alive():
                p.terminate()
        wire.close()
