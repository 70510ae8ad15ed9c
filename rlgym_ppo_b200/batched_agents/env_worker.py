"""Environment worker process: owns one gym-style environment and steps it on request.

Environment stepping stays on host processes (BASELINE.json north_star; the reference's worker is
rlgym_ppo/batched_agents/batched_agent.py).  The worker speaks a small tuple protocol over a multiprocessing
Pipe; observations travel as float32 arrays [n_agents, obs].  It imports neither torch nor CUDA.
"""
import time

import numpy as np


def _as_obs(x):
    a = np.asarray(x, dtype=np.float32)
    return a.reshape(1, -1) if a.ndim == 1 else np.ascontiguousarray(a)


def _space_info(env):
    """(flat observation size, number of actions, action space type 0/1/2) -- batched_agent.py:185-214."""
    kind = type(env.action_space).__name__
    space_type = {"MultiDiscrete": 1, "Box": 2}.get(kind, 0)
    if hasattr(env.action_space, "n"):
        n_acts = int(env.action_space.n)
    else:
        n_acts = int(np.prod(env.action_space.shape))
    return int(np.prod(env.observation_space.shape)), n_acts, space_type


def env_worker(conn, proc_id, seed, render, render_delay):
    env = None
    try:
        tag, build_env_fn, metrics_fn = conn.recv()
        assert tag == "init"
        env = build_env_fn()
        try:
            env.action_space.seed(seed)
        except Exception:
            pass
        obs = _as_obs(env.reset())
        conn.send(("reset", obs))
        while True:
            msg = conn.recv()
            if msg[0] == "act":
                n_agents = obs.shape[0]
                actions = np.asarray(msg[1], dtype=np.float32).reshape(n_agents, -1)
                out = env.step(actions)
                if len(out) == 4:
                    nxt, rew, done, info = out
                    truncated = False
                else:
                    nxt, rew, done, truncated, info = out
                rew = np.asarray(rew, dtype=np.float32).reshape(-1)
                metrics = None
                if metrics_fn is not None:
                    metrics = np.asarray(metrics_fn(info["state"]), dtype=np.float32)
                if done or truncated:
                    nxt = env.reset()      # the observation after a terminal step is the post-reset one (:130-131)
                obs = _as_obs(nxt)
                conn.send(("step", obs, rew, bool(done), bool(truncated), metrics))
                if render:
                    env.render()
                    if render_delay:
                        time.sleep(render_delay)
            elif msg[0] == "shapes":
                conn.send(("shapes",) + _space_info(env))
            elif msg[0] == "stop":
                break
    except (EOFError, KeyboardInterrupt):
        pass
    except Exception:
        import traceback
        print("ERROR IN ENVIRONMENT WORKER", proc_id)
        traceback.print_exc()
        try:
            conn.send(("error", traceback.format_exc()))
        except Exception:
            pass
    finally:
        try:
            conn.close()
        finally:
            if env is not None and hasattr(env, "close"):
                env.close()
