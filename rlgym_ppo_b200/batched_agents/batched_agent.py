"""Environment worker process speaking the reference's wire protocol (rlgym_ppo/batched_agents/batched_agent.py:1-227):
UDP datagrams on 127.0.0.1 for control, a shared float32 slab for the step data.

    child -> parent   b"0"                                     announces the child's endpoint
    parent -> child   pickle(("initialization_data", build_env_fn, metrics_fn))
    child -> parent   ENV_RESET_STATE_HEADER, n_shape, shape..., obs bytes
    parent -> child   ENV_SHAPES_HEADER                  ->    ENV_SHAPES_HEADER, obs size, n actions, space type
    parent -> child   POLICY_ACTIONS_HEADER, actions...  ->    slab written, then ENV_STEP_DATA_HEADER
    parent -> child   STOP_MESSAGE_HEADER
    slab layout (float32, at shm_offset bytes): [prev_n_agents, done, truncated, len(state_shape), len(metrics_shape),
        metrics_shape..., state_shape..., rewards[prev_n_agents], metrics..., observation (post-reset if the episode ended)]

Same signature and message layout as the reference, own code; environment stepping stays on the host (north_star).  The
action-space type is read from the class name, so `gym` is only needed by the environment itself."""
import pickle
import socket
import time

import numpy as np

from . import comm_consts as cc


def _as_f32(x):
    a = x if isinstance(x, np.ndarray) else np.asarray(x, dtype=np.float32)
    return a if a.dtype == np.float32 else a.astype(np.float32)


def _space_reply(env):
    """obs size, number of actions, space type 0 discrete / 1 multi-discrete / 2 continuous (batched_agent.py:185-214)."""
    kind = type(env.action_space).__name__
    space_type = {"MultiDiscrete": 1.0, "Box": 2.0}.get(kind, 0.0)
    if hasattr(env.action_space, "n"):
        n_acts = float(env.action_space.n)
    else:
        n_acts = float(np.prod(env.action_space.shape))
    return [float(np.prod(env.observation_space.shape)), n_acts, space_type]


def batched_agent_process(proc_id, endpoint, shm_buffer, shm_offset, shm_size, seed, render, render_delay):
    sock = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    sock.bind(("127.0.0.1", 0))
    sock.sendto(b"0", endpoint)
    env, metrics_fn = None, None
    step_header = cc.pack_message(cc.ENV_STEP_DATA_HEADER)
    try:
        while env is None:
            msg = pickle.loads(sock.recv(65536))
            if msg[0] == "initialization_data":
                env, metrics_fn = msg[1](), msg[2]
        try:
            env.action_space.seed(seed)
        except Exception:
            pass
        obs = _as_f32(env.reset())
        shape = [float(s) for s in obs.shape]
        n_agents = int(obs.shape[0]) if obs.ndim > 1 else 1
        sock.sendto(cc.pack_message(cc.ENV_RESET_STATE_HEADER + [float(len(shape))] + shape) + obs.tobytes(), endpoint)

        slab = np.frombuffer(shm_buffer, dtype=np.float32, offset=shm_offset, count=shm_size)
        act_width = None
        while True:
            message = np.frombuffer(sock.recv(65536), dtype=np.float32)
            tag = message[0]
            if tag == cc.POLICY_ACTIONS_HEADER[0]:
                prev_n_agents = n_agents
                data = message[cc.HEADER_LEN:]
                if act_width is None:
                    act_width = data.size // prev_n_agents
                actions = data[:prev_n_agents * act_width].reshape(prev_n_agents, act_width).copy()
                out = env.step(actions)
                if len(out) == 4:
                    nxt, rew, done, info = out
                    truncated = False
                else:
                    nxt, rew, done, truncated, info = out
                rew = np.asarray(rew, dtype=np.float32).reshape(-1)
                if done or truncated:
                    nxt = env.reset()            # the slab carries the post-reset observation (:130-138)
                obs = _as_f32(nxt)
                n_agents = int(obs.shape[0]) if obs.ndim > 1 else 1
                shape = [float(s) for s in obs.shape]
                if metrics_fn is not None:
                    metrics = np.asarray(metrics_fn(info["state"]), dtype=np.float32)
                    mshape = [float(s) for s in metrics.shape]
                else:
                    metrics, mshape = np.empty(0, np.float32), []
                head = [float(prev_n_agents), 1.0 if done else 0.0, 1.0 if truncated else 0.0, float(len(shape)),
                        float(len(mshape))] + mshape + shape
                count = len(head) + rew.size + metrics.size + obs.size
                if count > shm_size:
                    raise RuntimeError(f"step message of {count} floats exceeds the shared-memory slab ({shm_size})")
                o = len(head)
                slab[:o] = head
                slab[o:o + rew.size] = rew
                o += rew.size
                slab[o:o + metrics.size] = metrics.reshape(-1)
                o += metrics.size
                slab[o:o + obs.size] = obs.reshape(-1)
                sock.sendto(step_header, endpoint)
                if render:
                    env.render()
                    if render_delay:
                        time.sleep(render_delay)
            elif tag == cc.ENV_SHAPES_HEADER[0]:
                sock.sendto(cc.pack_message(cc.ENV_SHAPES_HEADER + _space_reply(env)), endpoint)
            elif tag == cc.STOP_MESSAGE_HEADER[0]:
                break
    except Exception:
        import traceback
        print("ERROR IN BATCHED AGENT LOOP", proc_id)
        traceback.print_exc()
    finally:
        sock.close()
        if env is not None and hasattr(env, "close"):
            env.close()
