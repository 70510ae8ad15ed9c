"""Manager-side endpoint of the reference's worker protocol: a connection object with the same send()/recv() tuple
interface as the multiprocessing.Pipe the default transport uses, so BatchedAgentManager does not care which one it holds.

Wire format: batched_agent.py (this package) / rlgym_ppo/batched_agents/batched_agent.py:60-167 (reference worker),
parsed as the reference's manager does (batched_agent_manager.py:254-300, :352-407)."""
import pickle
import socket

import numpy as np

from . import comm_consts as cc

PACKET_MAX_SIZE = 65536


class WireConn:
    """One worker: a UDP socket bound on 127.0.0.1 plus this worker's window of the shared float32 slab."""

    def __init__(self, shm_buffer, shm_offset, shm_size, timeout=120.0):
        self.sock = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
        self.sock.bind(("127.0.0.1", 0))
        self.sock.settimeout(timeout)
        self.endpoint = self.sock.getsockname()
        self.child = None
        self.slab = np.frombuffer(shm_buffer, dtype=np.float32, offset=shm_offset, count=shm_size)
        self._expect = None

    def accept(self):
        """The child's first datagram (b"0") tells us where it listens."""
        _, self.child = self.sock.recvfrom(1)

    # ---- the tuple protocol of env_worker.py, translated ----------------------------------------------------------
    def send(self, msg):
        tag = msg[0]
        if tag == "init":
            self.sock.sendto(pickle.dumps(("initialization_data", msg[1], msg[2])), self.child)
            self._expect = "reset"
        elif tag == "act":
            a = np.ascontiguousarray(msg[1], dtype=np.float32).reshape(-1)
            self.sock.sendto(cc.pack_message(cc.POLICY_ACTIONS_HEADER) + a.tobytes(), self.child)
            self._expect = "step"
        elif tag == "shapes":
            self.sock.sendto(cc.pack_message(cc.ENV_SHAPES_HEADER), self.child)
            self._expect = "shapes"
        elif tag == "stop":
            self.sock.sendto(cc.pack_message(cc.STOP_MESSAGE_HEADER), self.child)
        else:
            raise ValueError(f"unknown message {tag!r}")

    def recv(self):
        while True:
            data = self.sock.recv(PACKET_MAX_SIZE)
            head = np.frombuffer(data, dtype=np.float32, count=cc.HEADER_LEN)
            if self._expect == "reset" and head[0] == cc.ENV_RESET_STATE_HEADER[0]:
                body = np.frombuffer(data, dtype=np.float32, offset=4 * cc.HEADER_LEN)
                nd = int(body[0])
                shape = [int(s) for s in body[1:1 + nd]]
                if nd == 1:
                    shape = [1, shape[0]]
                return ("reset", body[1 + nd:].reshape(shape).copy())
            if self._expect == "shapes" and head[0] == cc.ENV_SHAPES_HEADER[0]:
                obs_size, n_acts, space_type = (int(x) for x in np.frombuffer(data, dtype=np.float32, offset=4 * cc.HEADER_LEN,
                                                                             count=3))
                return ("shapes", obs_size, n_acts, space_type)
            if self._expect == "step" and head[0] == cc.ENV_STEP_DATA_HEADER[0]:
                return self._parse_slab()
            # anything else (a late reply to an earlier request) is dropped, as the reference does (:259-260)

    def _parse_slab(self):
        s = self.slab
        prev_n_agents, done, truncated, nd, n_mshape = int(s[0]), bool(s[1]), bool(s[2]), int(s[3]), int(s[4])
        o = 5
        mshape = [int(x) for x in s[o:o + n_mshape]]
        o += n_mshape
        shape = [int(x) for x in s[o:o + nd]]
        if nd == 1:
            shape = [1, shape[0]]
        o += nd
        rew = s[o:o + prev_n_agents].copy()
        o += prev_n_agents
        n_metrics = int(np.prod(mshape)) if n_mshape else 0
        metrics = s[o:o + n_metrics].reshape(mshape).copy() if n_mshape else None
        o += n_metrics
        obs = s[o:o + int(np.prod(shape))].reshape(shape).copy()
        return ("step", obs, rew, done, truncated, metrics)

    def close(self):
        self.sock.close()
