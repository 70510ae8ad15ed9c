"""A pure-NumPy gym-style environment for the Learner integration test (no gym / rlgym_sim needed).
Two agents, 12-dim observations, 5 discrete actions, episodes of 17 steps; reward favours action == obs-argmax % 5."""
import numpy as np


class Discrete:
    def __init__(self, n):
        self.n = n

    def seed(self, s):
        pass


class Box:
    def __init__(self, shape):
        self.shape = shape


class FakeEnv:
    N_AGENTS, OBS, ACT, LEN = 2, 12, 5, 17

    def __init__(self):
        self.action_space = Discrete(self.ACT)
        self.observation_space = Box((self.OBS,))
        self.rng = np.random.RandomState(0)
        self.t = 0
        self.obs = None

    def reset(self):
        self.t = 0
        self.obs = self.rng.randn(self.N_AGENTS, self.OBS).astype(np.float32)
        return self.obs

    def step(self, actions):
        a = np.asarray(actions).reshape(self.N_AGENTS).astype(np.int64)
        assert a.min() >= 0 and a.max() < self.ACT
        rew = (a == (self.obs.argmax(-1) % self.ACT)).astype(np.float32)
        self.t += 1
        self.obs = self.rng.randn(self.N_AGENTS, self.OBS).astype(np.float32)
        done = self.t >= self.LEN
        return self.obs, list(rew), done, {"state": None}

    def close(self):
        pass


def make_env():
    return FakeEnv()


class SlowFakeEnv(FakeEnv):
    """FakeEnv whose steps take a random 0-3 ms: the processes answer in a different order on every pass, which is what
    the manager's asynchronous batching (min_inference_size < n_procs) has to cope with."""

    def __init__(self):
        super().__init__()
        import os
        self._lag = np.random.RandomState(os.getpid())

    def step(self, actions):
        import time
        time.sleep(float(self._lag.rand()) * 0.003)
        return super().step(actions)


def make_slow_env():
    return SlowFakeEnv()
