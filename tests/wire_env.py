"""A scripted, deterministic two-agent environment for the worker-protocol tests (importable by worker processes)."""
import numpy as np


class _Space:
    def __init__(self, n=None, shape=()):
        if n is not None:
            self.n = n
        self.shape = shape

    def seed(self, seed):
        pass


class ScriptedEnv:
    N_AGENTS, OBS, N_ACT, EP_LEN = 2, 7, 5, 6

    def __init__(self, tag=0):
        self.tag = tag
        self.action_space = _Space(n=self.N_ACT)
        self.observation_space = _Space(shape=(self.OBS,))
        self.t = 0
        self.episode = 0
        self.last = np.zeros(self.N_AGENTS, np.float32)

    def _obs(self):
        o = np.empty((self.N_AGENTS, self.OBS), np.float32)
        for a in range(self.N_AGENTS):
            o[a] = np.sin(np.arange(self.OBS) * 0.37 + self.t * 0.11 + self.episode * 1.3 + a * 0.5 + self.tag)
            o[a, 0] = self.last[a]           # the action the worker delivered comes back in the observation
        return o

    def reset(self):
        self.t = 0
        self.episode += 1
        self.last[:] = -1.0
        return self._obs()

    def step(self, actions):
        actions = np.asarray(actions, np.float32).reshape(self.N_AGENTS, -1)
        self.last = actions[:, 0].copy()
        self.t += 1
        rew = [float(self.t * 0.25 + a - self.last[a] * 0.125) for a in range(self.N_AGENTS)]
        done = self.t >= self.EP_LEN and self.episode % 2 == 1
        truncated = self.t >= self.EP_LEN and not done
        return self._obs(), rew, done, truncated, {"state": (self.t, self.episode)}


    def close(self):
        pass


def build_env():
    return ScriptedEnv()


def metrics_fn(state):
    return np.asarray([[state[0], state[1], 3.0]], np.float32)


def action_script(step, n_agents=ScriptedEnv.N_AGENTS):
    return np.asarray([[(3 * step + a) % ScriptedEnv.N_ACT] for a in range(n_agents)], np.float32)
