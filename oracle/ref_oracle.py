"""
CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (rlgym_ppo_b200) never imports it and has no CPU fallback.

This is a restatement, in NumPy / torch-CPU tensor math (no autograd, no nn.Module), of the learner-side
hot path of AechPro/rlgym-ppo v1.3.13.  Every function cites the reference file:line it follows.  The
reference's own arithmetic lives in two third-party dependencies that are not under /root/reference:

  * PyTorch (requirements.txt:9 `torch>1.13`; 2.11.0+cu128 in this image): nn.Linear / ReLU / Softmax /
    autograd / torch.optim.Adam / clip_grad_norm_.  Restated here explicitly: Linear is x @ W.T + b, the
    backward is the analytic chain rule of SURVEY.md Appendix A.3, Adam follows the published algorithm
    as implemented by torch.optim.Adam (bias-corrected, eps added after the sqrt, no weight decay).
  * NumPy (requirements.txt:7; 2.3.5 here): np.random.RandomState(seed).permutation -- the legacy MT19937
    stream, frozen by NumPy's compatibility policy.  The oracle calls NumPy itself for it and, separately,
    restates the algorithm (mt19937_permutation) so the product's C implementation can be checked.

PARITY PINNING: the reference has no tests or golden vectors.  The oracle is pinned
against outputs of the reference itself, imported in the authoring container from /root/reference by
tests/golden/make_golden.py; the resulting fixtures are committed under tests/golden/ and
tests/test_oracle_vs_golden.py checks every function here against them.  The two action heads that are next on
the path (MultiDiscreteFF, ContinuousPolicy: SURVEY.md 8(f)-4) are restated and pinned the same way
(tests/golden/make_golden_heads.py -> heads.npz) ahead of their CUDA epilogues.

Scalar-promotion note (compute_gae, WelfordRunningStat): the reference mixes np.float32 scalars, Python
floats and an f64 `truncated` array, so its rounding points depend on the NumPy major version.  The
functions below reproduce what NumPy >= 2 (NEP 50) executes, which is what the goldens were made with:
delta is rounded to f32, the A and R accumulations are f64.  `gae_fp64` is the all-f64 statement
(NumPy 1.x behaviour); the two differ by ~2e-6 abs, inside the 1e-5 tolerance.
"""
from __future__ import annotations

import ctypes
import math
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))

# --------------------------------------------------------------------------------------------------
# GAE  (rlgym_ppo/util/torch_functions.py:36-78)
# --------------------------------------------------------------------------------------------------


def gae_nep50(rews, dones, truncated, values, gamma=0.99, lmbda=0.95, return_std=1.0):
    """torch_functions.py:50-78 with the exact rounding points NumPy>=2 executes.

    rews, dones: f32 [N]; truncated: f64 (or f32) [N]; values: f32-representable [N+1];
    return_std: f32 scalar or None.  Returns (value_targets f32, advantages f32, returns f64).
    """
    rews = np.asarray(rews, dtype=np.float32)
    dones = np.asarray(dones, dtype=np.float32)
    trunc = np.asarray(truncated, dtype=np.float64)
    vals = np.asarray(values, dtype=np.float64)  # python floats in the reference (learner.py:352)
    n = rews.shape[0]
    adv = np.zeros(n, dtype=np.float64)
    rets = np.zeros(n, dtype=np.float64)
    f32 = np.float32
    gl32 = f32(gamma * lmbda)  # python float * np.float32 -> float32 (weak python scalar)
    last_gae = 0.0
    last_ret = 0.0
    for t in range(n - 1, -1, -1):
        nd = f32(1.0) - dones[t]  # torch_functions.py:59 (f32)
        nt = 1.0 - trunc[t]  # :60 (f64)
        if return_std is not None:  # :62-65
            nr = f32(rews[t] / f32(return_std))
            nr = f32(min(max(nr, f32(-10)), f32(10)))
        else:
            nr = rews[t]
        gv = f32(gamma * vals[t + 1])  # python-float product, rounded when it meets the f32 nd
        pred = f32(nr + f32(gv * nd))  # :67
        delta = f32(pred - f32(vals[t]))  # :68
        last_ret = float(rews[t]) + last_ret * gamma * float(nd) * nt  # :69 (f64)
        rets[t] = last_ret
        last_gae = float(delta) + float(f32(gl32 * nd)) * nt * last_gae  # :72 (f64)
        adv[t] = last_gae
    advantages = adv.astype(np.float32)  # :76
    value_targets = (vals[:-1] + adv).astype(np.float32)  # :77
    return value_targets, advantages, rets


def gae_fp64(rews, dones, truncated, values, gamma=0.99, lmbda=0.95, return_std=1.0):
    """the all-f64 statement of torch_functions.py:50-78 (NumPy 1.x promotion)."""
    r = np.asarray(rews, dtype=np.float64)
    d = np.asarray(dones, dtype=np.float64)
    tr = np.asarray(truncated, dtype=np.float64)
    v = np.asarray(values, dtype=np.float64)
    n = r.shape[0]
    adv = np.zeros(n)
    rets = np.zeros(n)
    la = 0.0
    lr = 0.0
    for t in range(n - 1, -1, -1):
        nd = 1.0 - d[t]
        nt = 1.0 - tr[t]
        nr = min(max(r[t] / float(return_std), -10.0), 10.0) if return_std is not None else r[t]
        delta = nr + gamma * v[t + 1] * nd - v[t]
        lr = r[t] + lr * gamma * nd * nt
        rets[t] = lr
        la = delta + gamma * lmbda * nd * nt * la
        adv[t] = la
    return (v[:-1] + adv).astype(np.float32), adv.astype(np.float32), rets


_C_LIB = None


def _c_lib():
    """The C restatement (oracle/gae_oracle.c), built by oracle/Makefile into oracle/_build/."""
    global _C_LIB
    if _C_LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
        lib = ctypes.CDLL(path)
        P = ctypes.c_void_p
        lib.oracle_gae_nep50.argtypes = [P, P, P, P, ctypes.c_int64, ctypes.c_double, ctypes.c_double,
                                         ctypes.c_int, ctypes.c_float, P, P, P]
        lib.oracle_gae_nep50.restype = None
        lib.oracle_welford_update.argtypes = [P, P, P, P, ctypes.c_int64]
        lib.oracle_welford_update.restype = None
        lib.oracle_mt19937_permutation.argtypes = [P, P, ctypes.c_int64, P]
        lib.oracle_mt19937_permutation.restype = None
        _C_LIB = lib
    return _C_LIB


def gae_nep50_c(rews, dones, truncated, values, gamma=0.99, lmbda=0.95, return_std=1.0):
    """Same as gae_nep50, through the plain-C restatement (fast enough for 1e8 steps)."""
    lib = _c_lib()
    r = np.ascontiguousarray(rews, dtype=np.float32)
    d = np.ascontiguousarray(dones, dtype=np.float32)
    tr = np.ascontiguousarray(truncated, dtype=np.float64)
    v = np.ascontiguousarray(values, dtype=np.float32)
    n = r.shape[0]
    adv = np.empty(n, np.float32)
    vt = np.empty(n, np.float32)
    rets = np.empty(n, np.float64)
    lib.oracle_gae_nep50(r.ctypes.data, d.ctypes.data, tr.ctypes.data, v.ctypes.data, n, gamma, lmbda,
                         0 if return_std is None else 1, 1.0 if return_std is None else float(return_std),
                         adv.ctypes.data, vt.ctypes.data, rets.ctypes.data)
    return vt, adv, rets


# --------------------------------------------------------------------------------------------------
# WelfordRunningStat  (rlgym_ppo/util/running_stats.py:15-98)
# --------------------------------------------------------------------------------------------------


class WelfordOracle:
    """running_stats.py:21-69.  State f32, intermediates f64 when the sample is f64 (NumPy>=2)."""

    def __init__(self, shape):
        self.shape = shape
        self.mean = np.zeros(shape, np.float32)  # :25
        self.m2 = np.zeros(shape, np.float32)  # :26 ("running_variance" holds M2)
        self.count = 0

    def update(self, sample):  # :37-46
        sample = np.asarray(sample)
        cc = self.count
        self.count += 1
        if sample.dtype == np.float64:
            delta = (sample.astype(np.float64) - self.mean.astype(np.float64)).reshape(self.mean.shape)
            delta_n = delta / self.count
            self.mean = (self.mean.astype(np.float64) + delta_n).astype(np.float32)
            self.m2 = (self.m2.astype(np.float64) + delta * delta_n * cc).astype(np.float32)
        else:
            s = sample.astype(np.float32)
            delta = (s - self.mean).reshape(self.mean.shape)
            delta_n = (delta / np.float32(self.count)).astype(np.float32)
            self.mean = self.mean + delta_n
            self.m2 = self.m2 + delta * delta_n * np.float32(cc)

    def increment(self, samples, num):  # :30-35
        if num > 1:
            for i in range(num):
                self.update(samples[i])
        else:
            self.update(samples)

    def get_mean(self):  # :54-58
        if self.count < 2:
            return np.zeros(self.shape, np.float32)
        return self.mean

    def get_std(self):  # :60-69
        if self.count < 2:
            return np.ones(self.shape, np.float32)
        var = self.m2 / np.float32(self.count - 1)
        var = np.where(var == 0, np.float32(1.0), var)
        return np.sqrt(var).astype(np.float32)

    def merge(self, other_mean, other_m2, other_count):  # :71-98
        if other_count == 0:
            return
        om = np.asarray(other_mean, np.float32).reshape(self.mean.shape)
        ov = np.asarray(other_m2, np.float32).reshape(self.m2.shape)
        count = self.count + other_count
        d = om - self.mean
        self.m2 = self.m2 + ov + d * d * self.count * other_count / count
        self.mean = (self.count * self.mean + other_count * om) / count
        self.count = count


# --------------------------------------------------------------------------------------------------
# ExperienceBuffer  (rlgym_ppo/ppo/experience_buffer.py:17-102)
# --------------------------------------------------------------------------------------------------

FIELDS = ("states", "actions", "log_probs", "rewards", "next_states", "dones", "truncated", "values",
          "advantages")


def fifo_cat(old, new, size):
    """experience_buffer.py:17-37 `_cat`: keep the newest `size` rows of old (+) new."""
    if len(new) > size:
        return new[-size:].copy()
    if len(new) == size:
        return new
    if len(old) + len(new) > size:
        return np.concatenate((old[len(new) - size:], new), 0)
    return np.concatenate((old, new), 0)


class BufferOracle:
    def __init__(self, max_size, seed):  # :39-52
        self.max_size = max_size
        self.rng = np.random.RandomState(seed)
        self.f = {k: np.zeros((0,), np.float32) for k in FIELDS}

    def submit(self, **fields):  # :54-80
        for k in FIELDS:
            new = np.asarray(fields[k], dtype=np.float32)
            old = self.f[k]
            if old.shape[0] == 0 and new.ndim > 1:
                old = np.zeros((0,) + new.shape[1:], np.float32)
            self.f[k] = fifo_cat(old, new, self.max_size)

    def batches(self, batch_size):  # :89-102 ; yields (indices, 5-tuple) per batch
        total = self.f["rewards"].shape[0]
        idx = self.rng.permutation(total)  # :98 one permutation per call (= per epoch)
        start = 0
        while start + batch_size <= total:  # :100 remainder dropped
            ii = idx[start:start + batch_size]
            yield ii, (self.f["actions"][ii], self.f["log_probs"][ii], self.f["states"][ii],
                       self.f["values"][ii], self.f["advantages"][ii])  # :82-87 field order
            start += batch_size


# --- legacy MT19937 permutation, restated (numpy/random/mtrand.pyx RandomState.permutation -> shuffle,
#     numpy/random/src/legacy + mt19937.c: rk_interval masked rejection, Fisher-Yates from the top) -----


def mt19937_permutation(key, pos, n):
    """Pure-Python restatement (small n only).  key: uint32[624], pos: int.  Returns (perm, key, pos)."""
    mt = [int(x) for x in key]

    def gen():
        nonlocal pos
        if pos == 624:
            for k in range(624):
                y = (mt[k] & 0x80000000) | (mt[(k + 1) % 624] & 0x7FFFFFFF)
                mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ (0x9908B0DF if (y & 1) else 0)
            pos = 0
        y = mt[pos]
        pos += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF

    def interval(mx):
        if mx == 0:
            return 0
        mask = mx
        for s in (1, 2, 4, 8, 16, 32):
            mask |= mask >> s
        if mx <= 0xFFFFFFFF:
            while True:
                v = gen() & mask
                if v <= mx:
                    return v
        while True:
            v = ((gen() << 32) | gen()) & mask
            if v <= mx:
                return v

    x = list(range(n))
    for i in range(n - 1, 0, -1):
        j = interval(i)
        x[i], x[j] = x[j], x[i]
    return np.asarray(x, np.int64), np.asarray(mt, np.uint32), pos


def mt19937_permutation_c(key, pos, n):
    lib = _c_lib()
    key = np.ascontiguousarray(key, dtype=np.uint32).copy()
    p = np.asarray([pos], dtype=np.int32)
    out = np.empty(n, np.int64)
    lib.oracle_mt19937_permutation(key.ctypes.data, p.ctypes.data, n, out.ctypes.data)
    return out, key, int(p[0])


# --------------------------------------------------------------------------------------------------
# Rollout flattening  (batched_agent_manager.py:100-172, batched_trajectory.py:20-105)
# --------------------------------------------------------------------------------------------------


def flatten_rollout(obs, acts, logp, rew, done, trunc, agents):
    """What collect_timesteps returns when every process answers every pass.  Inputs are time-major:
    obs [T+1, S, D], acts [T, S] or [T, S, A], logp / rew [T, S], done / trunc [T, P] (per ENV), agents = agents per
    process.  A process's open trajectory takes (state, action, log_prob, reward, next_state, done, truncated) per tick
    (:213-215, :338-341); update() closes it when done (batched_trajectory.py:36-54) and `_sync_trajectories` moves it to
    the completed list in process order within a tick (:174-178); at the end the open ones follow by process id
    (:126-128).  get_all() emits agent after agent, each agent's steps in time order (:70-101); the last step of every
    such run is truncated iff not done (:145).  Returns the seven arrays (truncated as float64, like np.asarray of the
    reference's mixed list)."""
    T = acts.shape[0]
    P = len(agents)
    slot0 = np.concatenate([[0], np.cumsum(agents)])
    open_steps = [[] for _ in range(P)]
    completed = []
    for t in range(T):
        for p in range(P):
            open_steps[p].append(t)
            if done[t, p]:
                completed.append((p, open_steps[p]))
                open_steps[p] = []
    for p in range(P):
        completed.append((p, open_steps[p]))
    cols = [[] for _ in range(7)]
    for p, steps in completed:
        if not steps:
            continue
        for s in range(slot0[p], slot0[p + 1]):
            run = [[obs[t, s] for t in steps], [acts[t, s] for t in steps], [logp[t, s] for t in steps],
                   [rew[t, s] for t in steps], [obs[t + 1, s] for t in steps], [np.float32(done[t, p]) for t in steps],
                   [float(trunc[t, p]) for t in steps]]
            run[6][-1] = 1.0 if run[5][-1] == 0 else 0.0
            for c, r in zip(cols, run):
                c.extend(r)
    dt = (np.float32,) * 6 + (np.float64,)
    return tuple(np.asarray(c, dtype=d) for c, d in zip(cols, dt))


# --------------------------------------------------------------------------------------------------
# Networks  (discrete_policy.py:22-42, value_estimator.py:19-36): params = [W0,b0,W1,b1,...], W [out,in]
# --------------------------------------------------------------------------------------------------


def mlp_forward(params, x, quant=None):
    """Linear/ReLU stack; returns (pre-softmax output, list of layer inputs).  `quant` optionally rounds
    both GEMM operands (e.g. to bf16) so the oracle can mirror the tensor-core kernels' rounding points."""
    q = quant if quant is not None else (lambda t: t)
    acts = []
    h = x
    nl = len(params) // 2
    for i in range(nl):
        acts.append(h)
        h = q(h) @ q(params[2 * i]).t() + params[2 * i + 1]
        if i < nl - 1:
            h = torch.relu(h)
            if quant is not None:
                h = q(h)  # the kernels store hidden activations in bf16
    return h, acts


def policy_probs(params, x, quant=None):
    """discrete_policy.py:35-42 get_output: softmax over the last Linear."""
    z, _ = mlp_forward(params, x, quant)
    return torch.softmax(z, dim=-1)


def action_logprob(params, x, actions, quant=None):
    """discrete_policy.py:54,60: clamp(1e-11,1) -> log -> gather."""
    p = torch.clamp(policy_probs(params, x, quant), 1e-11, 1.0)
    return torch.log(p).gather(-1, actions.view(-1, 1).long()).flatten()


def sample_inverse_cdf(probs, u):
    """Categorical sample by inverse CDF on the clamped probabilities (the product's K-a kernel contract):
    action = first j with cumsum(p)[j] > u * sum(p); torch.multinomial's own stream is not reproducible
    outside torch, so the sampler is pinned on injected uniforms + a chi-square test instead."""
    p = torch.clamp(probs, 1e-11, 1.0).double()
    c = torch.cumsum(p, -1)
    thr = (u.double() * c[:, -1]).unsqueeze(-1)
    a = (c > thr).float().argmax(-1)
    return a


def mlp_backward(params, acts, dout, quant=None):
    """Analytic backward of the Linear/ReLU stack.  Returns grads in params order."""
    q = quant if quant is not None else (lambda t: t)
    nl = len(params) // 2
    grads = [None] * len(params)
    g = dout
    for i in range(nl - 1, -1, -1):
        gq = q(g)
        grads[2 * i] = gq.t() @ q(acts[i])
        grads[2 * i + 1] = gq.sum(0)   # quant: the kernels sum the rounded tile the weight-gradient GEMM reads
        if i > 0:
            g = (gq @ q(params[2 * i])) * (acts[i] > 0).to(g.dtype)  # acts[i] = relu output of layer i-1
    return grads


# --------------------------------------------------------------------------------------------------
# The other two action heads (ppo_learner.py:36-50 selects by policy_type):
#   1  MultiDiscreteFF   multi_discrete_policy.py:17-89 + torch_functions.MultiDiscreteRolv (:81-122)
#   2  ContinuousPolicy  continuous_policy.py:21-120   + torch_functions.MapContinuousToAction (:15-33)
# Each returns (logp [mb], entropy scalar, backward) where backward(d_logp [mb], d_entropy scalar) is the
# analytic gradient with respect to the last Linear's output.  torch.distributions is restated, not called.
# --------------------------------------------------------------------------------------------------
ROLV_BINS = (3, 3, 3, 3, 3, 2, 2, 2)  # multi_discrete_policy.py:21


def head_multi_discrete(z, acts):
    """21 logits -> 8 categoricals (5 triplets, 3 duets; the reference pads the duets with -inf to triplets,
    which changes nothing: a -inf logit has probability 0 and, with Categorical.entropy's clamp of the
    logits to finfo.min, contributes 0 * finite = 0).  log_prob and entropy are SUMMED over the 8
    distributions (:115, :121), the entropy is then averaged over the minibatch
    (multi_discrete_policy.py:89)."""
    a = acts.view(z.shape[0], -1).long()
    logp = torch.zeros(z.shape[0], dtype=z.dtype)
    ent = torch.zeros(z.shape[0], dtype=z.dtype)
    groups = []
    start = 0
    for gi, nb in enumerate(ROLV_BINS):
        zg = z[:, start:start + nb]
        lsm = zg - torch.logsumexp(zg, -1, keepdim=True)  # Categorical(logits=...) normalisation
        pg = torch.softmax(lsm, -1)  # logits_to_probs
        hg = -(lsm * pg).sum(-1)
        logp = logp + lsm.gather(-1, a[:, gi:gi + 1]).flatten()
        ent = ent + hg
        groups.append((start, nb, lsm, pg, hg))
        start += nb

    def backward(d_logp, d_entropy):
        dz = torch.zeros_like(z)
        mb = z.shape[0]
        for gi, (st, nb, lsm, pg, hg) in enumerate(groups):
            onehot = torch.zeros_like(pg).scatter_(-1, a[:, gi:gi + 1], 1.0)
            # d logp / dz = onehot - p ;  dH/dz_i = -p_i (log p_i + H) ; entropy.mean() spreads d_entropy / mb
            dz[:, st:st + nb] = d_logp.unsqueeze(-1) * (onehot - pg) + \
                (d_entropy / mb) * (-pg * (lsm + hg.unsqueeze(-1)))
        return dz

    return logp, ent.mean(), backward


def head_continuous(u, acts, var_min=0.1, var_max=1.0):
    """u = output of the last Linear (2N columns).  continuous_policy.py:37 Tanh, torch_functions.py:24-33 affine map of
    the second half onto [var_min, var_max], :40-59 the four-term log-pdf as the reference writes it (f32), summed over
    the N actions (:112); entropy = Normal.entropy() = 0.5 + 0.5 log(2 pi) + log(std), averaged over ALL mb * N
    elements (:117-118)."""
    t = torch.tanh(u)
    n = t.shape[-1] // 2
    m = (var_max - var_min) / 2.0  # torch_functions.py:27-28
    b = var_min + m
    mean, std = t[:, :n], t[:, n:] * m + b
    x = acts.view(u.shape[0], n).to(u.dtype)
    msq, ssq, xsq = mean * mean, std * std, x * x
    term1 = -torch.divide(msq, (2 * ssq))
    term2 = torch.divide(mean * x, ssq)
    term3 = -torch.divide(xsq, (2 * ssq))
    term4 = torch.log(1 / torch.sqrt(2 * np.pi * ssq))
    logp = (term1 + term2 + term3 + term4).sum(dim=1)
    ent = (0.5 + 0.5 * math.log(2 * math.pi) + torch.log(std)).mean()

    def backward(d_logp, d_entropy):
        d = d_logp.unsqueeze(-1)
        d_mean = d * (x - mean) / ssq
        d_std = d * ((x - mean) ** 2 / (ssq * std) - 1.0 / std) + d_entropy / (std.numel() * std)
        dt = torch.cat([d_mean, d_std * m], dim=-1)
        return dt * (1.0 - t * t)  # Tanh

    return logp, ent, backward


def ppo_minibatch_head(policy_type, pol, val, obs, acts, old_logp, targets, adv, clip, ent_coef, batch_size,
                       var_range=(0.1, 1.0), quant=None):
    """ppo_minibatch for policy_type 1 / 2: the same loss block (ppo_learner.py:146-185) around another head."""
    w = 1.0 / batch_size
    mb = obs.shape[0]
    z, pacts = mlp_forward(pol, obs, quant)
    if policy_type == 1:
        logp, entropy, head_bwd = head_multi_discrete(z, acts)
    else:
        logp, entropy, head_bwd = head_continuous(z, acts, *var_range)
    log_ratio = logp - old_logp
    ratio = torch.exp(log_ratio)
    clipped = torch.clamp(ratio, 1.0 - clip, 1.0 + clip)
    kl = ((ratio - 1) - log_ratio).mean()
    clip_frac = ((ratio - 1).abs() > clip).float().mean()
    s1, s2 = ratio * adv, clipped * adv
    policy_loss = -torch.min(s1, s2).mean()
    v, vacts = mlp_forward(val, obs, quant)
    v = v.flatten()
    value_loss = ((v - targets) ** 2).mean()
    in_range = (ratio >= 1.0 - clip) & (ratio <= 1.0 + clip)
    lt, eq, gt = (s1 < s2).to(z.dtype), (s1 == s2).to(z.dtype), (s1 > s2).to(z.dtype)
    d_ratio = -w * adv * ((lt + 0.5 * eq) + (gt + 0.5 * eq) * in_range.to(z.dtype))
    # ppo_loss = (policy_loss - ent_coef * entropy) * (mb / B): d/d entropy = -ent_coef * mb / B
    dz = head_bwd(d_ratio * ratio, -ent_coef * mb * w)
    pg = mlp_backward(pol, pacts, dz, quant)
    vg = mlp_backward(val, vacts, (2.0 * w * (v - targets)).unsqueeze(-1), quant)
    metrics = dict(entropy=float(entropy), kl=float(kl), clip_fraction=float(clip_frac),
                   value_loss=float(value_loss), policy_loss=float(policy_loss))
    return pg, vg, metrics


def ppo_minibatch(pol, val, obs, acts, old_logp, targets, adv, clip, ent_coef, batch_size, quant=None):
    """ppo_learner.py:146-185 + discrete_policy.py:64-80 forward, and the analytic backward of
    SURVEY.md A.3.  Returns (policy grads, value grads, metrics dict with per-minibatch means)."""
    mb = obs.shape[0]
    w = 1.0 / batch_size  # (1/mb) * (mb/B), ppo_learner.py:175-177
    z, pacts = mlp_forward(pol, obs, quant)
    s = torch.softmax(z, -1)  # discrete_policy.py:30
    p = torch.clamp(s, 1e-11, 1.0)  # :74
    logp_all = torch.log(p)  # :76
    a = acts.view(-1, 1).long()  # :71
    logp = logp_all.gather(-1, a).flatten()  # :77
    ent_b = -(logp_all * p).sum(-1)  # :78
    entropy = ent_b.mean()  # :80
    log_ratio = logp - old_logp
    ratio = torch.exp(log_ratio)  # ppo_learner.py:153
    clipped = torch.clamp(ratio, 1.0 - clip, 1.0 + clip)  # :154-156
    kl = ((ratio - 1) - log_ratio).mean()  # :160-162
    clip_frac = ((ratio - 1).abs() > clip).float().mean()  # :165-169
    s1 = ratio * adv
    s2 = clipped * adv
    policy_loss = -torch.min(s1, s2).mean()  # :172-174
    v, vacts = mlp_forward(val, obs, quant)
    v = v.flatten()
    value_loss = ((v - targets) ** 2).mean()  # :176 (MSELoss), before the minibatch_ratio factor

    # ---- backward (A.3) ----
    # torch.min(a,b) backward: gradient to `a` where a < b, split 0.5/0.5 where a == b (ATen
    # min.other derivative: grad.masked_fill(self > other, 0) / (1 + (self == other))) -- the tie is the
    # common case here: whenever the ratio is inside the clip range, s1 == s2 exactly.
    in_range = (ratio >= 1.0 - clip) & (ratio <= 1.0 + clip)  # clamp passes gradient inside (inclusive)
    lt = (s1 < s2).to(z.dtype)
    eq = (s1 == s2).to(z.dtype)
    gt = (s1 > s2).to(z.dtype)
    d_s1 = lt + 0.5 * eq  # dmin/ds1
    d_s2 = gt + 0.5 * eq  # dmin/ds2
    d_ratio = -w * adv * (d_s1 + d_s2 * in_range.to(z.dtype))
    d_logp = d_ratio * ratio
    g = ent_coef * w * (logp_all + 1.0)  # d/dp of -ent_coef*w*sum(-p log p)  (both factors are the clamped p)
    g = g.scatter_add(-1, a, (d_logp / p.gather(-1, a).flatten()).unsqueeze(-1))
    g = g * ((s >= 1e-11) & (s <= 1.0)).to(z.dtype)  # clamp mask on the unclamped softmax
    dz = s * (g - (g * s).sum(-1, keepdim=True))  # softmax Jacobian
    pg = mlp_backward(pol, pacts, dz, quant)
    dv = (2.0 * w * (v - targets)).unsqueeze(-1)
    vg = mlp_backward(val, vacts, dv, quant)
    metrics = dict(entropy=float(entropy), kl=float(kl), clip_fraction=float(clip_frac),
                   value_loss=float(value_loss), policy_loss=float(policy_loss))
    return pg, vg, metrics


def clip_grad_norm(grads, max_norm=0.5):
    """torch.nn.utils.clip_grad_norm_ (called at ppo_learner.py:187-190): norm of per-tensor norms,
    coef = max_norm / (total + 1e-6) clamped to 1."""
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g) for g in grads]))
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return [g * coef for g in grads], float(total)


class AdamOracle:
    """torch.optim.Adam as configured at ppo_learner.py:56-59 (betas (0.9,0.999), eps 1e-8, wd 0)."""

    def __init__(self, params, lr, b1=0.9, b2=0.999, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, b1, b2, eps
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]
        self.step_count = 0

    def step(self, params, grads):
        self.step_count += 1
        t = self.step_count
        bc1 = 1.0 - self.b1 ** t
        bc2 = 1.0 - self.b2 ** t
        step_size = self.lr / bc1
        bc2s = math.sqrt(bc2)
        for i, (p, g) in enumerate(zip(params, grads)):
            self.m[i] = self.m[i] + (g - self.m[i]) * (1.0 - self.b1)  # lerp_
            self.v[i] = self.v[i] * self.b2 + (1.0 - self.b2) * g * g
            denom = self.v[i].sqrt() / bc2s + self.eps
            params[i] = p - step_size * (self.m[i] / denom)
        return params


class PPOLearnerOracle:
    """ppo_learner.py:92-238 `learn`, restated around ppo_minibatch/clip_grad_norm/AdamOracle."""

    def __init__(self, pol, val, batch_size, n_epochs, policy_lr, critic_lr, clip_range, ent_coef,
                 mini_batch_size, quant=None, policy_type=0, var_range=(0.1, 1.0)):
        self.policy_type, self.var_range = policy_type, var_range
        self.pol = [p.clone() for p in pol]
        self.val = [p.clone() for p in val]
        self.batch_size, self.n_epochs = batch_size, n_epochs
        self.mini_batch_size = mini_batch_size
        self.clip_range, self.ent_coef = clip_range, ent_coef
        self.popt = AdamOracle(self.pol, policy_lr)
        self.vopt = AdamOracle(self.val, critic_lr)
        self.cumulative_model_updates = 0
        self.quant = quant
        self.last_grads = None

    def learn(self, buf: BufferOracle):
        n_it = n_mb = 0
        m_ent = m_kl = m_vl = 0.0
        clips = []
        pb = torch.cat([p.flatten() for p in self.pol])
        vb = torch.cat([p.flatten() for p in self.val])
        for _ in range(self.n_epochs):  # :119
            for _, (acts, oldp, obs, tgt, adv) in buf.batches(self.batch_size):  # :121-129
                pg = [torch.zeros_like(p) for p in self.pol]
                vg = [torch.zeros_like(p) for p in self.val]
                for s in range(0, self.batch_size, self.mini_batch_size):  # :134
                    e = s + self.mini_batch_size
                    args = (self.pol, self.val, torch.from_numpy(obs[s:e]), torch.from_numpy(acts[s:e]),
                            torch.from_numpy(oldp[s:e]), torch.from_numpy(tgt[s:e]), torch.from_numpy(adv[s:e]),
                            self.clip_range, self.ent_coef, self.batch_size)
                    if self.policy_type == 0:
                        g1, g2, mt = ppo_minibatch(*args, self.quant)
                    else:  # ppo_learner.py:131 batch_acts.view(batch_size, -1): one row of actions per sample
                        g1, g2, mt = ppo_minibatch_head(self.policy_type, *args, self.var_range, self.quant)
                    pg = [a + b for a, b in zip(pg, g1)]
                    vg = [a + b for a, b in zip(vg, g2)]
                    m_vl += mt["value_loss"]
                    m_kl += mt["kl"]
                    m_ent += mt["entropy"]
                    clips.append(mt["clip_fraction"])
                    n_mb += 1
                vg, _ = clip_grad_norm(vg, 0.5)  # :187-189
                pg, _ = clip_grad_norm(pg, 0.5)  # :190
                self.last_grads = (pg, vg)
                self.pol = self.popt.step(self.pol, pg)  # :192
                self.val = self.vopt.step(self.val, vg)  # :193
                n_it += 1
        n_it = max(n_it, 1)
        n_mb = max(n_mb, 1)
        pa = torch.cat([p.flatten() for p in self.pol])
        va = torch.cat([p.flatten() for p in self.val])
        self.cumulative_model_updates += n_it
        return {
            "Cumulative Model Updates": self.cumulative_model_updates,
            "Policy Entropy": m_ent / n_mb,
            "Mean KL Divergence": m_kl / n_mb,
            "Value Function Loss": m_vl / n_mb,
            "SB3 Clip Fraction": float(np.mean(clips)) if clips else 0,
            "Policy Update Magnitude": float((pb - pa).norm()),
            "Value Function Update Magnitude": float((vb - va).norm()),
        }


def add_new_experience(val_params, buf: BufferOracle, stats: WelfordOracle, experience, gamma, lmbda,
                       standardize_returns=True, max_returns_per_stats_increment=150, quant=None):
    """learner.py:330-385."""
    states, actions, log_probs, rewards, next_states, dones, truncated = experience
    val_inp = np.zeros((states.shape[0] + 1, states.shape[1]))  # :347 (float64 staging)
    val_inp[:-1] = states
    val_inp[-1] = next_states[-1]
    v, _ = mlp_forward(val_params, torch.as_tensor(val_inp, dtype=torch.float32), quant)  # :352
    val_preds = v.flatten().numpy()
    ret_std = stats.get_std()[0] if standardize_returns else None  # :356
    vt, adv, rets = gae_nep50(rewards, dones, truncated, val_preds, gamma, lmbda, ret_std)  # :358
    if standardize_returns:  # :368-372
        n_inc = min(max_returns_per_stats_increment, len(rets))
        stats.increment(rets[:n_inc], n_inc)
    buf.submit(states=states, actions=actions, log_probs=log_probs, rewards=rewards, next_states=next_states,
               dones=dones, truncated=truncated, values=vt, advantages=adv)  # :375-385
    return val_preds, vt, adv, rets


def quant_bf16(t):
    return t.to(torch.bfloat16).to(torch.float32)
