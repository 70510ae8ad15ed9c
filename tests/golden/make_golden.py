"""
Generates the golden fixtures in this directory by running the UNMODIFIED reference
(/root/reference, AechPro/rlgym-ppo v1.3.13) in the authoring container.

    python tests/golden/make_golden.py

/root/reference does not exist on the GPU box, so nothing under tests/ or bench.py imports it at run
time; only this script does.  The reference needs `gym` at import time (rlgym_ppo/util/__init__.py:3 ->
rlgym_v2_gym_wrapper.py:1); a 3-line stub module stands in for it (no reference code is modified).

Fixtures (all float data stored exactly as the reference produced it):
  gae.npz        compute_gae                     (rlgym_ppo/util/torch_functions.py:36-78)
  welford.npz    WelfordRunningStat              (rlgym_ppo/util/running_stats.py:15-98)
  buffer.npz     ExperienceBuffer FIFO + shuffle (rlgym_ppo/ppo/experience_buffer.py:17-102)
  policy.npz     DiscreteFF / ValueEstimator     (discrete_policy.py:35-80, value_estimator.py:30-36)
  ppo_learn.npz  PPOLearner.learn                (rlgym_ppo/ppo/ppo_learner.py:92-238)
  add_exp.npz    Learner.add_new_experience      (rlgym_ppo/learner.py:330-385)
"""
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def import_reference():
    gym = types.ModuleType("gym")
    gym.Env = object
    gym.spaces = types.ModuleType("gym.spaces")
    sys.modules.setdefault("gym", gym)
    sys.modules.setdefault("gym.spaces", gym.spaces)
    sys.path.insert(0, REF)
    import rlgym_ppo  # noqa: F401
    from rlgym_ppo.learner import Learner
    from rlgym_ppo.ppo import DiscreteFF, ExperienceBuffer, PPOLearner, ValueEstimator
    from rlgym_ppo.util import WelfordRunningStat, torch_functions
    assert rlgym_ppo.__file__.startswith(REF)
    return SimpleNamespace(Learner=Learner, DiscreteFF=DiscreteFF, ExperienceBuffer=ExperienceBuffer,
                           PPOLearner=PPOLearner, ValueEstimator=ValueEstimator,
                           WelfordRunningStat=WelfordRunningStat, compute_gae=torch_functions.compute_gae)


def synth_rollout(rng, n, obs_dim, p_done=1 / 300, p_trunc=1 / 1500, rew_scale=0.1):
    """SURVEY.md 8(d) synthetic rollout in the layout of batched_agent_manager.py:159-168."""
    states = rng.randn(n, obs_dim).astype(np.float32)
    next_states = np.roll(states, -1, axis=0).copy()
    rewards = (rng.randn(n) * rew_scale).astype(np.float32)
    dones = (rng.rand(n) < p_done).astype(np.float32)
    truncated = ((rng.rand(n) < p_trunc) * (1 - dones)).astype(np.float64)
    truncated[-1] = 1.0 - dones[-1]  # batched_agent_manager.py:145
    return states, rewards, next_states, dones, truncated


def sd_to_list(module):
    return [p.detach().clone().numpy() for p in module.parameters()]


def gen_gae(ref):
    out = {}
    rng = np.random.RandomState(0)
    cases = []
    for name, n, pd, pt, std, scale in [
        ("basic", 1000, 1 / 300, 1 / 1500, np.float32(1.0), 0.1),
        ("std", 777, 1 / 50, 1 / 100, np.float32(0.37), 0.1),
        ("clip", 513, 1 / 20, 1 / 40, np.float32(0.01), 0.5),   # r/std hits the +-10 clip
        ("nostd", 300, 1 / 10, 1 / 10, None, 1.0),
        ("alldone", 64, 1.0, 0.0, np.float32(1.0), 0.1),
        ("nodone", 257, 0.0, 0.0, np.float32(2.5), 0.1),
        ("one", 1, 0.0, 0.0, np.float32(1.0), 0.1),
        ("done_edges", 130, 0.0, 0.0, np.float32(1.0), 0.1),
    ]:
        _, rewards, _, dones, truncated = synth_rollout(rng, n, 4, pd, pt, scale)
        if name == "done_edges":
            dones[0] = 1.0
            dones[-1] = 1.0
            truncated[-1] = 0.0
            truncated[64] = 1.0
        values = rng.randn(n + 1).astype(np.float32)
        vt, adv, rets = ref.compute_gae(rewards, dones, truncated, values.tolist(), gamma=0.99, lmbda=0.95,
                                        return_std=std)
        out[f"{name}.rew"] = rewards
        out[f"{name}.done"] = dones
        out[f"{name}.trunc"] = truncated
        out[f"{name}.val"] = values
        out[f"{name}.std"] = np.asarray([np.nan if std is None else std], np.float32)
        out[f"{name}.vt"] = vt.numpy()
        out[f"{name}.adv"] = adv.numpy()
        out[f"{name}.ret"] = np.asarray(rets, np.float64)
        cases.append(name)
    out["cases"] = np.asarray(cases)
    out["gamma_lambda"] = np.asarray([0.99, 0.95])
    np.savez_compressed(os.path.join(HERE, "gae.npz"), **out)


def gen_welford(ref):
    rng = np.random.RandomState(1)
    out = {}
    st = ref.WelfordRunningStat(1)
    out["std_empty"] = st.std.copy()
    out["mean_empty"] = st.mean.copy()
    samples = (rng.randn(400) * 3 + 1).astype(np.float64)  # f64 like compute_gae's returns list
    st.increment(list(samples[:150]), 150)
    out["s150.mean"], out["s150.m2"], out["s150.count"] = st.running_mean.copy(), st.running_variance.copy(), np.asarray([st.count])
    out["s150.std"] = np.asarray(st.std).copy()
    st.increment(list(samples[150:151]), 1)  # num == 1 path: update(samples) with a 1-list
    out["s151.mean"], out["s151.m2"] = st.running_mean.copy(), st.running_variance.copy()
    st.increment(list(samples[151:400]), 249)
    out["s400.mean"], out["s400.m2"], out["s400.count"] = st.running_mean.copy(), st.running_variance.copy(), np.asarray([st.count])
    out["s400.std"] = np.asarray(st.std).copy()
    out["samples"] = samples
    # zero-variance guard (running_stats.py:67-68)
    z = ref.WelfordRunningStat(1)
    z.increment([np.float64(2.0)] * 5, 5)
    out["const.std"] = np.asarray(z.std).copy()
    out["const.mean"] = np.asarray(z.mean).copy()
    # vector stats with f32 samples (obs stats, batched_agent_manager.py:377-380) + merge (:71-98)
    a = ref.WelfordRunningStat(5)
    b = ref.WelfordRunningStat(5)
    xs = rng.randn(60, 5).astype(np.float32)
    a.increment(xs[:40], 40)
    b.increment(xs[40:], 20)
    out["vec.samples"] = xs
    out["vec.a.mean"], out["vec.a.m2"] = a.running_mean.copy(), a.running_variance.copy()
    ser = b.serialize()
    out["vec.b.ser"] = np.asarray(ser, np.float64)
    a.increment_from_serialized_other(ser)
    out["vec.merged.mean"], out["vec.merged.m2"] = np.asarray(a.running_mean).copy(), np.asarray(a.running_variance).copy()
    out["vec.merged.count"] = np.asarray([a.count])
    out["vec.merged.std"] = np.asarray(a.std).copy()
    np.savez_compressed(os.path.join(HERE, "welford.npz"), **out)


def gen_buffer(ref):
    rng = np.random.RandomState(2)
    out = {}
    obs_dim, max_size, seed = 7, 100, 123
    buf = ref.ExperienceBuffer(max_size, seed, "cpu")
    sizes = [30, 50, 40, 100, 130, 5]  # grows, wraps, == size, > size, small
    out["sizes"] = np.asarray(sizes)
    out["cfg"] = np.asarray([obs_dim, max_size, seed])
    for k, n in enumerate(sizes):
        f = {
            "states": rng.randn(n, obs_dim).astype(np.float32),
            "actions": rng.randint(0, 90, n).astype(np.float32),
            "log_probs": rng.randn(n).astype(np.float32),
            "rewards": rng.randn(n).astype(np.float32),
            "next_states": rng.randn(n, obs_dim).astype(np.float32),
            "dones": (rng.rand(n) < 0.1).astype(np.float32),
            "truncated": (rng.rand(n) < 0.1).astype(np.float64),
            "values": rng.randn(n).astype(np.float32),
            "advantages": rng.randn(n).astype(np.float32),
        }
        for name, v in f.items():
            out[f"in{k}.{name}"] = v
        buf.submit_experience(f["states"], f["actions"], f["log_probs"], f["rewards"], f["next_states"],
                              f["dones"], f["truncated"], f["values"], f["advantages"])
        for name in f:
            out[f"after{k}.{name}"] = getattr(buf, name).numpy().copy()
        # two epochs of shuffled batches after every submit (rng state persists across calls)
        for ep in range(2):
            bs = 32
            batches = list(buf.get_all_batches_shuffled(bs))
            out[f"after{k}.ep{ep}.nbatches"] = np.asarray([len(batches)])
            for bi, (acts, lp, st, vals, adv) in enumerate(batches):
                out[f"after{k}.ep{ep}.b{bi}.actions"] = acts.numpy().copy()
                out[f"after{k}.ep{ep}.b{bi}.log_probs"] = lp.numpy().copy()
                out[f"after{k}.ep{ep}.b{bi}.states"] = st.numpy().copy()
                out[f"after{k}.ep{ep}.b{bi}.values"] = vals.numpy().copy()
                out[f"after{k}.ep{ep}.b{bi}.advantages"] = adv.numpy().copy()
    # plain permutation stream used by the bench shape
    r = np.random.RandomState(123)
    out["perm150000.head"] = r.permutation(150000)[:64]
    out["perm150000.second_head"] = r.permutation(150000)[:64]
    r = np.random.RandomState(7)
    out["perm1000"] = r.permutation(1000)
    out["perm1000.b"] = r.permutation(1000)
    np.savez_compressed(os.path.join(HERE, "buffer.npz"), **out)


def gen_policy(ref):
    torch.manual_seed(11)
    out = {}
    obs_dim, n_act, layers = 89, 90, (64, 64)
    pol = ref.DiscreteFF(obs_dim, n_act, layers, "cpu")
    val = ref.ValueEstimator(obs_dim, layers, "cpu")
    rng = np.random.RandomState(3)
    obs = rng.randn(200, obs_dim).astype(np.float32)
    with torch.no_grad():
        probs = pol.get_output(obs)
        acts, logp = pol.get_action(obs)
        v = val(obs.astype(np.float64))  # float64 input path, learner.py:347-352
        lp2, ent = pol.get_backprop_data(torch.from_numpy(obs), acts.view(-1, 1).float())
    out["cfg"] = np.asarray([obs_dim, n_act, *layers])
    for i, p in enumerate(sd_to_list(pol)):
        out[f"pol.{i}"] = p
    for i, p in enumerate(sd_to_list(val)):
        out[f"val.{i}"] = p
    out["pol.keys"] = np.asarray(list(pol.state_dict().keys()))
    out["val.keys"] = np.asarray(list(val.state_dict().keys()))
    out["obs"], out["probs"], out["actions"], out["logp"] = obs, probs.numpy(), acts.numpy(), logp.numpy()
    out["values"] = v.numpy()
    out["bp_logp"], out["bp_entropy"] = lp2.numpy(), np.asarray([ent.item()])
    np.savez_compressed(os.path.join(HERE, "policy.npz"), **out)


def capture_grads(opt, store):
    orig = opt.step

    def step(*a, **k):
        store.append([p.grad.detach().clone().numpy() for g in opt.param_groups for p in g["params"]])
        return orig(*a, **k)

    opt.step = step


def gen_ppo(ref):
    """PPOLearner.learn on a small buffer: 2 epochs x 2 batches x 2 minibatches, log-probs perturbed so
    the clip is active (clip fraction ~0.5), advantages un-normalised."""
    torch.manual_seed(123)
    out = {}
    obs_dim, n_act, layers = 89, 90, (64, 64)
    B, mb, epochs, total = 256, 128, 2, 600
    learner = ref.PPOLearner(obs_dim, n_act, 0, layers, layers, (0.1, 1.0), B, epochs, 3e-4, 3e-4, 0.2, 0.01,
                             mb, "cpu")
    out["cfg"] = np.asarray([obs_dim, n_act, B, mb, epochs, total, *layers])
    out["hyper"] = np.asarray([3e-4, 3e-4, 0.2, 0.01])
    for i, p in enumerate(sd_to_list(learner.policy)):
        out[f"pol0.{i}"] = p
    for i, p in enumerate(sd_to_list(learner.value_net)):
        out[f"val0.{i}"] = p
    rng = np.random.RandomState(4)
    states, rewards, next_states, dones, truncated = synth_rollout(rng, total, obs_dim, 1 / 50, 1 / 100)
    with torch.no_grad():
        acts, logp = learner.policy.get_action(states)
    logp = logp.numpy() + (rng.randn(total) * 0.3).astype(np.float32)
    values = rng.randn(total).astype(np.float32)
    adv = (rng.randn(total) * 0.5).astype(np.float32)
    buf = ref.ExperienceBuffer(1000, 123, "cpu")
    buf.submit_experience(states, acts.numpy().astype(np.float32), logp, rewards, next_states, dones, truncated,
                          values, adv)
    for name in ("states", "actions", "log_probs", "rewards", "next_states", "dones", "truncated", "values",
                 "advantages"):
        out[f"buf.{name}"] = getattr(buf, name).numpy().copy()
    pg, vg = [], []
    capture_grads(learner.policy_optimizer, pg)
    capture_grads(learner.value_optimizer, vg)
    report = learner.learn(buf)
    out["report.keys"] = np.asarray([k for k in report if k != "PPO Batch Consumption Time"])
    out["report.vals"] = np.asarray([float(report[k]) for k in report if k != "PPO Batch Consumption Time"])
    out["n_steps"] = np.asarray([len(pg)])
    for s in range(len(pg)):
        for i, g in enumerate(pg[s]):
            out[f"pgrad{s}.{i}"] = g  # post-clip grads seen by Adam at step s
        for i, g in enumerate(vg[s]):
            out[f"vgrad{s}.{i}"] = g
    for i, p in enumerate(sd_to_list(learner.policy)):
        out[f"pol1.{i}"] = p
    for i, p in enumerate(sd_to_list(learner.value_net)):
        out[f"val1.{i}"] = p
    psd = learner.policy_optimizer.state_dict()["state"]
    for i in psd:
        out[f"padam.{i}.m"] = psd[i]["exp_avg"].numpy()
        out[f"padam.{i}.v"] = psd[i]["exp_avg_sq"].numpy()
    out["padam.step"] = np.asarray([float(psd[0]["step"])])
    np.savez_compressed(os.path.join(HERE, "ppo_learn.npz"), **out)


def gen_add_exp(ref):
    """Unbound Learner.add_new_experience on a SimpleNamespace, two iterations so
    the second one runs with a non-trivial return_std."""
    torch.manual_seed(5)
    out = {}
    obs_dim, layers, n = 89, (64, 64), 400
    val = ref.ValueEstimator(obs_dim, layers, "cpu")
    for i, p in enumerate(sd_to_list(val)):
        out[f"val.{i}"] = p
    ns = SimpleNamespace(ppo_learner=SimpleNamespace(value_net=val), return_stats=ref.WelfordRunningStat(1),
                         standardize_returns=True, gae_gamma=0.99, gae_lambda=0.95,
                         max_returns_per_stats_increment=150,
                         experience_buffer=ref.ExperienceBuffer(600, 123, "cpu"))
    rng = np.random.RandomState(6)
    out["cfg"] = np.asarray([obs_dim, n, 600, *layers])
    for it in range(2):
        states, rewards, next_states, dones, truncated = synth_rollout(rng, n, obs_dim, 1 / 40, 1 / 80, 0.3)
        actions = rng.randint(0, 90, n).astype(np.float32)
        log_probs = (-np.abs(rng.randn(n)) - 3).astype(np.float32)
        exp = (states, actions, log_probs, rewards, next_states, dones, truncated)
        for name, a in zip(("states", "actions", "log_probs", "rewards", "next_states", "dones", "truncated"), exp):
            out[f"it{it}.{name}"] = a
        out[f"it{it}.std_before"] = np.asarray(ns.return_stats.std).copy()
        ref.Learner.add_new_experience(ns, exp)
        out[f"it{it}.stats"] = np.asarray([ns.return_stats.running_mean[0], ns.return_stats.running_variance[0],
                                           ns.return_stats.count], np.float64)
        out[f"it{it}.buf.values"] = ns.experience_buffer.values.numpy().copy()
        out[f"it{it}.buf.advantages"] = ns.experience_buffer.advantages.numpy().copy()
        out[f"it{it}.buf.len"] = np.asarray([ns.experience_buffer.rewards.shape[0]])
    np.savez_compressed(os.path.join(HERE, "add_exp.npz"), **out)


if __name__ == "__main__":
    ref = import_reference()
    print("reference imported from", REF, "| numpy", np.__version__, "| torch", torch.__version__)
    gen_gae(ref)
    gen_welford(ref)
    gen_buffer(ref)
    gen_policy(ref)
    gen_ppo(ref)
    gen_add_exp(ref)
    with open(os.path.join(HERE, "VERSIONS.txt"), "w") as f:
        f.write(f"reference: AechPro/rlgym-ppo v1.3.13 (/root/reference)\nnumpy {np.__version__}\ntorch {torch.__version__}\n")
    for fn in sorted(os.listdir(HERE)):
        print(f"{fn:20s} {os.path.getsize(os.path.join(HERE, fn)):>9d} B")
