"""GPU parity of the MultiDiscrete (21-logit Rolv) and Continuous (tanh-Gaussian) heads -- SURVEY.md 8(f)-4 -- against
the fixtures the unmodified reference produced (tests/golden/make_golden_heads.py -> heads.npz: get_backprop_data on a
minibatch, one PPOLearner.learn with the post-clip gradients of every optimiser step, updated weights, report) and
against the oracle's restated heads on injected random numbers.  multi_discrete_policy.py:16-89,
continuous_policy.py:23-120, torch_functions.py:15-33, 81-122, ppo_learner.py:34-50."""
import numpy as np
import pytest
import torch

from oracle import ref_oracle as O
from parity_helpers import capture_steps, close, rel_l2, unflatten

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
NAMES = ("states", "actions", "log_probs", "rewards", "next_states", "dones", "truncated", "values", "advantages")


def _learner(g, tag, ptype, precision):
    import contextlib
    import io
    from rlgym_ppo_b200.ppo import PPOLearner
    obs_dim, B, mb, epochs, total, n_cont, l0, l1 = [int(x) for x in g["cfg"]]
    plr, clr, clip, ent, vmin, vmax = [float(x) for x in g["hyper"]]
    with contextlib.redirect_stdout(io.StringIO()):
        lr = PPOLearner(obs_dim, n_cont, ptype, (l0, l1), (l0, l1), (vmin, vmax), B, epochs, plr, clr, clip, ent, mb, DEV,
                        precision=precision)
    for net, pre in ((lr.policy, f"{tag}.pol0"), (lr.value_net, f"{tag}.val0")):
        keys = list(net.state_dict().keys())
        net.load_state_dict({k: torch.from_numpy(g[f"{pre}.{i}"]) for i, k in enumerate(keys)})
    return lr, mb


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("tag,ptype", [("md", 1), ("ct", 2)])
def test_head_learn_matches_reference(golden, tag, ptype, precision):
    from rlgym_ppo_b200.ppo import ContinuousPolicy, ExperienceBuffer, MultiDiscreteFF
    g = golden("heads")
    lr, mb = _learner(g, tag, ptype, precision)
    assert isinstance(lr.policy, MultiDiscreteFF if ptype == 1 else ContinuousPolicy)
    exact = precision == "fp32"
    # state-dict layout = the reference's (continuous: the Sequential ends in Tanh, no extra keys)
    assert [tuple(v.shape) for v in lr.policy.state_dict().values()][-1] == ((21,) if ptype == 1 else (16,))
    # get_backprop_data on the first minibatch of the rollout
    obs = torch.from_numpy(g[f"{tag}.buf.states"][:mb])
    acts = torch.from_numpy(g[f"{tag}.buf.actions"][:mb])
    logp, ent = lr.policy.get_backprop_data(obs, acts)
    tol = 1e-4 if exact else 1e-2
    assert close(logp.cpu().numpy().reshape(-1), g[f"{tag}.bp_logp"].reshape(-1), tol)
    assert abs(float(ent) - float(g[f"{tag}.bp_entropy"][0])) < tol
    # PPOLearner.learn
    buf = ExperienceBuffer(1000, 123, DEV)
    buf.submit_experience(*[g[f"{tag}.buf.{n}"] for n in NAMES])
    assert np.array_equal(buf.actions.cpu().numpy(), g[f"{tag}.buf.actions"])           # action ROWS, byte-exact
    captured = capture_steps(lr)
    report = lr.learn(buf)
    n_steps = int(g[f"{tag}.n_steps"][0])
    assert len(captured) == n_steps
    ref_rep = dict(zip([str(k) for k in g[f"{tag}.report.keys"]], g[f"{tag}.report.vals"]))
    mtol = 1e-3 if exact else 3e-3
    for k in ("Policy Entropy", "Mean KL Divergence", "Value Function Loss", "SB3 Clip Fraction"):
        assert abs(report[k] - ref_rep[k]) <= mtol * max(1.0, abs(ref_rep[k])), (k, report[k], ref_rep[k])
    n_pol = len(list(lr.policy.parameters()))
    errs = []
    for s, flat in enumerate(captured):
        tensors = unflatten(lr, flat)
        wants = [g[f"{tag}.pgrad{s}.{i}"] if i < n_pol else g[f"{tag}.vgrad{s}.{i - n_pol}"] for i in range(len(tensors))]
        # a 1-element tensor (the value head's bias gradient, a sum that nearly cancels) has no meaningful relative error
        # of its own: it is measured against the norm of its network's whole gradient
        net_norm = [np.sqrt(sum(float((w.astype(np.float64) ** 2).sum()) for w in wants[:n_pol])),
                    np.sqrt(sum(float((w.astype(np.float64) ** 2).sum()) for w in wants[n_pol:]))]
        for i, (got, want) in enumerate(zip(tensors, wants)):
            if want.size == 1:
                errs.append((s, i, float(np.abs(got - want).max()) / net_norm[0 if i < n_pol else 1]))
            else:
                errs.append((s, i, rel_l2(got, want)))
    worst = max(e for _, _, e in errs)
    print(tag, precision, "max grad rel-L2 vs reference:", worst, [(s, i, float(f"{e:.1e}")) for s, i, e in errs if e > 1e-4])
    # Step 0 starts from the reference's weights: pure kernel error, held to the north-star 1e-3 in fp32 mode (measured
    # ~7e-6).  Later steps start from OUR step-0 weights: Adam normalises every gradient element to a +-lr move, so an
    # element whose gradient is ~0 (the value head's scalar bias gradient is a sum that nearly cancels) can move by +lr here
    # and -lr in the reference; that one weight shifts every value prediction and the NEXT gradient by a few per cent
    # (measured on B200: one element off by lr after step 0 -> 2-3.5e-2 on the md value net at step 1; the same happens
    # between two fp32 evaluations of the reference with different summation orders).
    assert max(e for s, _, e in errs if s == 0) < (1e-3 if exact else 0.25), errs
    assert worst < (6e-2 if exact else 0.25), errs
    for name, net in (("pol1", lr.policy), ("val1", lr.value_net)):
        for i, p in enumerate(net.parameters()):
            got, want = p.detach().cpu().numpy(), g[f"{tag}.{name}.{i}"]
            if exact:
                assert close(got, want, 1e-3), (name, i, np.abs(got - want).max())
                assert rel_l2(got, want) < 1e-3, (name, i)
            else:
                # bf16 operands: an element whose tiny gradient changes sign moves by up to 2 * lr per Adam step
                plr = float(g["hyper"][0])
                assert np.abs(got - want).max() <= 2 * plr * n_steps + 1e-6, (name, i, np.abs(got - want).max())


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_multi_discrete_sampling(golden, precision):
    """Inverse-CDF draws per distribution on injected uniforms equal the oracle's; log-probs are consistent with
    get_backprop_data; the deterministic branch returns the reference's [8, n] argmax layout."""
    from rlgym_ppo_b200 import ops
    g = golden("heads")
    lr, _ = _learner(g, "md", 1, precision)
    pol = lr.policy
    rng = np.random.RandomState(3)
    obs = g["md.buf.states"][:300]
    z = pol.get_output(obs).cpu()
    assert z.shape == (300, 21)
    zref, _ = O.mlp_forward([torch.from_numpy(g[f"md.pol0.{i}"]) for i in range(6)], torch.from_numpy(obs))
    assert close(z.numpy(), zref.numpy(), 1e-4 if precision == "fp32" else 2e-2)
    u = torch.from_numpy(rng.rand(300, 8).astype(np.float32)).to(DEV)
    zview, n, ws, _ = pol._logits(obs)
    acts = torch.empty((300, 8), device=DEV)
    logp = torch.empty(300, device=DEV)
    ops.head_multi_discrete_sample(zview[0], zview[1], zview[2], 300, acts, logp, u=u)
    # oracle draw from the DEVICE logits (so ties in the CDF do not depend on GEMM rounding)
    zd = z.double()
    start, want_a, want_lp = 0, [], torch.zeros(300, dtype=torch.float64)
    for gi, nb in enumerate(O.ROLV_BINS):
        lsm = zd[:, start:start + nb] - torch.logsumexp(zd[:, start:start + nb], -1, keepdim=True)
        c = torch.cumsum(lsm.exp().float(), -1)          # the kernel accumulates the CDF in f32
        a = (c > u[:, gi:gi + 1].cpu()).float().argmax(-1)
        a = torch.where((c > u[:, gi:gi + 1].cpu()).any(-1), a, torch.full_like(a, nb - 1))
        want_a.append(a)
        want_lp += lsm.gather(-1, a.view(-1, 1)).flatten()
        start += nb
    want_a = torch.stack(want_a, -1).float()
    mism = (acts.cpu() != want_a).float().mean()
    assert mism < 0.01, float(mism)         # CDF ties within one f32 ulp may land on the neighbouring bin
    same = (acts.cpu() == want_a).all(-1)
    assert close(logp.cpu().numpy()[same.numpy()], want_lp.numpy()[same.numpy()], 1e-4)
    lp2, _ = pol.get_backprop_data(obs, acts)
    assert close(lp2.cpu().numpy(), logp.cpu().numpy(), 1e-5)
    a_cpu, lp_cpu = pol.get_action(obs)
    assert a_cpu.shape == (300, 8) and a_cpu.dtype == torch.int64 and not a_cpu.is_cuda and lp_cpu.shape == (300,)
    assert int(a_cpu[:, :5].max()) <= 2 and int(a_cpu[:, 5:].max()) <= 1 and int(a_cpu.min()) >= 0
    det, zero = pol.get_action(obs, deterministic=True)
    assert det.shape == (8, 300) and zero == 0
    assert np.array_equal(det[0], z[:, 0:3].argmax(-1).numpy()) and np.array_equal(det[7], z[:, 19:21].argmax(-1).numpy())


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_continuous_sampling(golden, precision):
    from rlgym_ppo_b200 import ops
    g = golden("heads")
    lr, _ = _learner(g, "ct", 2, precision)
    pol = lr.policy
    obs = g["ct.buf.states"][:257]
    mean, std = pol.get_output(obs)
    assert mean.shape == (257, 8) and std.shape == (257, 8)
    assert float(std.min()) >= 0.1 - 1e-6 and float(std.max()) <= 1.0 + 1e-6 and float(mean.abs().max()) <= 1.0
    uref, _ = O.mlp_forward([torch.from_numpy(g[f"ct.pol0.{i}"]) for i in range(6)], torch.from_numpy(obs))
    assert close(mean.cpu().numpy(), torch.tanh(uref)[:, :8].numpy(), 1e-4 if precision == "fp32" else 2e-2)
    nrm = torch.randn(257, 8, generator=torch.Generator().manual_seed(1)).to(DEV)
    zview, n, ws, _ = pol._logits(obs)
    acts = torch.empty((257, 8), device=DEV)
    logp = torch.empty(257, device=DEV)
    ops.head_continuous_sample(zview[0], zview[1], zview[2], 257, 8, 0.1, 1.0, acts, logp, normals=nrm)
    want = (mean + std * nrm).clamp(-1, 1)
    assert close(acts.cpu().numpy(), want.cpu().numpy(), 1e-5)
    lp2, ent = pol.get_backprop_data(obs, acts)
    assert close(lp2.cpu().numpy(), logp.cpu().numpy(), 1e-4)
    want_ent = (0.5 + 0.5 * np.log(2 * np.pi) + torch.log(std)).mean()
    assert abs(float(ent) - float(want_ent)) < 1e-4
    a_cpu, lp_cpu = pol.get_action(obs)
    assert a_cpu.shape == (257, 8) and not a_cpu.is_cuda and float(a_cpu.abs().max()) <= 1.0
    # the Philox Box-Muller normals are standard normal: moments of (a - mean) / std over unclamped samples
    zs = ((a_cpu.to(DEV) - mean) / std)[(a_cpu.to(DEV).abs() < 1.0)]
    assert abs(float(zs.mean())) < 0.1
    det, zero = pol.get_action(obs, deterministic=True)
    assert zero == 0 and close(det.cpu().numpy(), mean.cpu().numpy(), 1e-6)
