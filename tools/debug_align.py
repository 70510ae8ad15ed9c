"""Diagnostic: does the value net's gradient at the SECOND optimiser step depend on the alignment of its arena offset?"""
import sys, os, contextlib, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import ref_oracle as O
from parity_helpers import capture_steps, rel_l2, unflatten
from rlgym_ppo_b200.ppo import ExperienceBuffer, PPOLearner

DEV = "cuda:0"
NAMES = ("states", "actions", "log_probs", "rewards", "next_states", "dones", "truncated", "values", "advantages")
for n_act in (21, 24, 90):
    for precision in ("fp32",):
        torch.manual_seed(0)
        rng = np.random.RandomState(1)
        obs, B, total = 89, 256, 600
        with contextlib.redirect_stdout(io.StringIO()):
            lr = PPOLearner(obs, n_act, 0, (64, 64), (64, 64), (0.1, 1.0), B, 1, 3e-4, 3e-4, 0.2, 0.01, 128, DEV, precision=precision)
        pol = [p.detach().cpu().clone() for p in lr.policy.parameters()]
        val = [p.detach().cpu().clone() for p in lr.value_net.parameters()]
        f = dict(states=rng.randn(total, obs).astype(np.float32), actions=rng.randint(0, n_act, total).astype(np.float32),
                 log_probs=(-np.abs(rng.randn(total)) - 2).astype(np.float32), rewards=np.zeros(total, np.float32),
                 next_states=rng.randn(total, obs).astype(np.float32), dones=np.zeros(total, np.float32),
                 truncated=np.zeros(total), values=rng.randn(total).astype(np.float32),
                 advantages=rng.randn(total).astype(np.float32))
        buf = ExperienceBuffer(1000, 123, DEV)
        buf.submit_experience(*[f[k] for k in NAMES])
        cap = capture_steps(lr)
        lr.learn(buf)
        ob = O.BufferOracle(1000, 123)
        ob.submit(**f)
        orc = O.PPOLearnerOracle([p.double() for p in pol], [p.double() for p in val], B, 1, 3e-4, 3e-4, 0.2, 0.01, 128)
        seen = []
        orig = orc.popt.step
        def spy(p, g, orig=orig, orc=orc, seen=seen):
            seen.append([t.numpy() for t in orc.last_grads[0] + orc.last_grads[1]])
            return orig(p, g)
        orc.popt.step = spy
        _pm = O.ppo_minibatch
        O.ppo_minibatch = lambda pol_, val_, o, a, ol, tg, ad, *r, **k: _pm(pol_, val_, o.double(), a, ol.double(), tg.double(), ad.double(), *r, **k)
        orc.learn(ob)
        O.ppo_minibatch = _pm
        print(f"n_act={n_act} n_p={int(lr._seg[1])} (mod 4 = {int(lr._seg[1]) % 4}) {precision}")
        for s in range(len(cap)):
            print("   step", s, " ".join(f"{rel_l2(a, b):.1e}" for a, b in zip(unflatten(lr, cap[s]), seen[s])))
        w = [p.detach().cpu().numpy() for p in list(lr.policy.parameters()) + list(lr.value_net.parameters())]
        print("   weights", " ".join(f"{rel_l2(a, b.numpy()):.1e}" for a, b in zip(w, orc.pol + orc.val)))
        # operands vs master weights
        from rlgym_ppo_b200 import ops
        for name, st in (("pol", lr.policy._stack), ("val", lr.value_net._stack)):
            errs = []
            for i, l in enumerate(st.linears):
                ps = ops.pad64(l.in_features)
                wq = sum(st.wq[i][:l.out_features, q * ps:q * ps + l.in_features].float() for q in range(3))
                errs.append(float((wq - st.w[i]).abs().max()))
            print("   ", name, "operand - master max abs:", errs)
