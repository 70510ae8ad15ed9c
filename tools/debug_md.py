import sys, os, contextlib, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import ref_oracle as O
from parity_helpers import capture_steps, rel_l2, unflatten
from rlgym_ppo_b200.ppo import ExperienceBuffer, PPOLearner
DEV = "cuda:0"
NAMES = ("states", "actions", "log_probs", "rewards", "next_states", "dones", "truncated", "values", "advantages")
g = np.load(os.path.join(ROOT, "tests/golden/heads.npz"))
for tag, ptype in (("md", 1), ("ct", 2)):
    obs_dim, B, mb, epochs, total, n_cont, l0, l1 = [int(x) for x in g["cfg"]]
    plr, clr, clip, ent, vmin, vmax = [float(x) for x in g["hyper"]]
    with contextlib.redirect_stdout(io.StringIO()):
        lr = PPOLearner(obs_dim, n_cont, ptype, (l0, l1), (l0, l1), (vmin, vmax), B, epochs, plr, clr, clip, ent, mb, DEV, precision="fp32")
    for net, pre in ((lr.policy, f"{tag}.pol0"), (lr.value_net, f"{tag}.val0")):
        keys = list(net.state_dict().keys())
        net.load_state_dict({k: torch.from_numpy(g[f"{pre}.{i}"]) for i, k in enumerate(keys)})
    buf = ExperienceBuffer(1000, 123, DEV)
    buf.submit_experience(*[g[f"{tag}.buf.{n}"] for n in NAMES])
    cap = capture_steps(lr)
    wsnap = []
    orig = lr._optimizer_step
    def hook(orig=orig, lr=lr, wsnap=wsnap):
        orig()
        wsnap.append(lr._params.detach().cpu().numpy().copy())
        wsnap.append(lr._m.detach().cpu().numpy().copy())
    lr._optimizer_step = hook
    lr.learn(buf)
    # oracle (fp32, reproduces the golden to 1e-7)
    ob = O.BufferOracle(1000, 123)
    ob.submit(**{n: g[f"{tag}.buf.{n}"] for n in O.FIELDS})
    pol0 = [torch.from_numpy(g[f"{tag}.pol0.{i}"]) for i in range(6)]
    val0 = [torch.from_numpy(g[f"{tag}.val0.{i}"]) for i in range(6)]
    orc = O.PPOLearnerOracle(pol0, val0, B, epochs, plr, clr, clip, ent, mb, policy_type=ptype, var_range=(vmin, vmax))
    snaps = []
    o2 = orc.vopt.step
    def spy(p, gr, o2=o2, orc=orc, snaps=snaps):
        out = o2(p, gr)
        snaps.append((torch.cat([t.flatten() for t in orc.pol]).numpy().copy(), torch.cat([t.flatten() for t in out]).numpy().copy(),
                      torch.cat([t.flatten() for t in orc.vopt.m]).numpy().copy(), [t.numpy().copy() for t in orc.last_grads[1]]))
        return out
    orc.vopt.step = spy
    orc.learn(ob)
    n_p = int(lr._seg[1])
    for s in range(2):
        ours_v = wsnap[2 * s][n_p:]
        print(tag, "after step", s, "value weights rel", rel_l2(ours_v, snaps[s][1]), "max abs", float(np.abs(ours_v - snaps[s][1]).max()),
              "| m rel", rel_l2(wsnap[2 * s + 1][n_p:], snaps[s][2]))
        vg = unflatten(lr, cap[s])[6:]
        print("     value grads vs oracle:", " ".join(f"{rel_l2(a, b):.1e}" for a, b in zip(vg, snaps[s][3])))
    print(tag, "steps", lr._steps.tolist(), "sqnorm", lr._sqnorm.tolist())
