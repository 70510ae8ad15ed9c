"""Device time of the fused training launch alone (C2 shape: 50k rows, obs 89, 90 actions, 256x3 nets): N back-to-back
launches replayed from one CUDA graph between two CUDA events, after warm-up; weights / inputs stay in L2 as in a training step.
For A/B comparisons of kernel variants: python tools/time_fused.py [launches] [repeats]"""
import contextlib, io, os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from rlgym_ppo_b200 import ops
from rlgym_ppo_b200.ppo import ExperienceBuffer, PPOLearner
N = int(sys.argv[1]) if len(sys.argv) > 1 else 50
R = int(sys.argv[2]) if len(sys.argv) > 2 else 5
M = int(os.environ.get("ROWS", "50000"))
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    lr = PPOLearner(89, 90, 0, (256,) * 3, (256,) * 3, (0.1, 1.0), M, 1, 3e-4, 3e-4, 0.2, 0.01, M, "cuda:0")
lr.use_cuda_graph = False
rng = np.random.RandomState(0)
b = ExperienceBuffer(M, 1, "cuda:0")
b.submit_experience(rng.randn(M, 89).astype(np.float32), rng.randint(0, 90, M).astype(np.float32), np.full(M, -4.5, np.float32),
                    np.zeros(M, np.float32), np.zeros((M, 89), np.float32), np.zeros(M, np.float32), np.zeros(M),
                    rng.randn(M).astype(np.float32), rng.randn(M).astype(np.float32))
lr.learn(b)
mb = lr._minibatch_buffers(M)
ps, vs = lr.policy._stack, lr.value_net._stack
wp, wv = ps.workspace(M), vs.workspace(M)
x = mb["x"]
metrics = lr._step_metrics()
def launch():
    ops.policy_value_train_fused(ps.fused_net(x.stride(0), wp, policy_head=True), vs.fused_net(x.stride(0), wv), x, M,
                                 lr.policy.n_actions, mb["actions"], mb["old_logp"], mb["adv"], 1.0 / M, 0.2, 0.01, vs.w[-1],
                                 mb["targets"], vs.gw[-1], metrics)
for _ in range(10):
    launch()
torch.cuda.synchronize()
# the launches replay from a CUDA graph (as in a training step): eagerly, the host-side launch preparation (tensor-map
# encoding, ctypes) takes about as long as the kernel and the loop would measure the host
g = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    with torch.cuda.graph(g, stream=side):
        for _ in range(N):
            launch()
torch.cuda.synchronize()
g.replay()
torch.cuda.synchronize()
out = []
for _ in range(R):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    out.append(e0.elapsed_time(e1) * 1e3 / N)
print("fused train launch us (graph of %d):" % N, " ".join(f"{t:.1f}" for t in out), " min %.1f" % min(out))
