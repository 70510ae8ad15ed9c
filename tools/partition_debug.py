"""Debug: gradients of one 2048-row batch in one chunk vs two / sixteen chunks, per parameter tensor."""
import contextlib, io, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.dp_check import make_buffer
from rlgym_ppo_b200.ppo import PPOLearner
DEV = "cuda:0"
B, n = 2048, 3 * 2048
layers = tuple(int(x) for x in os.environ.get("LAYERS", "256,256").split(","))
torch.manual_seed(5)
with contextlib.redirect_stdout(io.StringIO()):
    lr = PPOLearner(89, 90, 0, layers, layers, (0.1, 1.0), B, 1, 3e-4, 3e-4, 0.2, 0.01, B, DEV)
lr.use_cuda_graph = False
lr.policy._stack.refresh_operands(); lr.value_net._stack.refresh_operands(); lr._sync_lr()
buf = make_buffer(100, n, DEV)
perm = buf.next_permutation_device()
idx = perm[:B].contiguous()
grads = {}
for chunk in (2048, 1024, 128):
    for rep in range(2):
        lr._tail.zero_()
        lr._backward_body(buf, idx, B, chunk)
        torch.cuda.synchronize()
        grads[(chunk, rep)] = (lr._grads.clone(), lr._tail.clone())
names = []
off = 0
for net, tag in ((lr.policy, "pol"), (lr.value_net, "val")):
    for k, p in net.named_parameters():
        names.append((tag + "." + k, off, p.numel())); off += p.numel()
ref = grads[(2048, 0)][0]
for key in ((2048, 1), (1024, 0), (1024, 1), (128, 0)):
    g = grads[key][0]
    print("chunk", key, "max abs diff", float((g - ref).abs().max()), "metrics diff", float((grads[key][1] - grads[(2048, 0)][1]).abs().max()))
    for name, o, cnt in names:
        a, b = g[o:o + cnt], ref[o:o + cnt]
        print(f"   {name:28s} rel_l2 {float((a - b).norm() / (b.norm() + 1e-30)):.3e}  max_abs {float((a - b).abs().max()):.3e}  ref_norm {float(b.norm()):.3e}")
