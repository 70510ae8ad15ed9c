"""Debug driver for the duo kernel: one backward pass of a small batch, eager, errors printed."""
import contextlib, io, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.dp_check import make_buffer
from rlgym_ppo_b200.ppo import PPOLearner
DEV = "cuda:0"
B = int(os.environ.get("ROWS", "512"))
torch.manual_seed(5)
with contextlib.redirect_stdout(io.StringIO()):
    lr = PPOLearner(89, 90, 0, (256, 256), (256, 256), (0.1, 1.0), B, 1, 3e-4, 3e-4, 0.2, 0.01, B, DEV)
lr.use_cuda_graph = False
lr.policy._stack.refresh_operands(); lr.value_net._stack.refresh_operands(); lr._sync_lr()
buf = make_buffer(100, 3 * B, DEV)
idx = buf.next_permutation_device()[:B].contiguous()
try:
    lr._backward_body(buf, idx, B, B)
    torch.cuda.synchronize()
    print("ok grads norm", float(lr._grads.norm()), "metrics", lr._tail[:8].tolist())
except Exception as e:
    print("FAILED:", e)
