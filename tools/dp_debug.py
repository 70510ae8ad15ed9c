import contextlib, io, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.dp_check import make_buffer
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]); lr_ = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr_); dev = f"cuda:{lr_}"
dist.init_process_group("nccl", device_id=torch.device(dev))
solo = None
for r in range(world):
    g = dist.new_group([r])
    if r == rank: solo = g
from rlgym_ppo_b200.ppo import PPOLearner
B, n = 2048, 3 * 2048
def learner(group, mode):
    torch.manual_seed(5)
    with contextlib.redirect_stdout(io.StringIO()):
        return PPOLearner(89, 90, 0, (256, 256), (256, 256), (0.1, 1.0), B, 1, 3e-4, 3e-4, 0.2, 0.01, B, dev, process_group=group, dp_mode=mode)
dp, alone = learner(None, "replicated"), learner(solo, "replicated")
print(rank, "init equal", torch.equal(dp._params, alone._params), dp.world_size, dp.rank, alone.world_size, alone.rank)
for L in (dp, alone):
    L.use_cuda_graph = False
    L.policy._stack.refresh_operands(); L.value_net._stack.refresh_operands(); L._sync_lr()
b1, b2 = make_buffer(100, n, dev), make_buffer(100, n, dev)
p1, p2 = b1.next_permutation_device(), b2.next_permutation_device()
print(rank, "perm equal", torch.equal(p1, p2))
local = B // world
dp._backward_body(b1, p1[rank * local:(rank + 1) * local], local, local)
g_local = dp._grads.clone()
dist.all_reduce(dp._grads)
alone._backward_body(b2, p2[:B], B, B)
torch.cuda.synchronize()
def rl(a, b): return float((a - b).norm() / b.norm())
print(rank, "grad rel err dp-vs-alone", rl(dp._grads, alone._grads), "local norm", float(g_local.norm()), "sum norm", float(dp._grads.norm()), "alone norm", float(alone._grads.norm()))
# also: alone on the two halves separately
alone._grads.zero_(); 
alone._train_chunk(b2, p2[:local], local); ga = alone._grads.clone(); alone._grads.zero_()
alone._train_chunk(b2, p2[local:B], local); gb = alone._grads.clone()
torch.cuda.synchronize()
print(rank, "alone halves sum vs full", rl(ga + gb, dp._grads), "my half vs alone half", rl(g_local, ga if rank == 0 else gb))
dist.destroy_process_group()
