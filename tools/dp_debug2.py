"""Where does a replicated data-parallel run separate from a one-rank run?  (torchrun, 2 ranks)"""
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
B = 2048
def learner(group, epochs):
    torch.manual_seed(5)
    with contextlib.redirect_stdout(io.StringIO()):
        return PPOLearner(89, 90, 0, (256, 256), (256, 256), (0.1, 1.0), B, epochs, 3e-4, 3e-4, 0.2, 0.01, B, dev, process_group=group, dp_mode="replicated")
def rl(a, b): return float((a - b).norm() / b.norm())
from rlgym_ppo_b200.ppo import ExperienceBuffer
_orig_into = ExperienceBuffer.next_permutation_into
def _main_stream_into(self, dst):
    perm = self.next_permutation()
    dst.copy_(perm, non_blocking=True)
    ev = torch.cuda.Event(); ev.record()
    self._pin_ev[self._pin_of_last] = ev
for name, sync_after, main_stream, order in (("baseline", False, False, "dp_first"), ("sync after make_buffer", True, False, "dp_first"),
                                      ("perm upload on main stream", False, True, "dp_first"), ("alone first", False, False, "alone_first"),
                                      ("two alone learners", False, False, "alone_alone")):
    ExperienceBuffer.next_permutation_into = _main_stream_into if main_stream else _orig_into
    dp = learner(solo if order == "alone_alone" else None, 2)
    alone = learner(solo, 2)
    dp.use_cuda_graph = alone.use_cuda_graph = False
    p0 = dp._params.clone()
    for it in range(3):
        b_dp, b_al = make_buffer(100 + it, 3 * B, dev), make_buffer(100 + it, 3 * B, dev)
        if sync_after:
            torch.cuda.synchronize()
        if order == "alone_first":
            r2 = alone.learn(b_al); r1 = dp.learn(b_dp)
        else:
            r1 = dp.learn(b_dp); r2 = alone.learn(b_al)
        torch.cuda.synchronize()
        if rank == 0:
            print(f"{name:28s} call {it}: update rel-L2 {rl(dp._params - p0, alone._params - p0):.2e}  m {rl(dp._m, alone._m):.2e}", flush=True)
dist.destroy_process_group()
