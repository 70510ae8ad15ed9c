"""compute_gae on the B200 scan kernel (replaces rlgym_ppo/util/torch_functions.py:36-78).

Same signature and return types as the reference: (value_targets f32 tensor, advantages f32 tensor, returns
list).  The learner's own hot path (Learner.add_new_experience) calls ops.gae directly on device-resident arrays
and never materialises the Python list; this wrapper is the drop-in for code that calls compute_gae itself.
"""
import numpy as np
import torch

from .. import _lib, ops


def _dev(x, dtype=None):
    if isinstance(x, torch.Tensor):
        t = x.detach()
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(x)))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    elif t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float32)
    return t.to("cuda:%d" % torch.cuda.current_device()).contiguous()


def compute_gae_device(rews, dones, truncated, values, gamma=0.99, lmbda=0.95, return_std=1, n_head=0):
    """Device-resident variant: returns (value_targets, advantages, returns) f32 device tensors (+ the first
    n_head returns in f64 when n_head > 0).  return_std: None, a Python/NumPy scalar, or a 1-element f32 device
    tensor (e.g. WelfordRunningStat.device_std())."""
    _lib.require_device()
    r = _dev(rews, torch.float32)
    d = _dev(dones, torch.float32)
    t = _dev(truncated)
    v = _dev(values, torch.float32)
    if return_std is None:
        std = None
    elif isinstance(return_std, torch.Tensor) and return_std.is_cuda:
        std = return_std.to(torch.float32)
    else:
        std = torch.tensor([float(return_std)], dtype=torch.float32).to(r.device)
    head = torch.zeros(min(n_head, r.numel()), dtype=torch.float64, device=r.device) if n_head else None
    vt, adv, ret = ops.gae(r, d, t, v, gamma, lmbda, std, ret_head64=head)
    return (vt, adv, ret) if head is None else (vt, adv, ret, head)


def compute_gae(rews, dones, truncated, values, gamma=0.99, lmbda=0.95, return_std=1):
    """
    Function to estimate the advantage function for a series of states and actions using the
    general advantage estimator (GAE).  Drop-in for rlgym_ppo.util.torch_functions.compute_gae.
    :return: Bootstrapped value function estimates, GAE results, returns.
    """
    n = len(rews)
    if n == 0:
        return torch.zeros(0), torch.zeros(0), []
    vt, adv, _, ret64 = compute_gae_device(rews, dones, truncated, values, gamma, lmbda, return_std, n_head=n)
    return vt.cpu(), adv.cpu(), ret64.cpu().tolist()   # the reference's returns are Python floats (f64)
