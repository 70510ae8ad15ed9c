"""MultiDiscreteFF on the B200 kernels (replaces rlgym_ppo/ppo/multi_discrete_policy.py).

21 logits parametrise eight categorical distributions (bins 3,3,3,3,3,2,2,2; torch_functions.MultiDiscreteRolv).  Same
constructor, `.model` state-dict, `get_output` / `get_action` / `get_backprop_data` as the reference; the Linear/ReLU
stack runs on the tcgen05 GEMMs, the distributions in rlppo_head_multi_discrete_{sample,train}.
"""
import numpy as np
import torch

from .. import ops
from ._head_policy import HeadPolicy
from ._mlp import build_sequential

BINS = (3, 3, 3, 3, 3, 2, 2, 2)      # multi_discrete_policy.py:21


class MultiDiscreteFF(HeadPolicy):
    act_width = len(BINS)

    def __init__(self, input_shape, layer_sizes, device, precision=None):
        super().__init__()
        self.device = device
        self.splits = list(BINS)
        self.model = build_sequential(input_shape, layer_sizes, sum(BINS), softmax=False)     # :22-33
        self._init_stack(device, precision)

    def get_output(self, obs):
        """The 21 logits (multi_discrete_policy.py:37-45), f32 on the device."""
        zview, n, _, lead = self._logits(obs)
        if n == 0:
            return torch.empty((*lead, sum(BINS)), dtype=torch.float32, device=self._stack.device)
        return self._logits_f32(zview, n).view(*lead, sum(BINS))

    def forward(self, obs):
        return self.get_output(obs)

    def get_action_device(self, obs, deterministic=False):
        st = self._stack
        zview, n, _, lead = self._logits(obs)
        acts = torch.empty((n, 8), dtype=torch.float32, device=st.device)
        logp = torch.empty(n, dtype=torch.float32, device=st.device)
        if n:
            z, parts, ps = zview
            ops.head_multi_discrete_sample(z, parts, ps, n, acts, logp, seed=self._seed, offset=self._offset,
                                           deterministic=deterministic)
            self._offset += n
        return acts, logp, lead

    def get_action(self, obs, deterministic=False):
        """(action [.., 8] int64, log_prob [..]) as CPU tensors (:47-72).  The deterministic branch returns the reference's
        layout: a NumPy array with the EIGHT DISTRIBUTIONS FIRST (`torch.stack(action)`, :60-66) and 0."""
        acts, logp, lead = self.get_action_device(obs, bool(deterministic))
        a = acts.to(torch.int64).view(*lead, 8)
        if deterministic:
            return np.moveaxis(a.cpu().numpy(), -1, 0), 0
        return a.cpu(), logp.view(*lead).cpu()

    def _head_train_kernel(self, zview, M, actions, old_logp, adv, inv_b, clip, ent_coef, dzview, metrics, logp_out):
        z, parts, ps = zview
        dz, dparts, dps, dcols = dzview
        ops.head_multi_discrete_train(z, parts, ps, M, actions, old_logp, adv, inv_b, clip, ent_coef, dz, dparts, dps,
                                      dcols, metrics, logp_out=logp_out)

    def get_backprop_data(self, obs, acts):
        """(summed log-prob of `acts` [n], mean entropy) (:74-89); forward values only (the backward is fused into
        PPOLearner.learn's kernels)."""
        return self._backprop_forward(obs, acts)
