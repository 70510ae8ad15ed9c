"""Shared plumbing of the policies whose head is a per-row kernel over the logits of the last Linear
(MultiDiscreteFF, ContinuousPolicy): staging, sampler stream, the training hook PPOLearner calls."""
import numpy as np
import torch
import torch.nn as nn

from ._mlp import Stack


class HeadPolicy(nn.Module):
    act_width = 1       # columns of the stored action rows

    def _init_stack(self, device, precision):
        self._stack = Stack(self.model, device, precision)
        dev = self._stack.device
        self._stack.bind(torch.zeros(self._stack.n_params, device=dev), torch.zeros(self._stack.n_params, device=dev))
        # Philox stream of the sampler: seeded from torch's generator at construction (Learner seeds torch first,
        # learner.py:95-97), advanced by the number of rows sampled
        self._seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        self._offset = 0
        self._obs_stats = None

    def _apply(self, fn, recurse=True):
        probe = fn(torch.empty(0, device=self._stack.device))
        if probe.device != self._stack.device or probe.dtype != torch.float32:
            raise RuntimeError(f"{type(self).__name__} lives on its CUDA device in fp32; it cannot be moved or cast")
        return self

    def load_state_dict(self, state_dict, strict=True, assign=False):
        if assign:
            raise RuntimeError("assign=True would detach the parameters from the flat arena the kernels read")
        out = super().load_state_dict(state_dict, strict=strict)
        self._stack.refresh_operands(force=True)
        return out

    def _stage_obs(self, obs):
        st = self._stack
        if not isinstance(obs, torch.Tensor):
            obs = torch.as_tensor(np.asarray(obs))
        if obs.dtype != torch.float32:
            obs = obs.to(torch.float32)
        lead = tuple(obs.shape[:-1])
        obs = obs.reshape(-1, st.in_dim)
        if not obs.is_cuda:
            obs = obs.to(st.device, non_blocking=True)
        obs = obs.contiguous()
        n = obs.shape[0]
        ws = st.workspace(n)
        if self._obs_stats is not None:
            mean, std, clip = self._obs_stats
            st.stage_rows(obs, ws["x"], mean, std, clip)
        else:
            st.stage_rows(obs, ws["x"])
        return ws["x"], n, ws, lead

    def _logits(self, obs):
        """obs -> (z view of the last Linear's output, n rows, workspace, leading shape)."""
        st = self._stack
        st.refresh_operands()
        x, n, ws, lead = self._stage_obs(obs)
        if n == 0:
            return None, 0, ws, lead
        h = st.forward_hidden(x, n, ws)
        return st.logits(h, n, ws), n, ws, lead

    def _logits_f32(self, zview, n):
        z, parts, ps = zview
        st = self._stack
        return sum(z[:n, q * ps:q * ps + st.out_dim].float() for q in range(parts))

    # PPOLearner._train_chunk: forward through the head + loss + backward into dL/dH_last
    def train_head(self, h, M, ws, actions, old_logp, adv, inv_b, clip, ent_coef, metrics, logp_out=None):
        st = self._stack
        zview = st.logits(h, M, ws)
        self._head_train_kernel(zview, M, actions, old_logp, adv, inv_b, clip, ent_coef, st.dz_view(ws), metrics, logp_out)
        return st.head_backward(h, M, ws)

    def _backprop_forward(self, obs, acts):
        """Forward values of get_backprop_data: (log-probs [n], mean entropy)."""
        st = self._stack
        zview, n, ws, _ = self._logits(obs)
        dev = st.device
        acts_f = torch.as_tensor(acts).to(device=dev, dtype=torch.float32).reshape(n, -1).contiguous()
        zeros = torch.zeros(n, dtype=torch.float32, device=dev)
        logp = torch.empty(n, dtype=torch.float32, device=dev)
        metrics = torch.zeros(8, dtype=torch.float32, device=dev)
        self._head_train_kernel(zview, n, acts_f, zeros, zeros, 0.0, 0.2, 0.0, st.dz_view(ws), metrics, logp)
        return logp, metrics[0] / metrics[4]
