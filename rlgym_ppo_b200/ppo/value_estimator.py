"""ValueEstimator on the B200 kernels (replaces rlgym_ppo/ppo/value_estimator.py).

Linear/ReLU stack as bf16 tcgen05 GEMMs; the final Linear(., 1) is a fused SIMT pass over the last hidden
activation (rlppo_value_head) -- a 1-wide output would waste a 128-row tensor-core tile.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import ops
from ._mlp import Stack, build_sequential


class ValueEstimator(nn.Module):
    def __init__(self, input_shape, layer_sizes, device, precision=None):
        super().__init__()
        self.device = device
        self.model = build_sequential(input_shape, layer_sizes, 1, softmax=False)   # value_estimator.py:19-28
        self._stack = Stack(self.model, device, precision)
        dev = self._stack.device
        self._stack.bind(torch.zeros(self._stack.n_params, device=dev), torch.zeros(self._stack.n_params, device=dev))

    def load_state_dict(self, state_dict, strict=True, assign=False):
        """Stock nn.Module.load_state_dict into the arena views, then the bf16 GEMM operands are rebuilt at once."""
        if assign:
            raise RuntimeError("assign=True would detach the parameters from the flat arena the kernels read")
        out = super().load_state_dict(state_dict, strict=strict)
        self._stack.refresh_operands(force=True)
        return out

    def _apply(self, fn, recurse=True):
        probe = fn(torch.empty(0, device=self._stack.device))
        if probe.device != self._stack.device or probe.dtype != torch.float32:
            raise RuntimeError("ValueEstimator lives on its CUDA device in fp32; it cannot be moved or cast")
        return self

    def values_from_bf16(self, x, n, out=None):
        """x: bf16 [>=n, in_pad] device rows -> f32 [n] values (no host sync)."""
        st = self._stack
        st.refresh_operands()
        if out is None:
            out = torch.empty(n, dtype=torch.float32, device=st.device)
        if n == 0:
            return out
        if st.fused_ok:
            ops.value_infer_fused(st.fused_net(x.stride(0)), x, n, st.w[-1], out)
            return out
        ws = st.workspace(n)
        h = st.forward_hidden(x, n, ws)
        st.value_head_infer(h, n, out)
        return out

    def forward(self, x):
        """[n, obs] (numpy incl. float64, list or tensor) -> [n, 1] f32 device tensor (value_estimator.py:30-36)."""
        st = self._stack
        if not isinstance(x, torch.Tensor):
            x = torch.as_tensor(np.asarray(x))
        lead = tuple(x.shape[:-1])
        x = x.reshape(-1, st.in_dim)
        if not x.is_cuda:
            x = x.to(st.device, non_blocking=True)
        if x.dtype != torch.float32:
            x = x.to(torch.float32)   # as_tensor(dtype=float32), :35
        x = x.contiguous()
        n = x.shape[0]
        ws = st.workspace(n)
        st.stage_rows(x, ws["x"])
        return self.values_from_bf16(ws["x"], n).view(*lead, 1)
