"""Host -> HBM staging of rollout arrays.

The reference hands the learner NumPy arrays (batched_agent_manager.py:159-168) and copies them around on the
host (torch.cat in experience_buffer.py:17-37, per-minibatch .to(device) in ppo_learner.py:139-143).  Here each
array crosses PCIe exactly once: arrays that already live in pinned memory (our collection path writes its
trajectories there) are copied with one async cudaMemcpy; anything else goes through a grow-only pinned bounce
buffer.  dtype conversion (f64 -> f32, as torch.as_tensor(..., dtype=float32) does) happens on the device.
"""
import numpy as np
import torch

_F = (torch.float32, torch.float64)


class Stager:
    """Grow-only pinned bounce buffers, one per named slot, so concurrent uploads never alias."""

    def __init__(self, device):
        self.device = torch.device(device)
        self._pinned = {}
        self._events = {}

    def _bounce(self, slot, nbytes):
        buf = self._pinned.get(slot)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(nbytes, 1 << 12), dtype=torch.uint8, pin_memory=True)
            self._pinned[slot] = buf
        return buf

    def to_device(self, x, slot, out=None):
        """Returns a contiguous f32/f64 device tensor holding `x` (async on the current stream).
        `out` (optional): preallocated device tensor of the right dtype/shape to copy into."""
        if isinstance(x, torch.Tensor) and x.is_cuda:
            t = x if x.dtype in _F else x.to(torch.float32)
            t = t.contiguous()
            if out is not None:
                out.copy_(t)
                return out
            return t
        if isinstance(x, torch.Tensor):
            t = x.detach()
            if t.dtype not in _F:
                t = t.to(torch.float32)
            t = t.contiguous()
        else:
            a = np.asarray(x)
            if a.dtype not in (np.float32, np.float64):
                a = a.astype(np.float32)
            t = torch.from_numpy(np.ascontiguousarray(a))
        if out is None:
            out = torch.empty(t.shape, dtype=t.dtype, device=self.device)
        if t.numel() == 0:
            return out
        if t.is_pinned():
            out.copy_(t, non_blocking=True)
            return out
        nbytes = t.numel() * t.element_size()
        # the bounce buffer may still be feeding an earlier async copy
        ev = self._events.get(slot)
        if ev is not None:
            ev.synchronize()
        bounce = self._bounce(slot, nbytes)[:nbytes].view(t.dtype).view(t.shape)
        bounce.copy_(t)
        out.copy_(bounce, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._events[slot] = ev
        return out
