"""ExperienceBuffer, device resident (replaces rlgym_ppo/ppo/experience_buffer.py).

The reference keeps nine CPU tensors and rebuilds every one of them with torch.cat on each submit
(experience_buffer.py:17-37, 54-80), then fancy-indexes five of them per batch on the CPU and ships every
minibatch to the GPU (:82-102, ppo_learner.py:139-143).  Here the nine fields are rings preallocated in HBM:
logical row i (oldest = 0, the order `_cat` keeps) lives at physical row (start + i) % max_size, submit is one
append kernel per field, and a batch is one permutation-gather kernel.  The permutation itself is NumPy's
legacy RandomState stream (bit-exact by construction: produced on the host by the same MT19937/Fisher-Yates
algorithm, rlppo_host_permutation) so indices and gathered bytes equal the reference's.

A bf16 copy of `states`, zero padded to a multiple of 8 columns, rides along: it is the TMA-legal A operand of
the first-layer tensor-core GEMMs (row stride 16-byte aligned), so the update never re-converts observations.
"""
import numpy as np
import torch

from .. import _lib, ops
from .._staging import Stager

_PERM_POOL = None
_NEXT_UID = [0]
FIELDS = ("states", "actions", "log_probs", "rewards", "next_states", "dones", "truncated", "values", "advantages")
_WIDE = ("states", "next_states")


class ExperienceBuffer(object):
    @staticmethod
    def _cat(t1, t2, size):
        """experience_buffer.py:17-37 -- kept for API compatibility (torch tensors on any device)."""
        if len(t2) > size:
            t = t2[-size:].clone()
        elif len(t2) == size:
            t = t2
        elif len(t1) + len(t2) > size:
            t = torch.cat((t1[len(t2) - size:], t2), 0)
        else:
            t = torch.cat((t1, t2), 0)
        return t

    def __init__(self, max_size, seed, device):
        # The reference's Learner passes device="cpu" here (learner.py:124-126).  The buffer IS the device-side
        # store in this implementation, so "cpu"/"auto" mean "the current CUDA device".
        if device in (None, "cpu", "auto", "gpu"):
            device = "cuda:%d" % torch.cuda.current_device() if torch.cuda.is_available() else "cuda:0"
        self.device = torch.device(device)
        self.seed = seed
        self.max_size = int(max_size)
        self.rng = np.random.RandomState(seed)
        self.capacity = self.max_size
        self.start = 0          # physical row of logical row 0
        self.size = 0           # number of valid rows
        self.obs_dim = None
        self.obs_pad = None
        self.act_dim = 1        # columns of an action row (8 for MultiDiscreteFF, N for ContinuousPolicy)
        self._rings = None
        self.states_bf16 = None
        self._stager = None
        self.start_dev = None
        self.state_dev = None
        self._spec = None
        self._copy_stream = None
        self._pin = [None, None, None]
        self._pin_ev = [None, None, None]
        self._pin_i = 0
        self._pin_of_last = 0
        self._last_rows = 0

    # ---- storage -------------------------------------------------------------------------------------
    def _allocate(self, obs_dim, act_dim=1):
        _lib.require_device()
        self.obs_dim = int(obs_dim)
        self.obs_pad = ops.pad8(self.obs_dim)
        self.act_dim = int(act_dim)
        cap = self.capacity
        self._rings = {}
        for f in FIELDS:
            shape = (cap, self.obs_dim) if f in _WIDE else (cap,)
            if f == "actions" and self.act_dim > 1:
                shape = (cap, self.act_dim)      # multi-dimensional actions stay rows, as the reference's tensor does
            self._rings[f] = torch.zeros(shape, dtype=torch.float32, device=self.device)
        self.states_bf16 = torch.zeros((cap, self.obs_pad), dtype=torch.bfloat16, device=self.device)
        # {start, size} mirrored on the device: the append and gather kernels read the ring position from here, so a
        # captured CUDA graph keeps following the ring as it fills and wraps
        self.state_dev = torch.zeros(2, dtype=torch.int64, device=self.device)
        self.start_dev = self.state_dev[0:1]
        _NEXT_UID[0] += 1
        self.uid = _NEXT_UID[0]     # identity of this set of rings (a freed buffer's addresses can be handed out again)
        self._stager = Stager(self.device)

    def ring(self, field):
        """Physical ring tensor of a field (row i of it is NOT logical row i; see `start`)."""
        return self._rings[field]

    def sync_late(self):
        """Make the current stream wait for a `next_states` block that Learner.add_new_experience is still copying /
        appending on its late stream (nothing on the learner path reads that ring; see learner.py)."""
        ev = getattr(self, "_late_ev", None)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
            self._late_ev = None

    def _logical(self, field):
        if field == "next_states":
            self.sync_late()
        if self._rings is None:
            return torch.empty(0, dtype=torch.float32, device=self.device)
        r = self._rings[field]
        end = self.start + self.size
        if end <= self.capacity:
            return r[self.start:end]
        return torch.cat((r[self.start:], r[:end - self.capacity]), 0)

    def __len__(self):
        return self.size

    # ---- submit (experience_buffer.py:54-80) ------------------------------------------------------------
    def submit_experience(self, states, actions, log_probs, rewards, next_states, dones, truncated, values,
                          advantages):
        new = dict(zip(FIELDS, (states, actions, log_probs, rewards, next_states, dones, truncated, values,
                                advantages)))
        self.sync_late()
        if self._rings is None:
            st = states
            obs_dim = st.shape[1] if hasattr(st, "shape") and len(st.shape) == 2 else np.asarray(st).shape[1]
            self._allocate(obs_dim, _act_dim(actions))
        dev = {f: self._stager.to_device(new[f], "sub." + f) for f in FIELDS}
        self.submit_device(dev)

    def submit_device(self, dev):
        """Append rows that already live in HBM: dict field -> contiguous f32/f64 device tensor."""
        rows = self.append_device(dev)
        if rows:
            self.advance_host(rows)

    def append_device(self, dev):
        """Device half of submit_device (enqueue only, CUDA-graph capturable): all nine rings in one launch, at the
        position held in state_dev, which the launch also advances.  Returns the number of rows appended; the caller
        mirrors it on the host with advance_host()."""
        n = int(dev["rewards"].shape[0])
        for f in FIELDS:
            if f == "next_states" and f not in dev:
                continue        # appended separately by the caller (append_next_states_late)
            assert int(dev[f].shape[0]) == n, f"field {f} has {dev[f].shape[0]} rows, expected {n}"
        if n == 0:
            return 0
        cap = self.capacity
        skip = max(0, n - cap)          # `_cat`: when the new block alone exceeds max_size keep its tail
        rows = n - skip
        fields = [(self._rings[f], (dev[f][skip:] if skip else dev[f]),
                   self.states_bf16 if f == "states" else None) for f in FIELDS if f in dev]
        ops.ring_append_fields(fields, cap, 0, rows, state_dev=self.state_dev)
        return rows

    def append_next_states_late(self, src, stream):
        """Append `src` [n, obs] to the next_states ring on `stream`, at the position the NEXT append_device will use --
        call it BEFORE advance_host.  The position comes from the host mirrors (same arithmetic as the device state)."""
        n = int(src.shape[0])
        cap = self.capacity
        skip = max(0, n - cap)
        rows = n - skip
        pos = (self.start + self.size) % cap
        with torch.cuda.stream(stream):
            ops.ring_append(self._rings["next_states"], pos, src[skip:] if skip else src, rows)
            ev = torch.cuda.Event()
            ev.record(stream)
        self._late_ev = ev

    def advance_host(self, rows):
        """Host mirror of the device-side ring advance (the arithmetic of `_cat`, experience_buffer.py:17-37): rows were
        written at (start + size) % capacity; once full, the oldest rows fall out."""
        cap = self.capacity
        self._last_rows = rows
        over = max(0, self.size + rows - cap)
        self.start = (self.start + over) % cap
        self.size = min(cap, self.size + rows)

    # ---- sampling (experience_buffer.py:82-102) ------------------------------------------------------------
    def _index_tensor(self, indices):
        if isinstance(indices, torch.Tensor):
            return indices.to(device=self.device, dtype=torch.int64).contiguous()
        idx = np.ascontiguousarray(indices, dtype=np.int64)
        return torch.from_numpy(idx).to(self.device, non_blocking=False)

    def _get_samples(self, indices):
        idx = self._index_tensor(indices)
        B = idx.numel()
        out = [torch.empty(B, dtype=torch.float32, device=self.device) for _ in range(4)]
        if self.act_dim > 1:
            out[0] = torch.empty((B, self.act_dim), dtype=torch.float32, device=self.device)
        st = torch.empty((B, self.obs_dim), dtype=torch.float32, device=self.device)
        self.gather(idx, out_actions=out[0], out_logp=out[1], out_values=out[2], out_adv=out[3], out_states=st)
        return out[0], out[1], st, out[2], out[3]

    def gather(self, idx, **outs):
        """Device permutation gather of LOGICAL indices into caller-provided outputs (see ops.gather_batch)."""
        view = _RingView(self)
        if self.act_dim > 1 and outs.get("out_actions") is not None:
            # action ROWS: a second pass of the same kernel over the actions ring as a wide field (exact f32 copy)
            ops.gather_batch(_WideView(self, self._rings["actions"], self.act_dim), idx, out_states=outs.pop("out_actions"))
        ops.gather_batch(view, idx, **outs)

    def next_permutation(self):
        """One `self.rng.permutation(total)` (experience_buffer.py:98), NumPy's legacy stream bit for bit, returned as
        a PINNED int64 tensor (async upload).

        The draw is a sequential ~1 ms host computation, comparable to the GPU time of a whole example-size iteration,
        so the NEXT permutation is always computed speculatively on a worker thread (ctypes releases the GIL) from a
        copy of the generator state, for the buffer length we expect next.  `self.rng` itself only ever advances here, on
        the caller's thread, when a permutation is handed out: if the speculation guessed the wrong length it is thrown
        away and the draw is redone, so the stream is the reference's in every case."""
        total = int(self.size)
        spec, self._spec = self._spec, None
        perm = None
        if spec is not None:
            fut, spec_total, spec_state = spec
            out, key, pos = fut.result()
            st = self.rng.get_state()
            if spec_total == total and st[2] == spec_state[1] and np.array_equal(st[1], spec_state[0]):
                self.rng.set_state((st[0], key, pos, st[3], st[4]))
                perm, self._pin_of_last = out
        if perm is None:
            st = self.rng.get_state()
            (perm, self._pin_of_last), key, pos = self._draw(st[1], st[2], total)
            self.rng.set_state((st[0], key, pos, st[3], st[4]))
        self._speculate(total)
        return perm

    def next_permutation_device(self):
        """next_permutation() uploaded to the device (async); the pinned buffer is fenced with an event so the
        speculative draw two calls later cannot overwrite it before the copy engine has read it."""
        perm = self.next_permutation()
        dev = perm.to(self.device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._pin_ev[self._pin_of_last] = ev
        return dev

    def next_permutation_into(self, dst, fresh_alloc=False):
        """next_permutation() uploaded (async) into the caller's int64 device buffer `dst` (len == len(self)): a fixed
        address, so a captured CUDA graph can read its indices from it."""
        perm = self.next_permutation()
        assert dst.numel() == perm.numel() and dst.dtype == torch.int64
        # on a side stream: the copy engine uploads the indices while the device is still busy with the work enqueued
        # before (value inference, GAE, ring appends of this iteration); the caller's stream only waits for the event
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        # (readers of dst's previous contents are done: PPOLearner.learn synchronises before it returns)
        main = torch.cuda.current_stream()
        # A freshly allocated `dst` (fresh_alloc) may be a block the caching allocator just recycled from main-stream
        # temporaries whose kernels are still in flight: then the side-stream copy must not start before them.  In steady
        # state the caller reuses one buffer and the copy overlaps the main stream's work (waiting here every time cost
        # 0.17 ms per end-to-end step).
        if fresh_alloc:
            self._copy_stream.wait_stream(main)
        with torch.cuda.stream(self._copy_stream):
            dst.copy_(perm, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        dst.record_stream(self._copy_stream)
        self._pin_ev[self._pin_of_last] = ev
        main.wait_event(ev)

    def _draw(self, key, pos, total):
        self._pin_i = (self._pin_i + 1) % 3
        ev = self._pin_ev[self._pin_i]
        if ev is not None:
            ev.synchronize()
            self._pin_ev[self._pin_i] = None
        buf = self._pin[self._pin_i]
        if buf is None or buf.numel() < total:
            buf = torch.empty(max(total, 1), dtype=torch.int64).pin_memory() if torch.cuda.is_available() \
                else torch.empty(max(total, 1), dtype=torch.int64)
            self._pin[self._pin_i] = buf
        view = buf[:total]
        _, key, pos = _lib.host_permutation_raw(key, pos, total, out=view.numpy())
        return (view, self._pin_i), key, pos

    def _speculate(self, total):
        if total < 4096:
            return                      # cheaper to draw on demand
        global _PERM_POOL
        if _PERM_POOL is None:
            from concurrent.futures import ThreadPoolExecutor
            _PERM_POOL = ThreadPoolExecutor(max_workers=1, thread_name_prefix="rlppo-perm")
        st = self.rng.get_state()
        key, pos = st[1].copy(), int(st[2])
        guess = min(self.capacity, total + self._last_rows) if total < self.capacity else total
        self._spec = (_PERM_POOL.submit(self._draw, key, pos, guess), guess, (key, pos))

    def get_all_batches_shuffled(self, batch_size):
        total_samples = self.size
        idx_dev = self.next_permutation_device()
        start_idx = 0
        while start_idx + batch_size <= total_samples:
            yield self._get_samples(idx_dev[start_idx:start_idx + batch_size])
            start_idx += batch_size

    def clear(self):
        self.__init__(self.max_size, self.seed, self.device)


class _RingView:
    """What ops.gather_batch reads: physical rings + (capacity, start)."""

    def __init__(self, buf):
        r = buf._rings
        self.actions, self.log_probs = r["actions"], r["log_probs"]
        self.values, self.advantages = r["values"], r["advantages"]
        self.states, self.states_bf16 = r["states"], buf.states_bf16
        self.obs_dim, self.capacity, self.start = buf.obs_dim, buf.capacity, buf.start
        self.start_dev = buf.start_dev


class _WideView:
    """A [capacity, width] ring presented to ops.gather_batch as its `states` field (scalar fields absent)."""

    def __init__(self, buf, ring, width):
        self.actions = self.log_probs = self.values = self.advantages = None
        self.states, self.states_bf16 = ring, None
        self.obs_dim, self.capacity, self.start = int(width), buf.capacity, buf.start
        self.start_dev = buf.start_dev


def _act_dim(actions):
    shape = tuple(actions.shape) if hasattr(actions, "shape") else np.asarray(actions).shape
    return int(shape[1]) if len(shape) == 2 else 1


def _make_field_property(name):
    def get(self):
        return self._logical(name)
    return property(get, doc=f"`{name}` in logical (oldest -> newest) order, as the reference's tensor attribute")


for _f in FIELDS:
    setattr(ExperienceBuffer, _f, _make_field_property(_f))
