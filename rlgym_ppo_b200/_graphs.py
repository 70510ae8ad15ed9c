"""CUDA-graph cache: replay a fixed sequence of C-ABI launches instead of re-issuing it from Python.

The example-size learner iteration is ~40 launches of 2-100 us kernels; issued one by one from Python (ctypes marshalling,
torch bookkeeping, the driver's launch path) the host cannot keep the device busy.  A body of enqueue-only code is run
eagerly the first time a key is seen (workspaces get allocated, kernel attributes configured), captured on the second
use, and replayed from then on.  Keys must name everything the launches bake in: buffer addresses (identities /
generations), shapes, scalar arguments.  Ring positions, step counters, learning rates and the return-normalisation
scale are read from device memory by the kernels, so they are NOT part of a key.
"""
from collections import OrderedDict

import torch

from . import _lib


class GraphCache:
    def __init__(self, max_entries=12):
        self._graphs = OrderedDict()
        self._warm = set()
        self.max_entries = max_entries

    def replay(self, key, body):
        """True if `body`'s device work was enqueued by replaying (or just capturing and replaying) its graph; False
        if the caller has to run `body()` eagerly (first sighting of `key`)."""
        entry = self._graphs.get(key)
        if entry is None:
            if key not in self._warm:
                self._warm.add(key)
                return False
            calls0 = _lib.CALLS
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                body()
            entry = (graph, _lib.CALLS - calls0)
            _lib.CALLS = calls0
            self._graphs[key] = entry
            while len(self._graphs) > self.max_entries:
                self._graphs.popitem(last=False)
        else:
            self._graphs.move_to_end(key)
        entry[0].replay()
        _lib.CALLS += entry[1]
        return True
