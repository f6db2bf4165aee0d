"""CUDA-graph replay of a whole training step on a batch that stays resident in HBM.

One SSWL+ step is ~470 kernel launches; enqueueing them from Python takes 14.8 ms of host
time against 16.4 ms of device time (profiles/r1_graph_probe.json), so the step is within
10 % of being launch-bound on ONE process and becomes host-bound as soon as eight ranks
share the host's cores.  Every operator of this package is stream-ordered, allocation-free
inside the C ABI and free of host synchronisation once the plans of a batch are cached, so
the whole step (forward, loss, backward, NCCL all-reduce, fused AdamW with
``capturable=True``) can be captured once per resident batch and replayed with a single
launch.  Batches of a different shape need their own graph; the end-to-end (host-fed) path
therefore stays eager.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import _lib


class StepGraph:
    """Capture ``step_fn()`` (no arguments; closes over a device-resident batch, the model
    and a ``capturable`` optimizer) into a CUDA graph.

    ``warmup`` eager calls run first on a side stream (allocator warm-up, lazy plan
    builds, cuBLAS handle/workspace creation -- none of which may happen under capture).
    ``replay()`` launches the graph on the current stream and returns the static output
    (e.g. the loss tensor, overwritten by every replay).  ``launches`` is the number of
    pygho_b200 kernel launches one replay performs."""

    def __init__(self, step_fn: Callable[[], Optional[torch.Tensor]], warmup: int = 2,
                 pool=None):
        if not torch.cuda.is_available():
            raise RuntimeError("StepGraph needs a CUDA device")
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                step_fn()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        before = _lib.launches()
        with torch.cuda.graph(self.graph, pool=pool):
            self.output = step_fn()
        self.launches = _lib.launches() - before

    def replay(self):
        self.graph.replay()
        return self.output

    def pool(self):
        return self.graph.pool()
