"""Data-parallel plumbing: graph sharding and a flat-bucket gradient all-reduce.

The reference has no distributed code; the natural partition of
its workload is by graph (a batch is a block-diagonal concatenation, docs/HoData.md),
so each rank owns whole graphs, builds its own plans, and the only exchange per step is
one all-reduce of the flattened gradient bucket (NCCL over NVLink; ``gloo`` in CPU tests).
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist


def shard_by_cost(costs: Sequence[int], world_size: int) -> List[List[int]]:
    """Greedy longest-processing-time partition: item ids per rank, balanced by cost
    (use the per-graph tuple or triple count).  Deterministic."""
    order = sorted(range(len(costs)), key=lambda i: (-int(costs[i]), i))
    loads = [0] * world_size
    parts: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        parts[r].append(i)
        loads[r] += int(costs[i])
    return [sorted(p) for p in parts]


def shard_contiguous(n_items: int, rank: int, world_size: int) -> range:
    """Rank r takes items [r*n/W, (r+1)*n/W)."""
    lo = (n_items * rank) // world_size
    hi = (n_items * (rank + 1)) // world_size
    return range(lo, hi)


class FlatGradBucket:
    """All gradients of a model live in ONE contiguous buffer (``p.grad`` are views), so
    a step needs a single all-reduce and ``zero()`` is a single memset."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dtype = self.params[0].device, self.params[0].dtype
        # every parameter starts at a multiple of 4 elements (16 bytes): the fused kernels add
        # into these views with 128-bit accesses, and the flat optimizer walks the buffer as float4
        self.offsets = []
        total = 0
        for p in self.params:
            self.offsets.append(total)
            total += (p.numel() + 3) // 4 * 4
        self.flat = torch.zeros(total, dtype=dtype, device=dev)
        for p, off in zip(self.params, self.offsets):
            p.grad = self.flat[off:off + p.numel()].view_as(p)

    def zero(self) -> None:
        self.flat.zero_()

    def allreduce_mean(self, group=None) -> None:
        """Average gradients over the ranks (no-op without an initialised group)."""
        if self.allreduce_sum(group) > 1:
            self.flat.div_(dist.get_world_size(group))

    def allreduce_sum(self, group=None) -> int:
        """Sum gradients over the ranks; returns the world size (the caller folds the 1 / world
        of the average into its optimizer step, see :class:`FlatAdamW`)."""
        if not (dist.is_available() and dist.is_initialized()):
            return 1
        world = dist.get_world_size(group)
        if world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        return world

    def nbytes(self) -> int:
        return self.flat.numel() * self.flat.element_size()


class FlatAdamW:
    """AdamW whose parameters, gradients and moments each live in ONE flat buffer: the whole
    optimisation step is a single kernel launch (``pgh_adamw_flat_f32``) instead of torch's
    multi-tensor launches (3 launches / 85 us per SSWL+ step for 5 MB of parameters -- 3 % of a
    128-graph step at the 8-GPU strong-scaling point).  Same update rule as
    ``torch.optim.AdamW`` (decoupled weight decay, bias correction); the step counter is a device
    scalar, so the step is CUDA-graph capturable.  ``bucket`` is the model's
    :class:`FlatGradBucket`; the parameters are re-pointed to views of a flat buffer laid out
    like the bucket (``p.data`` keeps its values)."""

    def __init__(self, bucket: "FlatGradBucket", lr: float = 1e-3, betas=(0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 1e-2):
        self.bucket, self.lr, self.betas, self.eps, self.weight_decay = bucket, lr, betas, eps, weight_decay
        flat = bucket.flat
        if flat.dtype != torch.float32 or not flat.is_cuda:
            raise TypeError("FlatAdamW needs float32 CUDA parameters")
        self.param = torch.zeros_like(flat)
        with torch.no_grad():
            for p, off in zip(bucket.params, bucket.offsets):
                view = self.param[off:off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.steps = torch.zeros((1,), dtype=torch.float32, device=flat.device)

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0) -> None:
        from . import _lib
        flat = self.bucket.flat
        _lib.call("pgh_adamw_flat_f32", self.param.data_ptr(), flat.data_ptr(), self.exp_avg.data_ptr(),
                  self.exp_avg_sq.data_ptr(), flat.numel(), self.steps.data_ptr(), float(self.lr),
                  float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.weight_decay),
                  float(grad_scale), _lib.stream_ptr(flat.device))
        _lib.count_launch()
        self.steps.add_(1.0)

    def zero_grad(self) -> None:
        self.bucket.zero()


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every replica start from rank ``src``'s weights and buffers."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


def enable_sync_batchnorm(module: torch.nn.Module, group=None) -> int:
    """Cross-rank BatchNorm statistics for every fused Linear-BatchNorm-activation block of
    ``module`` (reference semantics: ``honn/utils.py:46-61`` normalises over ALL tuples of the
    batch; under graph sharding that means the tuples of all ranks -- SURVEY.md Q11).  The fused
    block then all-gathers the per-rank (mean, M2, count) triples forward and all-reduces the
    two backward sums (csrc/fused_mlp.cu), which makes N-rank training on a sharded batch equal
    to single-process training on the whole batch.  Returns the number of layers switched."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    grp = group if group is not None else dist.group.WORLD
    n = 0
    for m in module.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m._pgh_sync_group = grp
            n += 1
    return n


def disable_sync_batchnorm(module: torch.nn.Module) -> None:
    for m in module.modules():
        if hasattr(m, "_pgh_sync_group"):
            del m._pgh_sync_group
