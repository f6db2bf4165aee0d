"""``torch_scatter_reduce`` -- the single reduction primitive of the reference
(``pygho/backend/utils.py:6-56``) as a deterministic segmented reduce on the GPU."""
from __future__ import annotations

from typing import Tuple

import torch
from torch import LongTensor, Tensor

from .. import _lib
from .. import plans as P
from ..ops import seg_gmr


def _flatten_dense(values: Tensor) -> Tuple[Tensor, Tuple[int, ...]]:
    """(rows, *dense) -> ((rows, prod(dense)) contiguous, dense shape)."""
    dshape = tuple(values.shape[1:])
    width = 1
    for s_ in dshape:
        width *= int(s_)
    flat = values.reshape(values.shape[0], width)
    return (flat if flat.is_contiguous() else flat.contiguous()), dshape


def _seg_reduce_int(values: Tensor, plan: P.TriplePlan, reduce: str) -> Tensor:
    """Integer values (tuple features while coalescing): no autograd, int64 kernel."""
    if values.dtype not in (torch.int64, torch.int32, torch.int16, torch.uint8, torch.bool):
        raise TypeError(f"unsupported value dtype {values.dtype}: float32 or integer expected")
    _lib.require_cuda(values)
    flat, dshape = _flatten_dense(values.to(torch.int64))
    g = plan.group("a")
    if g.rowptr is None:
        raise RuntimeError("integer reduce needs a CSR grouping")
    out = torch.empty((plan.n_out, flat.shape[1]), dtype=torch.int64, device=values.device)
    if out.numel():
        P._launch("pgh_seg_reduce_i64", P.ptr(flat), P.ptr(g.first), P.ptr(g.rowptr), plan.n_out,
                  flat.shape[1], _lib.AGGR_CODE[reduce], P.ptr(out), P.stream_ptr(values.device))
    return out.reshape((plan.n_out,) + dshape).to(values.dtype)


def torch_scatter_reduce(dim: int, src: Tensor, ind: LongTensor, dim_size: int,
                         aggr: str) -> Tensor:
    """out[i] = aggr of the rows of ``src`` with ``ind == i``; rows that receive nothing
    are 0 for every ``aggr`` (zero init + ``include_self=False`` in the reference).

    The CSR plan of ``ind`` is built once and cached on the ``ind`` tensor."""
    assert dim == 0, "other dim not implemented"
    assert ind.ndim == 1, "indice must be 1-d"
    plan = P.plan_from_key(ind, dim_size)
    if src.dtype != torch.float32:
        return _seg_reduce_int(src, plan, aggr)
    flat, dshape = _flatten_dense(src)
    return seg_gmr(flat, None, plan, aggr).reshape((dim_size,) + dshape)
