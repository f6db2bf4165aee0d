"""Differentiable CPU restatement of the value ops  --  TEST INFRASTRUCTURE ONLY.

Same role and rules as ``oracle/pygho_oracle.py`` (never imported by the product).  The
functions below express the reference's arithmetic with stock torch CPU ops so that
gradients come from torch's own autograd: gather = ``index_select``, reduction =
``scatter_reduce(include_self=False)`` on a zero tensor -- the very ATen kernels the
reference reaches through ``backend/utils.py:50-55`` -- so values AND gradients
(including the even split among max/min ties) are the reference's.  Pinned by
``tests/test_oracle_golden.py::test_torch_oracle_matches_golden`` against outputs of the
real reference.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

_RED = {"sum": "sum", "mean": "mean", "max": "amax", "min": "amin"}


def scatter_reduce(src: Tensor, ind: Tensor, dim_size: int, aggr: str) -> Tensor:
    """backend/utils.py:6-56."""
    out = src.new_zeros((int(dim_size),) + tuple(src.shape[1:]))
    if src.shape[0] == 0:
        return out
    idx = ind.reshape((-1,) + (1,) * (src.ndim - 1)).expand_as(src)
    return out.scatter_reduce(0, idx, src, _RED[aggr], include_self=False)


def spspmm(a_val: Optional[Tensor], b_val: Optional[Tensor], acd: Tensor, n_out: int,
           aggr: str = "sum") -> Tensor:
    """backend/Spspmm.py:307-321 (values only)."""
    if a_val is None:
        msg = b_val.index_select(0, acd[2])
    elif b_val is None:
        msg = a_val.index_select(0, acd[1])
    else:
        msg = a_val.index_select(0, acd[1]) * b_val.index_select(0, acd[2])
    return scatter_reduce(msg, acd[0], n_out, aggr)


def spmm(a_ind: Tensor, a_val: Optional[Tensor], a_shape, dim1: int, x: Tensor,
         aggr: str = "sum") -> Tensor:
    """backend/Spmm.py:6-44."""
    src, tar = (a_ind[0], a_ind[1]) if dim1 == 0 else (a_ind[1], a_ind[0])
    n_tar = a_shape[1] if dim1 == 0 else a_shape[0]
    msg = x.index_select(0, src)
    if a_val is not None:
        msg = a_val * msg
    return scatter_reduce(msg, tar, n_tar, aggr)


def sp_pool(ind: Tensor, val: Tensor, shape, keep_dim: int, aggr: str) -> Tensor:
    """backend/SpTensor.py:388-394 (one surviving sparse dim)."""
    return scatter_reduce(val, ind[keep_dim], shape[keep_dim], aggr)


def sp_unpool(ind: Tensor, dim: int, x: Tensor) -> Tensor:
    """backend/SpTensor.py:470-476."""
    return x.index_select(0, ind[dim])


def mamamm(a: Tensor, dim1: int, b: Tensor, dim2: int, mask: Tensor) -> Tensor:
    """backend/Mamamm.py:7-64 with zeroed pads and a masked result (intended semantics)."""
    sa = "bji" if dim1 == 1 else "bij"
    sb = "bjk" if dim2 == 1 else "bkj"
    out = torch.einsum(f"{sa}d,{sb}d->bikd", a, b)
    return out * mask.unsqueeze(-1)


def ma_pool(data: Tensor, mask: Tensor, dims, aggr: str) -> Tensor:
    """backend/MaTensor.py:175-206 with a true minimum."""
    m = mask.reshape(mask.shape + (1,) * (data.ndim - mask.ndim))
    dims = tuple(dims)
    if aggr == "sum":
        return (data * m).sum(dim=dims)
    if aggr == "mean":
        cnt = mask.sum(dim=dims).clamp_min(1)
        return (data * m).sum(dim=dims) / cnt.reshape(cnt.shape + (1,) * (data.ndim - mask.ndim))
    fill = float("-inf") if aggr == "max" else float("inf")
    red = data.masked_fill(~m, fill)
    red = red.amax(dim=dims) if aggr == "max" else red.amin(dim=dims)
    return red.masked_fill(torch.isinf(red), 0.0)
