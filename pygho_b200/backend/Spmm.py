"""2-D SparseTensor x dense matrix (reference ``pygho/backend/Spmm.py:6-44``)."""
from __future__ import annotations

import torch
from torch import Tensor

from .. import plans as P
from ..ops import seg_gmr
from .SpTensor import SparseTensor
from .utils import _flatten_dense


def _spmm_plan(A: SparseTensor, dim1: int) -> P.TriplePlan:
    cache = P._cache(A.indices)
    plan = cache.get(("spmm", dim1))
    if plan is None:
        src_row, tar_row = (0, 1) if dim1 == 0 else (1, 0)
        tar = P.to_i32(A.indices[tar_row])
        src = P.to_i32(A.indices[src_row])
        # triples: output row tar[t], operand A row t (the edge value), operand B row src[t]
        plan = P.TriplePlan(A.nnz, A.shape[tar_row], A.nnz, A.shape[src_row], tar, None, src,
                            sorted_by="a" if tar_row == 0 else "d")
        cache[("spmm", dim1)] = plan
    return plan


def spmm(A: SparseTensor, dim1: int, X: Tensor, aggr: str = "sum") -> Tensor:
    """``out[tar] = aggr_{edges (tar, src)} A.values[edge] * X[src]`` where ``dim1`` is the
    contracted (source) dim of ``A``.  ``A.values`` may be None (= 1) or broadcast along
    trailing dims like the reference's elementwise product."""
    assert A.sparse_dim == 2, "can only use 2-dim sparse tensor"
    plan = _spmm_plan(A, dim1)
    val = A.values
    if val is None:
        fx, dshape = _flatten_dense(X)
        out = seg_gmr(None, fx, plan, aggr)
    else:
        if val.shape[1:] != X.shape[1:]:
            dshape = torch.broadcast_shapes(val.shape[1:], X.shape[1:])
            val = val.expand((val.shape[0],) + tuple(dshape))
            X = X.expand((X.shape[0],) + tuple(dshape))
        fv, dshape = _flatten_dense(val)
        fx, _ = _flatten_dense(X)
        out = seg_gmr(fv, fx, plan, aggr)
    return out.reshape((plan.n_out,) + tuple(dshape))
