"""SparseTensor x MaskedTensor contraction (reference ``pygho/backend/Spmamm.py:12-67``): the
"SD" mode, a sparse batched adjacency (b, n, m) against dense tuple features.

``out[b, i, *r, :] = aggr_{j : A[b, i, j] != 0}  A.values[b, i, j, :] * B[b, j, *r, :]``
(for ``dim1 = 2``; ``dim1 = 1`` contracts A's dim 1 and keeps its dim 2), with the contracted
position of ``B`` given by ``dim2`` and ``r`` ranging over B's other masked dims.  Every
(non-zero, r) pair is one entry of a gather-multiply-segmented-reduce plan, cached on
``A.indices``; the values go through the same ``seg_gmr`` kernel as spspmm.

Intended semantics, not the reference's accidents: masked-out positions of B
never contribute (the reference's ``masked_fill`` result is dropped, Spmamm.py:60), A's values
broadcast over ALL other dims of B (the reference's ``unsqueeze(1)`` only works for exactly one),
rows that receive nothing are 0.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import BoolTensor

from .. import ops
from .. import plans as P
from .MaTensor import MaskedTensor
from .SpTensor import SparseTensor


def _prod(xs) -> int:
    out = 1
    for x in xs:
        out *= int(x)
    return out


def _plan(A: SparseTensor, dim1: int, n: int, m: int, K: int, bmask: Optional[BoolTensor],
          dim2: int):
    """TriplePlan with one entry per (non-zero of A, other position r): output row
    ``(b*n + i)*K + r``, A row ``t``, B row ``(b*m + j)*K + r``.  With ``bmask`` (B's mask;
    max / min only) the entries whose B position is masked out are dropped.  Cached on
    ``A.indices`` (per mask object: masks are shared by all layers of a model)."""
    ind = A.indices
    cache = P._cache(ind)
    key = ("spmamm", dim1, dim2, n, m, K, None if bmask is None else id(bmask))
    hit = cache.get(key)
    if hit is not None:
        return hit[0]
    tmask = None if bmask is None else torch.movedim(bmask, dim2, 1).contiguous()
    bidx = ind[0]
    cidx, tidx = (ind[1], ind[2]) if dim1 == 1 else (ind[2], ind[1])
    nnz = ind.shape[1]
    r = torch.arange(K, device=ind.device)
    a = ((bidx * n + tidx).unsqueeze(1) * K + r).reshape(-1)
    c = torch.arange(nnz, device=ind.device).unsqueeze(1).expand(-1, K).reshape(-1)
    d = ((bidx * m + cidx).unsqueeze(1) * K + r).reshape(-1)
    acd = torch.stack((a, c, d))
    if tmask is not None:
        acd = acd[:, tmask.reshape(-1)[d]].contiguous()
    plan = P.plan_from_acd(acd, int(A.shape[0]) * n * K, nnz, int(A.shape[0]) * m * K)
    cache[key] = (plan, acd, bmask)          # keeps the keyed mask object alive
    return plan


def spmamm(A: SparseTensor, dim1: int, B: MaskedTensor, dim2: int,
           mask: Optional[BoolTensor] = None, aggr: str = "sum") -> MaskedTensor:
    assert A.sparse_dim == 3, f"A should have 3 sparse dims, but input has {A.sparse_dim}"
    assert aggr != "mean", "not implemented"
    if dim1 == 1:
        n = int(A.shape[2])
    elif dim1 == 2:
        n = int(A.shape[1])
    else:
        raise NotImplementedError
    md = B.masked_dim
    assert 1 <= dim2 < md, "dim2 must be a masked, non-batch dim of B"
    b = int(A.shape[0])
    tB = torch.movedim(B.fill_masked(0.0), dim2, 1)
    m = int(tB.shape[1])
    others, dense = tuple(tB.shape[2:md]), tuple(tB.shape[md:])
    K, D = _prod(others), _prod(dense)
    plan = _plan(A, dim1, n, m, K, B.mask if aggr in ("max", "min") else None, dim2)
    aval = A.values
    if aval is not None:
        assert tuple(aval.shape[1:]) == dense, "A.values and B must have the same dense shape"
        aval = aval.reshape(aval.shape[0], D)
    out = ops.seg_gmr(aval, tB.reshape(b * m * K, D), plan, aggr)
    ret = torch.movedim(out.reshape((b, n) + others + dense), 1, dim2)
    return MaskedTensor(ret, mask if mask is not None else B.mask)
