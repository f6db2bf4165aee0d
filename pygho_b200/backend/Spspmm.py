"""Sparse x sparse contraction over shared sparse dims -- plans and value ops.

API mirror of the reference ``pygho/backend/Spspmm.py`` (``ptr2batch`` :9, ``deg2batch``
:34, ``spspmm_ind`` :57, ``spsphadamard_ind`` :146, ``filterind`` :186, ``spsphadamard``
:225, ``spspmm`` :270, ``spspmpnn`` :334).  Plans are built by the device plan kernels
(stable radix sorts, binary searches, scans); the value ops are single fused
gather-multiply-segmented-reduce launches over a cached CSR regrouping of ``acd``.
"""
from __future__ import annotations

import warnings
from typing import Callable, Optional, Tuple

import torch
from torch import LongTensor, Tensor

from .. import plans as P
from ..ops import seg_gmr
from .SpTensor import SparseTensor
from .utils import _flatten_dense, torch_scatter_reduce


def deg2batch(deg: LongTensor, dim_size: Optional[int] = None) -> LongTensor:
    """``[0]*deg[0] + [1]*deg[1] + ...`` (reference Spspmm.py:34-54).  Built as the
    expansion of the ranges [cumsum-deg, cumsum) with the match/expand plan kernels."""
    assert deg.ndim == 1, "ptr should be 1-d"
    dev = P._lib.require_cuda(deg)
    n = deg.numel()
    ptr = torch.zeros((n + 1,), dtype=torch.int64, device=dev)
    torch.cumsum(deg, 0, out=ptr[1:])
    return ptr2batch(ptr, dim_size)


def ptr2batch(ptr: LongTensor, dim_size: Optional[int] = None) -> LongTensor:
    """``batch[ptr[i]:ptr[i+1]] = i`` (reference Spspmm.py:9-31)."""
    assert ptr.ndim == 1, "ptr should be 1-d"
    dev = P._lib.require_cuda(ptr)
    total = int(ptr[-1].item()) if dim_size is None else int(dim_size)
    if __debug__:
        assert int(ptr[0].item()) == 0 and bool((ptr[1:] >= ptr[:-1]).all()), \
            "should put in a ptr tensor"
        assert int(ptr[-1].item()) == total, "dim_size should match ptr"
    m = ptr.numel() - 1
    lo = torch.zeros((max(m, 1),), dtype=torch.int32, device=dev)
    c = torch.empty((total,), dtype=torch.int32, device=dev)
    d = torch.empty((total,), dtype=torch.int32, device=dev)
    if total:
        P._launch("pgh_expand_pairs", P.ptr(ptr.contiguous()), P.ptr(lo), None, m, total,
                  P.ptr(c), P.ptr(d), P.stream_ptr(dev))
    return P.to_i64(c)


def spspmm_ind(ind1: LongTensor, dim1: int, ind2: LongTensor, dim2: int,
               is_k2_sorted: bool = False) -> Tuple[LongTensor, LongTensor]:
    """Contraction plan: eliminate ``dim1`` of ``ind1`` against ``dim2`` of ``ind2``.

    Returns ``tarind`` (sorted unique output coordinates = remaining dims of ind1 then of
    ind2) and ``bcd`` (3, T0): ``val1[c] * val2[d]`` contributes to output column ``b``.
    ``bcd`` is sorted by (b, c) -- canonical, where the reference's order inside one
    output segment is unspecified (unstable argsort, Spspmm.py:142)."""
    assert 0 <= dim1 < ind1.shape[0], f"ind1's reduced dim {dim1} is out of range"
    assert 0 <= dim2 < ind2.shape[0], f"ind2's reduced dim {dim2} is out of range"
    k2_sorted = bool(is_k2_sorted) or dim2 == 0
    if __debug__ and k2_sorted:
        assert P.is_sorted(ind2[dim2]), "ind2[0] should be sorted"
    tarind, b, c, d = P.spspmm_ind_i32(ind1, dim1, ind2, dim2, k2_sorted)
    bcd = torch.stack((P.to_i64(b), P.to_i64(c), P.to_i64(d)))
    plan = P.TriplePlan(b.numel(), tarind.shape[1], ind1.shape[1], ind2.shape[1], b, c, d,
                        sorted_by="a")
    P._cache(bcd)[("acd", tarind.shape[1], ind1.shape[1], ind2.shape[1])] = plan
    bcd._pgh_i32 = (b, c, d)
    return tarind, bcd


def spsphadamard_ind(tar_ind: LongTensor, ind: LongTensor) -> LongTensor:
    """``b2a[i]`` = column of ``tar_ind`` equal to ``ind[:, i]`` or -1
    (reference Spspmm.py:146-183)."""
    assert tar_ind.shape[0] == ind.shape[0]
    tkey = P.pack_keys(tar_ind, check=__debug__)
    if __debug__:
        assert P.is_sorted(tkey, strict=True), "tar_ind should be sorted and coalesce"
    return P.to_i64(P.lookup_sorted(tkey, P.pack_keys(ind, check=__debug__)))


def filterind(tar_ind: LongTensor, ind: LongTensor, bcd: LongTensor) -> LongTensor:
    """Restrict a plan to the output pattern ``tar_ind`` (Hadamard with the product):
    ``acd = (b2a[b], c, d)`` for the triples whose output coordinate is present
    (reference Spspmm.py:186-222)."""
    i32 = getattr(bcd, "_pgh_i32", None)
    if i32 is None:
        bcd = bcd.contiguous()
        i32 = (P.to_i32(bcd[0]), P.to_i32(bcd[1]), P.to_i32(bcd[2]))
    a, c, d = P.filter_triples(tar_ind, ind, *i32, check=__debug__)
    if a.numel() == 0:
        return torch.zeros((3, 0), dtype=torch.int64, device=tar_ind.device)
    acd = torch.stack((P.to_i64(a), P.to_i64(c), P.to_i64(d)))
    acd._pgh_i32 = (a.contiguous(), c.contiguous(), d.contiguous())
    return acd


def _dense_pair(va: Optional[Tensor], vb: Optional[Tensor]):
    """Flatten the dense dims of the two operands to one common width (broadcasting
    trailing dims like the reference's elementwise product does)."""
    if va is None or vb is None:
        v = va if va is not None else vb
        flat, dshape = _flatten_dense(v)
        return (flat, None, dshape) if va is not None else (None, flat, dshape)
    if va.shape[1:] != vb.shape[1:]:
        dshape = torch.broadcast_shapes(va.shape[1:], vb.shape[1:])
        va = va.expand((va.shape[0],) + tuple(dshape))
        vb = vb.expand((vb.shape[0],) + tuple(dshape))
    fa, dshape = _flatten_dense(va)
    fb, _ = _flatten_dense(vb)
    return fa, fb, dshape


def spsphadamard(A: SparseTensor, B: SparseTensor, b2a: Optional[LongTensor] = None
                 ) -> SparseTensor:
    """Elementwise product on the common sparsity pattern (reference Spspmm.py:225-267)."""
    assert A.is_coalesced(), "A should be coalesced"
    assert B.is_coalesced(), "B should be coalesced"
    assert A.sparseshape == B.sparseshape, "A, B should be of the same sparse shape"
    if b2a is None:
        b2a = spsphadamard_ind(A.indices, B.indices)
    keep = torch.nonzero(b2a >= 0).flatten()           # pattern bookkeeping (host-sized)
    retind = B.indices[:, keep]
    n = keep.numel()
    plan = P.TriplePlan(n, n, A.nnz, B.nnz, None, P.to_i32(b2a[keep]), P.to_i32(keep))
    va, vb, dshape = _dense_pair(A.values, B.values)
    val = seg_gmr(va, vb, plan, "sum").reshape((n,) + tuple(dshape))
    return SparseTensor(retind, val, shape=A.sparseshape + tuple(val.shape[1:]), is_coalesced=True)


def _out_sparseshape(A: SparseTensor, dim1: int, B: SparseTensor, dim2: int):
    return (A.sparseshape[:dim1] + A.sparseshape[dim1 + 1:] + B.sparseshape[:dim2] +
            B.sparseshape[dim2 + 1:])


def spspmm(A: SparseTensor, dim1: int, B: SparseTensor, dim2: int, aggr: str = "sum",
           bcd: Optional[LongTensor] = None, tar_ind: Optional[LongTensor] = None,
           acd: Optional[LongTensor] = None) -> SparseTensor:
    """``out[a] = aggr_t A.values[acd[1,t]] * B.values[acd[2,t]]`` over ``acd[0,t] == a``
    on the pattern ``tar_ind`` (reference Spspmm.py:270-331).  An operand without values
    counts as 1.  One fused kernel launch; the CSR regrouping of ``acd`` is cached on the
    ``acd`` tensor, so all layers of a model share it."""
    assert A.is_coalesced(), "A should be coalesced"
    assert B.is_coalesced(), "B should be coalesced"
    if acd is None:
        warnings.warn("acd is not found")
        ind = None
        if bcd is None:
            ind, bcd = spspmm_ind(A.indices, dim1, B.indices, dim2)
        if tar_ind is not None:
            if ind is None:
                raise ValueError("bcd without acd needs the product pattern it refers to; "
                                 "pass acd instead (the reference fails here too, Q6)")
            acd = filterind(tar_ind, ind, bcd)
        else:
            warnings.warn("tar_ind is not found")
            if ind is None:
                raise ValueError("bcd given without tar_ind")
            acd, tar_ind = bcd, ind
    assert tar_ind is not None
    plan = P.plan_from_acd(acd, tar_ind.shape[1], A.nnz, B.nnz)
    va, vb, dshape = _dense_pair(A.values, B.values)
    val = seg_gmr(va, vb, plan, aggr).reshape((tar_ind.shape[1],) + tuple(dshape))
    return SparseTensor(tar_ind, val, shape=_out_sparseshape(A, dim1, B, dim2) + tuple(val.shape[1:]),
                        is_coalesced=True)


def spspmpnn(A: SparseTensor, dim1: int, B: SparseTensor, dim2: int, C: SparseTensor,
             acd: LongTensor, message_func: Callable[[Tensor, Tensor, Tensor, LongTensor], Tensor],
             aggr: str = "sum") -> SparseTensor:
    """Message passing with a user message function (reference Spspmm.py:334-380).  The
    callable is arbitrary Python, so the per-triple operands are materialised with the
    gather kernel and only the final reduction is the fused segmented reduce."""
    plan = P.plan_from_acd(acd, C.nnz, A.nnz, B.nnz)

    def gathered(X: SparseTensor, which: str, n_src: int):
        if X.values is None:
            return None
        flat, dshape = _flatten_dense(X.values)
        gp = _gather_rows_plan(plan, which, n_src)
        return seg_gmr(flat, None, gp, "sum").reshape((plan.T,) + dshape)

    mult = message_func(gathered(A, "c", A.nnz), gathered(B, "d", B.nnz),
                        gathered(C, "a", C.nnz), acd[0])
    retval = torch_scatter_reduce(0, mult, acd[0], C.nnz, aggr)
    return SparseTensor(C.indices, retval,
                        shape=_out_sparseshape(A, dim1, B, dim2) + tuple(retval.shape[1:]),
                        is_coalesced=True)


def _gather_rows_plan(plan: P.TriplePlan, which: str, n_src: int) -> P.TriplePlan:
    """out[t] = src[idx_which[t]] as a (cached) plan."""
    cache = plan.__dict__.setdefault("_row_gathers", {})
    gp = cache.get(which)
    if gp is None:
        gp = P.TriplePlan(plan.T, plan.T, n_src, 0, None, plan.idx[which], None)
        cache[which] = gp
    return gp
