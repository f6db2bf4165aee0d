"""CPU oracle for the ``pygho.backend`` tensor-operator path  --  TEST INFRASTRUCTURE ONLY.

This file restates, in numpy, what the reference (GraphPKU/PygHO, pure Python on
torch) computes on the hot path.  It is the checker the CUDA kernels are compared
with; it is never imported by the product package (``pygho_b200``), only by
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py``.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the real reference
from ``/root/reference`` (possible because it is Python), runs it on seeded inputs and
stores inputs + outputs in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
checks every function below against those vectors and against the reference's own
known-answer tests (``tests/test_backend_sparse.py:35-99``,
``tests/test_backend_masked.py:45-59``).

Conventions: indices are int64 arrays of shape (sparse_dim, nnz); values are
(nnz, *dense).  Floating reductions are accumulated in float64 and rounded once to
the input dtype, so the oracle is at least as accurate as the reference's fp32
scatter; comparisons use 1e-5 relative (BASELINE.json north_star).

Reference citations are relative to ``/root/reference/pygho``.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np

I64 = np.int64


# --------------------------------------------------------------------------- hashing
def hash_bits(sparse_dim: int) -> int:
    """Bits per coordinate of the packed key: ``63 // sparse_dim`` (backend/SpTensor.py:34)."""
    return 63 // sparse_dim


def indicehash(ind: np.ndarray) -> np.ndarray:
    """Order-preserving pack of the rows of ``ind`` into one int64 key per column.

    backend/SpTensor.py:10-42 -- row 0 occupies the most significant field."""
    ind = np.asarray(ind, dtype=I64)
    assert ind.ndim == 2
    sd = ind.shape[0]
    if sd == 1:
        return ind[0].copy()
    bits = hash_bits(sd)
    if ind.size:
        assert ind.min() >= 0, "indice cannot be negative"
        assert int(ind.max()) < (1 << bits), "too large indice, hash is not injective"
    key = np.zeros(ind.shape[1], dtype=I64)
    for r in range(sd):
        key = (key << bits) | ind[r]
    return key


def decodehash(key: np.ndarray, sparse_dim: int) -> np.ndarray:
    """Inverse of :func:`indicehash` (backend/SpTensor.py:45-87)."""
    key = np.asarray(key, dtype=I64)
    if sparse_dim == 1:
        return key[None, :].copy()
    bits = hash_bits(sparse_dim)
    field = (1 << bits) - 1
    rows = [(key >> (bits * (sparse_dim - 1 - r))) & field for r in range(sparse_dim)]
    return np.stack(rows).astype(I64)


def indicehash_tight(ind: np.ndarray, dimsize: Sequence[int]) -> np.ndarray:
    """Mixed-radix (row-major) flattening (backend/SpTensor.py:90-126)."""
    ind = np.asarray(ind, dtype=I64)
    dimsize = [int(s) for s in dimsize]
    assert ind.shape[0] == len(dimsize)
    key = np.zeros(ind.shape[1], dtype=I64)
    for r, s in enumerate(dimsize):
        if ind.shape[1]:
            assert 0 <= ind[r].min() and ind[r].max() < s, "indice exceeds dimsize"
        key = key * s + ind[r]
    return key


def decodehash_tight(key: np.ndarray, dimsize: Sequence[int]) -> np.ndarray:
    """Inverse of :func:`indicehash_tight` (backend/SpTensor.py:129-164)."""
    key = np.asarray(key, dtype=I64).copy()
    rows = []
    for s in reversed([int(s) for s in dimsize]):
        rows.append(key % s)
        key //= s
    return np.stack(rows[::-1]).astype(I64)


# ----------------------------------------------------------------- segmented reduce
_AGGR = ("sum", "mean", "max", "min")


def scatter_reduce(src: np.ndarray, ind: np.ndarray, dim_size: int, aggr: str) -> np.ndarray:
    """``torch_scatter_reduce(0, src, ind, dim_size, aggr)`` (backend/utils.py:6-56).

    Rows that receive nothing are 0 for every aggr (zero init + include_self=False);
    integer ``mean`` floors like torch's integer division."""
    assert aggr in _AGGR, aggr
    src = np.asarray(src)
    ind = np.asarray(ind, dtype=I64)
    assert ind.ndim == 1 and src.shape[0] == ind.shape[0]
    out_shape = (int(dim_size),) + src.shape[1:]
    if ind.size == 0:
        return np.zeros(out_shape, dtype=src.dtype)
    assert ind.min() >= 0 and ind.max() < dim_size
    order = np.argsort(ind, kind="stable")
    keys = ind[order]
    starts = np.flatnonzero(np.r_[True, keys[1:] != keys[:-1]])
    rows = keys[starts]
    is_float = np.issubdtype(src.dtype, np.floating)
    work = src[order].astype(np.float64 if is_float else I64)
    if aggr in ("sum", "mean"):
        red = np.add.reduceat(work, starts, axis=0)
        if aggr == "mean":
            cnt = np.diff(np.r_[starts, keys.shape[0]]).reshape((-1,) + (1,) * (src.ndim - 1))
            red = red / cnt if is_float else np.floor_divide(red, cnt)
    elif aggr == "max":
        red = np.maximum.reduceat(work, starts, axis=0)
    else:
        red = np.minimum.reduceat(work, starts, axis=0)
    out = np.zeros(out_shape, dtype=src.dtype)
    out[rows] = red.astype(src.dtype)
    return out


def coalesce(ind: np.ndarray, val: Optional[np.ndarray], reduce: str = "sum"):
    """Sort by packed key, merge duplicates with ``reduce`` (backend/SpTensor.py:167-197)."""
    ind = np.asarray(ind, dtype=I64)
    key = indicehash(ind)
    ukey, inverse = np.unique(key, return_inverse=True)
    out_ind = decodehash(ukey, ind.shape[0])
    if val is None:
        return out_ind, None
    return out_ind, scatter_reduce(np.asarray(val), inverse.reshape(-1), ukey.shape[0], reduce)


# ------------------------------------------------------------------------ index plans
def ptr2batch(ptr: np.ndarray, dim_size: Optional[int] = None) -> np.ndarray:
    """``batch[ptr[i]:ptr[i+1]] = i`` (backend/Spspmm.py:9-31)."""
    ptr = np.asarray(ptr, dtype=I64)
    assert ptr.ndim == 1 and ptr[0] == 0 and np.all(np.diff(ptr) >= 0)
    if dim_size is not None:
        assert ptr[-1] == dim_size
    return np.repeat(np.arange(ptr.shape[0] - 1, dtype=I64), np.diff(ptr))


def deg2batch(deg: np.ndarray, dim_size: Optional[int] = None) -> np.ndarray:
    """backend/Spspmm.py:34-54."""
    deg = np.asarray(deg, dtype=I64)
    assert deg.ndim == 1 and np.all(deg >= 0)
    return np.repeat(np.arange(deg.shape[0], dtype=I64), deg)


def canonical_plan(plan: np.ndarray) -> np.ndarray:
    """Sort the columns of a (3, T) plan by (row0, row1, row2).

    The reference's final ``argsort`` (backend/Spspmm.py:142) is not stable, so the
    order inside one output segment is unspecified; plans are compared in this
    canonical order."""
    plan = np.asarray(plan, dtype=I64)
    order = np.lexsort((plan[2], plan[1], plan[0]))
    return plan[:, order]


def spspmm_ind(ind1: np.ndarray, dim1: int, ind2: np.ndarray, dim2: int):
    """Contraction plan of two sparse index sets (backend/Spspmm.py:57-143).

    For every nonzero ``p`` of ``ind1`` and every nonzero ``q`` of ``ind2`` with
    ``ind1[dim1, p] == ind2[dim2, q]`` there is one triple; the output coordinate is
    the remaining dims of ``ind1`` followed by the remaining dims of ``ind2``.
    Returns ``tarind`` (sorted unique output coordinates) and ``bcd`` (3, T0) in
    canonical order: ``b`` = column of ``tarind``, ``c`` = p, ``d`` = q."""
    ind1 = np.asarray(ind1, dtype=I64)
    ind2 = np.asarray(ind2, dtype=I64)
    sd1, sd2 = ind1.shape[0], ind2.shape[0]
    assert 0 <= dim1 < sd1 and 0 <= dim2 < sd2
    k1, k2 = ind1[dim1], ind2[dim2]
    perm = np.argsort(k2, kind="stable")
    k2s = k2[perm]
    lo = np.searchsorted(k2s, k1, side="left")
    hi = np.searchsorted(k2s, k1, side="right")
    cnt = hi - lo
    c = np.repeat(np.arange(ind1.shape[1], dtype=I64), cnt)
    first = np.cumsum(cnt) - cnt
    within = np.arange(c.shape[0], dtype=I64) - np.repeat(first, cnt)
    d = perm[np.repeat(lo, cnt) + within].astype(I64)
    rest1 = np.delete(ind1, dim1, axis=0)[:, c]
    rest2 = np.delete(ind2, dim2, axis=0)[:, d]
    out_sd = sd1 + sd2 - 2
    key = indicehash(np.concatenate([rest1, rest2], axis=0))
    ukey, b = np.unique(key, return_inverse=True)
    tarind = decodehash(ukey, out_sd)
    return tarind, canonical_plan(np.stack([b.reshape(-1).astype(I64), c, d]))


def spsphadamard_ind(tar_ind: np.ndarray, ind: np.ndarray) -> np.ndarray:
    """Column of ``tar_ind`` equal to each column of ``ind``, -1 if absent
    (backend/Spspmm.py:146-183).  ``tar_ind`` must be sorted and duplicate free."""
    tkey = indicehash(np.asarray(tar_ind, dtype=I64))
    assert np.all(np.diff(tkey) > 0), "tar_ind should be sorted and coalesce"
    key = indicehash(np.asarray(ind, dtype=I64))
    if tkey.size == 0:
        return np.full(key.shape, -1, dtype=I64)
    pos = np.clip(np.searchsorted(tkey, key, side="right") - 1, 0, None)
    return np.where(tkey[pos] == key, pos, -1).astype(I64)


def filterind(tar_ind: np.ndarray, ind: np.ndarray, bcd: np.ndarray) -> np.ndarray:
    """Keep the triples whose output coordinate is in ``tar_ind`` and renumber the
    output row (backend/Spspmm.py:186-222).  Canonical order."""
    b2a = spsphadamard_ind(tar_ind, ind)
    bcd = np.asarray(bcd, dtype=I64)
    a = b2a[bcd[0]]
    keep = a >= 0
    return canonical_plan(np.stack([a[keep], bcd[1][keep], bcd[2][keep]]))


# -------------------------------------------------------------------------- value ops
def _bmul(x: Optional[np.ndarray], y: Optional[np.ndarray]) -> np.ndarray:
    if x is None:
        return y
    if y is None:
        return x
    return x * y


def spspmm(a_val: Optional[np.ndarray], b_val: Optional[np.ndarray], acd: np.ndarray,
           n_out: int, aggr: str = "sum") -> np.ndarray:
    """Values of ``spspmm(A, dim1, B, dim2, aggr, acd=acd, tar_ind=...)``
    (backend/Spspmm.py:307-321): ``out[a] = aggr_t A[c_t] * B[d_t]``; a ``None``
    operand counts as 1."""
    acd = np.asarray(acd, dtype=I64)
    msg = _bmul(None if a_val is None else a_val[acd[1]],
                None if b_val is None else b_val[acd[2]])
    return scatter_reduce(msg, acd[0], n_out, aggr)


def spsphadamard(ind1, val1, ind2, val2):
    """Elementwise product on the common pattern (backend/Spspmm.py:225-267)."""
    b2a = spsphadamard_ind(ind1, ind2)
    m = b2a >= 0
    if val1 is None:
        val = val2[m]
    elif val2 is None:
        val = val1[b2a[m]]
    else:
        val = val1[b2a[m]] * val2[m]
    return np.asarray(ind2)[:, m], val


def spmm(a_ind: np.ndarray, a_val: Optional[np.ndarray], a_shape: Sequence[int], dim1: int,
         x: np.ndarray, aggr: str = "sum") -> np.ndarray:
    """2-D sparse times dense (backend/Spmm.py:6-44); ``dim1`` is the contracted
    dim of A."""
    a_ind = np.asarray(a_ind, dtype=I64)
    src, tar = (a_ind[0], a_ind[1]) if dim1 == 0 else (a_ind[1], a_ind[0])
    n_tar = a_shape[1] if dim1 == 0 else a_shape[0]
    msg = x[src] if a_val is None else a_val * x[src]
    return scatter_reduce(msg, tar, n_tar, aggr)


def sp_pool_dense(ind: np.ndarray, val: np.ndarray, shape: Sequence[int],
                  dims: Sequence[int], aggr: str) -> np.ndarray:
    """``SparseTensor.sum/mean/max(dims)`` -> dense (backend/SpTensor.py:382-409)."""
    ind = np.asarray(ind, dtype=I64)
    sd = ind.shape[0]
    keep = [i for i in range(sd) if i not in list(dims)]
    if len(keep) == 1:
        return scatter_reduce(val, ind[keep[0]], shape[keep[0]], aggr)
    kshape = [int(shape[i]) for i in keep]
    flat = scatter_reduce(val, indicehash_tight(ind[keep], kshape), int(np.prod(kshape)), aggr)
    return flat.reshape(tuple(kshape) + val.shape[1:])


def sp_pool_sparse(ind: np.ndarray, val: np.ndarray, dims: Sequence[int], aggr: str):
    """``SparseTensor.sum/mean/max(dims, return_sparse=True)``
    (backend/SpTensor.py:368-380 -> coalesce)."""
    ind = np.asarray(ind, dtype=I64)
    keep = [i for i in range(ind.shape[0]) if i not in list(dims)]
    return coalesce(ind[keep], val, aggr)


def sp_unpool_dense(ind: np.ndarray, dim: int, x: np.ndarray) -> np.ndarray:
    """``unpooling_fromdense1dim`` (backend/SpTensor.py:470-476)."""
    return x[np.asarray(ind, dtype=I64)[dim]]


def sp_unpool_sparse(src_ind, src_val, tar_ind, dims: Sequence[int]) -> np.ndarray:
    """``SparseTensor.unpooling`` (backend/SpTensor.py:447-468): rows of the target
    pattern take the value of the matching (reduced-dims removed) source tuple, 0 when
    there is none."""
    tar_ind = np.asarray(tar_ind, dtype=I64)
    keep = [i for i in range(tar_ind.shape[0]) if i not in list(dims)]
    b2a = spsphadamard_ind(src_ind, tar_ind[keep])
    out = np.zeros((tar_ind.shape[1],) + src_val.shape[1:], dtype=src_val.dtype)
    m = b2a >= 0
    out[m] = src_val[b2a[m]]
    return out


def sp_diag_dense(ind, val, shape, dims: Sequence[int]) -> np.ndarray:
    """``SparseTensor.diag(dims)`` with all sparse dims listed
    (backend/SpTensor.py:322-333): value of tuple (i, i, ...) or 0."""
    ind = np.asarray(ind, dtype=I64)
    assert len(dims) == ind.shape[0]
    n = int(shape[dims[0]])
    diag = np.tile(np.arange(n, dtype=I64), (len(dims), 1))
    b2a = spsphadamard_ind(ind, diag)
    out = np.zeros((n,) + val.shape[1:], dtype=val.dtype)
    m = b2a >= 0
    out[m] = val[b2a[m]]
    return out


# ------------------------------------------------------------------------ masked path
def ma_fill(data: np.ndarray, mask: np.ndarray, value: float) -> np.ndarray:
    """``MaskedTensor.fill_masked`` with the *intended* semantics
    (backend/MaTensor.py:113-128; the reference skips the fill when the requested
    value equals the recorded padvalue)."""
    m = mask.reshape(mask.shape + (1,) * (data.ndim - mask.ndim))
    return np.where(m, data, np.asarray(value, dtype=data.dtype))


def ma_pool(data: np.ndarray, mask: np.ndarray, dims: Sequence[int], aggr: str):
    """``MaskedTensor.sum/mean/max/min`` (backend/MaTensor.py:175-206).  Returns
    (data, mask).  ``min`` is a true minimum (the reference calls amax, Q2)."""
    dims = tuple(int(d) for d in dims)
    m = mask.reshape(mask.shape + (1,) * (data.ndim - mask.ndim))
    omask = mask.any(axis=dims)
    d64 = data.astype(np.float64)
    if aggr in ("sum", "mean"):
        out = np.where(m, d64, 0.0).sum(axis=dims)
        if aggr == "mean":
            cnt = np.maximum(mask.sum(axis=dims), 1)
            out = out / cnt.reshape(cnt.shape + (1,) * (out.ndim - cnt.ndim))
    elif aggr == "max":
        out = np.where(m, d64, -np.inf).max(axis=dims)
        out = np.where(np.isinf(out), 0.0, out)
    elif aggr == "min":
        out = np.where(m, d64, np.inf).min(axis=dims)
        out = np.where(np.isinf(out), 0.0, out)
    else:
        raise ValueError(aggr)
    return out.astype(data.dtype), omask


def mamamm(a: np.ndarray, a_mask: np.ndarray, dim1: int, b: np.ndarray, b_mask: np.ndarray,
           dim2: int, out_mask: np.ndarray) -> np.ndarray:
    """``mamamm(A, dim1, B, dim2, mask, broadcast_firstdim=True)`` for
    (batch, n, n, *dense) operands (backend/Mamamm.py:7-64): contract masked dim
    ``dim1`` of A with ``dim2`` of B, batched over dim 0 and elementwise over the dense
    dims; masked operand entries count as 0 and the result is zeroed outside
    ``out_mask`` (intended semantics, Q1)."""
    assert a_mask.ndim == 3 and b_mask.ndim == 3 and dim1 in (1, 2) and dim2 in (1, 2)
    a0 = ma_fill(a, a_mask, 0.0).astype(np.float64)
    b0 = ma_fill(b, b_mask, 0.0).astype(np.float64)
    sa = "bji" if dim1 == 1 else "bij"
    sb = "bjk" if dim2 == 1 else "bkj"
    out = np.einsum(f"{sa}...,{sb}...->bik...", a0, b0)
    return ma_fill(out.astype(a.dtype), out_mask, 0.0)


def ma_unpool(data: np.ndarray, dims: Sequence[int], tar_shape: Sequence[int]) -> np.ndarray:
    """``MaskedTensor.unpooling`` (backend/MaTensor.py:225-234): insert and broadcast."""
    out = data
    for d in sorted(dims):
        out = np.expand_dims(out, d)
    shape = list(out.shape)
    for d in dims:
        shape[d] = int(tar_shape[d])
    return np.broadcast_to(out, shape)


def ma_diag(data: np.ndarray, mask: np.ndarray, dims: Sequence[int]):
    """``MaskedTensor.diag`` for two dims (backend/MaTensor.py:208-223)."""
    d0, d1 = sorted(dims)
    dd = np.moveaxis(np.diagonal(data, axis1=d0, axis2=d1), -1, d0)
    mm = np.moveaxis(np.diagonal(mask, axis1=d0, axis2=d1), -1, d0)
    return dd, mm


def spmamm(a_ind: np.ndarray, a_val: Optional[np.ndarray], a_shape: Sequence[int], dim1: int,
           b_data: np.ndarray, b_mask: np.ndarray, dim2: int, out_mask: Optional[np.ndarray],
           aggr: str = "sum"):
    """backend/Spmamm.py:12-67 with its intended semantics: ``dim1`` of the sparse (b, n, m)
    operand is contracted with masked dim ``dim2`` of B; masked-out positions of B do not
    contribute (the fill of :60 applied), A's values broadcast over B's other masked dims,
    empty reductions are 0 (filterinf, :65-66), the result is zero outside ``out_mask``
    (default B's mask).  Returns (data, mask)."""
    md = b_mask.ndim
    if dim1 == 1:
        n, cidx, tidx = a_shape[2], a_ind[1], a_ind[2]          # :44-47
    elif dim1 == 2:
        n, cidx, tidx = a_shape[1], a_ind[2], a_ind[1]          # :48-51
    else:
        raise NotImplementedError
    b = a_shape[0]
    tb = np.moveaxis(b_data, dim2, 1)                           # :55
    tm = np.moveaxis(b_mask, dim2, 1)                           # :56
    others, dense = tb.shape[2:md], tb.shape[md:]
    fill = {"sum": 0.0, "max": -np.inf, "min": np.inf}[aggr]    # :8
    rows = tb[a_ind[0], cidx].astype(np.float64)                # (nnz, *others, *dense)
    if a_val is not None:
        rows = a_val.reshape((a_val.shape[0],) + (1,) * len(others) + dense) * rows   # :57-58
    valid = tm[a_ind[0], cidx]
    rows = np.where(valid.reshape(valid.shape + (1,) * len(dense)), rows, fill)       # :60-61
    out = np.full((b * n,) + others + dense, fill, dtype=np.float64)
    tar = n * a_ind[0] + tidx
    red = {"sum": np.add, "max": np.maximum, "min": np.minimum}[aggr]
    red.at(out, tar, rows)                                      # :62
    out = np.where(np.isinf(out), 0.0, out)                     # :65-66
    out = np.moveaxis(out.reshape((b, n) + others + dense), 1, dim2)
    mask = b_mask if out_mask is None else out_mask
    out = np.where(mask.reshape(mask.shape + (1,) * (out.ndim - mask.ndim)), out, 0.0)
    return out.astype(np.float32), mask
