"""``SparseTensor`` -- COO tensor with sparse leading dims and dense trailing dims, on B200.

API mirror of the reference ``pygho/backend/SpTensor.py`` (class at :200-527, hashing
helpers at :10-164, ``coalesce`` at :167-197): same constructor, properties and methods,
same assertion messages.  Every computation goes to the CUDA kernels of
``libpygho_b200.so``: hashing/sorting/unique are the plan kernels, reductions are the
deterministic segmented-reduce kernel (no atomics), and the CSR plans they need are
cached on the ``indices`` tensor, which all tuplewise results share.
"""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional, Tuple, Union

import torch
from torch import LongTensor, Tensor

from .. import plans as P
from ..ops import seg_gmr
from .utils import _flatten_dense, _seg_reduce_int


def indicehash(indice: LongTensor) -> LongTensor:
    """Lexicographic pack of (sparse_dim, nnz) indices into one int64 per column,
    ``63 // sparse_dim`` bits per coordinate (reference SpTensor.py:10-42)."""
    assert indice.ndim == 2
    return P.pack_keys(indice, check=__debug__)


def decodehash(indhash: LongTensor, sparse_dim: int) -> LongTensor:
    """Inverse of :func:`indicehash` (reference SpTensor.py:45-87)."""
    if sparse_dim > 1:
        assert indhash.ndim == 1, "indhash should of shape (nnz) "
    return P.unpack_keys(indhash, sparse_dim)


def _dims_list(dimsize) -> List[int]:
    return [int(v) for v in (dimsize.tolist() if isinstance(dimsize, Tensor) else dimsize)]


def indicehash_tight(indice: LongTensor, dimsize: LongTensor) -> LongTensor:
    """Row-major flattening by the given dim sizes (reference SpTensor.py:90-126)."""
    assert indice.ndim == 2, "indice shoule be of shape (sparse_dim, nnz) "
    dims = _dims_list(dimsize)
    assert len(dims) == indice.shape[0], "indice dim and dim size not match"
    total = 1
    for s in dims:
        total *= s
    assert total < (1 << 62), "total size exceeds the range that torch.long can express"
    return P.pack_tight(indice, dims, check=__debug__)


def decodehash_tight(indhash: LongTensor, dimsize: LongTensor) -> LongTensor:
    """Inverse of :func:`indicehash_tight` (reference SpTensor.py:129-164)."""
    assert indhash.ndim == 1, "indhash should of shape (nnz) "
    return P.unpack_tight(indhash, _dims_list(dimsize))


def _coalesce_plan(indices: LongTensor):
    """sorted unique coordinates + the plan that merges duplicate columns."""
    key = P.pack_keys(indices, check=__debug__)
    sd = indices.shape[0]
    ks, perm = P.sort_keys(key, P.hash_bits(sd) * sd if sd > 1 else 63)
    ukey, seg, count = P.unique_sorted(ks)
    plan = P.TriplePlan(indices.shape[1], count, indices.shape[1], 0, seg, perm, None,
                        sorted_by="a")
    return P.unpack_keys(ukey, sd), plan


def coalesce(edge_index: LongTensor, edge_attr: Optional[Tensor] = None,
             reduce: str = "sum") -> Tuple[Tensor, Optional[Tensor]]:
    """Sort coordinates, merge duplicates with ``reduce`` (reference SpTensor.py:167-197).
    Differentiable w.r.t. floating ``edge_attr``."""
    out_ind, plan = _coalesce_plan(edge_index)
    if edge_attr is None:
        return out_ind, None
    return out_ind, _reduce_values(edge_attr, plan, reduce)


def _reduce_values(values: Tensor, plan: P.TriplePlan, reduce: str) -> Tensor:
    if values.dtype == torch.float32:
        flat, dshape = _flatten_dense(values)
        return seg_gmr(flat, None, plan, reduce).reshape((plan.n_out,) + dshape)
    return _seg_reduce_int(values, plan, reduce)


class SparseTensor:
    """Coalesced COO tensor: ``indices`` (sparse_dim, nnz) int64 sorted lexicographically,
    ``values`` (nnz, *denseshape) or None.  See the reference class docstring
    (SpTensor.py:200-239) for the semantics this mirrors."""

    def __init__(self, indices: LongTensor, values: Optional[Tensor] = None,
                 shape: Optional[List[int]] = None, is_coalesced: bool = False,
                 reduce: str = "sum"):
        assert indices.ndim == 2, "indice should of shape (#sparsedim, #nnz)"
        if values is not None:
            assert indices.shape[1] == values.shape[0], \
                "indices and values should have the same number of nnz"
        self._sd = indices.shape[0]
        if shape is None:
            lead = [int(v) + 1 for v in torch.max(indices, dim=1).values.tolist()]
            self._shape = tuple(lead) + (tuple(values.shape[1:]) if values is not None else ())
        else:
            self._shape = tuple(int(s) for s in shape)
            if values is not None:
                assert self.denseshape == tuple(values.shape[1:]), "shape, value not match"
        if not is_coalesced:
            indices, values = coalesce(indices, values, reduce)
        self._indices, self._values = indices, values

    # ---- container protocol ---------------------------------------------------------
    def is_coalesced(self) -> bool:
        return True

    def to(self, device, non_blocking: bool = False):
        self._indices = self._indices.to(device, non_blocking=non_blocking)
        if self._values is not None:
            self._values = self._values.to(device, non_blocking=non_blocking)
        return self

    @property
    def indices(self) -> LongTensor:
        return self._indices

    @property
    def values(self) -> Optional[Tensor]:
        return self._values

    @property
    def sparse_dim(self) -> int:
        return self._sd

    @property
    def nnz(self) -> int:
        return self._indices.shape[1]

    @property
    def shape(self) -> Tuple[int, ...]:
        return self._shape

    @property
    def sparseshape(self) -> Tuple[int, ...]:
        return self._shape[:self._sd]

    @property
    def denseshape(self) -> Tuple[int, ...]:
        return self._shape[self._sd:]

    def __repr__(self):
        return f"SparseTensor(shape={self.shape}, sparse_dim={self.sparse_dim}, nnz={self.nnz})"

    def _same_pattern(self, values: Tensor) -> "SparseTensor":
        return SparseTensor(self._indices, values, self.sparseshape + tuple(values.shape[1:]),
                            is_coalesced=True)

    def _check_dims(self, dims) -> List[int]:
        dims = [int(d) for d in dims]
        assert all(d < self._sd for d in dims), \
            "please use tuplewiseapply for operation on dense dims"
        assert all(d >= 0 for d in dims), "do not support negative dims"
        return dims

    # ---- cached plans keyed on the shared indices tensor ------------------------------
    def _key_plan(self, keep: Tuple[int, ...]) -> P.TriplePlan:
        """Plan scattering tuple t to the (flattened) coordinate of its ``keep`` dims."""
        cache = P._cache(self._indices)
        ck = ("pool", keep)
        plan = cache.get(ck)
        if plan is None:
            if len(keep) == 1:
                key, n_rows = self._indices[keep[0]], self._shape[keep[0]]
            else:
                dims = [self._shape[i] for i in keep]
                key, n_rows = P.pack_tight(self._indices, dims, rows=keep), 1
                for s in dims:
                    n_rows *= s
            # the leading dims of a coalesced tensor are non-decreasing
            lead = keep == tuple(range(len(keep)))
            plan = P.TriplePlan(self.nnz, n_rows, self.nnz, 0, P.to_i32(key), None, None,
                                "a" if lead else "")
            cache[ck] = plan
        return plan

    def _pair_plan(self):
        """Index arrays of :class:`pygho_b200.ops.GatherProduct` for sparse dims (0, 1): the int32
        coordinates, and for each of the two dims the CSR over its coordinate together with the
        tuple ids and the OTHER coordinate in that order.  Cached on the indices."""
        cache = P._cache(self._indices)
        hit = cache.get("pairplan")
        if hit is None:
            p0, p1 = self._key_plan((0,)), self._key_plan((1,))
            i32, j32 = p0.idx["a"], p1.idx["a"]
            g0, g1 = p0.group("a"), p1.group("a")       # first = tuple ids in that order (or None)
            j_i = j32 if g0.first is None else P.gather_i32(j32, g0.first)
            i_j = i32 if g1.first is None else P.gather_i32(i32, g1.first)
            hit = (i32, j32, g0.rowptr, g0.first, j_i, g1.rowptr, g1.first, i_j)
            cache["pairplan"] = hit
        return hit

    def gather_product(self, X0: Tensor, X1: Tensor) -> "SparseTensor":
        """values[t] = X0[indices[0, t]] * X1[indices[1, t]] * values[t] -- what the reference
        models write as ``X.tuplewiseapply(lambda val: X0[X.indices[0]] * X1[X.indices[1]] * val)``
        (example/zinc.py:270-276), without materialising the two gathered factors (extension:
        not in the reference API; 2-D float32 dense features on CUDA, else the generic path)."""
        v = self._values
        if (self._sd >= 2 and v is not None and v.ndim == 2 and X0.ndim == 2 and X1.ndim == 2
                and v.is_cuda and v.dtype == X0.dtype == X1.dtype == torch.float32
                and X0.shape[1] == X1.shape[1] == v.shape[1] and v.shape[1] % 4 == 0
                and X0.shape[0] == self._shape[0] and X1.shape[0] == self._shape[1] and self.nnz):
            from ..ops import GatherProduct
            return self._same_pattern(GatherProduct.apply(X0.contiguous(), X1.contiguous(),
                                                          v.contiguous(), self._pair_plan()))
        root = self.unpooling_fromdense1dim(0, X0).values
        node = self.unpooling_fromdense1dim(1, X1).values
        return self._same_pattern(root * node * v)

    def _sparse_pool_plan(self, keep: Tuple[int, ...]):
        cache = P._cache(self._indices)
        ck = ("spool", keep)
        hit = cache.get(ck)
        if hit is None:
            hit = _coalesce_plan(self._indices[list(keep)])
            cache[ck] = hit
        return hit

    # ---- reductions over sparse dims (reference SpTensor.py:368-445) -------------------
    def _reduce_to_sparse(self, dims: Iterable[int], reduce: str) -> "SparseTensor":
        dims = self._check_dims(dims)
        keep = tuple(i for i in range(self._sd) if i not in dims)
        out_ind, plan = self._sparse_pool_plan(keep)
        vals = _reduce_values(self._values, plan, reduce)
        return SparseTensor(out_ind, vals, tuple(self._shape[i] for i in keep) + self.denseshape,
                            is_coalesced=True)

    def _reduce_to_dense(self, dims: Iterable[int], reduce: str) -> Tensor:
        dims = self._check_dims(dims)
        keep = tuple(i for i in range(self._sd) if i not in dims)
        plan = self._key_plan(keep)
        out = _reduce_values(self._values, plan, reduce)
        if len(keep) > 1:
            out = out.reshape(tuple(self._shape[i] for i in keep) + tuple(out.shape[1:]))
        return out

    def _pool(self, dims, return_sparse: bool, reduce: str):
        if isinstance(dims, int):
            dims = [dims]
        if dims is None:
            # the reference raises here (torch.sum(..., dims=0), SpTensor.py:417); give the
            # obviously intended result instead: reduce over all tuples
            flat = self._values
            if reduce == "sum":
                return flat.sum(0)
            if reduce == "mean":
                return flat.mean(0)
            return flat.amax(0) if reduce == "max" else flat.amin(0)
        if return_sparse:
            return self._reduce_to_sparse(dims, reduce)
        return self._reduce_to_dense(dims, reduce)

    def sum(self, dims: Union[int, Optional[Iterable[int]]], return_sparse: bool = False):
        return self._pool(dims, return_sparse, "sum")

    def max(self, dims: Union[int, Optional[Iterable[int]]], return_sparse: bool = False):
        return self._pool(dims, return_sparse, "max")

    def mean(self, dims: Union[int, Optional[Iterable[int]]], return_sparse: bool = False):
        return self._pool(dims, return_sparse, "mean")

    def min(self, dims: Union[int, Optional[Iterable[int]]], return_sparse: bool = False):
        return self._pool(dims, return_sparse, "min")

    # ---- diagonal (reference SpTensor.py:304-366) --------------------------------------
    def _diag_lookup(self, dims: List[int]):
        """b2a (int32, -1 = absent) of the diagonal coordinates (i, i, ...) of ``dims``."""
        cache = P._cache(self._indices)
        ck = ("diag", tuple(dims))
        hit = cache.get(ck)
        if hit is None:
            n = self._shape[dims[0]]
            ar = torch.arange(n, device=self._indices.device)
            diag_key = P.pack_keys(ar.unsqueeze(0).expand(len(dims), -1).contiguous())
            self_key = P.pack_keys(self._indices, rows=dims)
            hit = (P.lookup_sorted(self_key, diag_key), n)
            cache[ck] = hit
        return hit

    def _diag_to_dense(self, dims: List[int]) -> Tensor:
        if len(dims) != self._sd:
            return self._diag_subset_to_dense(dims)
        pos, n = self._diag_lookup(dims)
        plan = _gather_plan(pos, self.nnz)
        flat, dshape = _flatten_dense(self._values)
        return seg_gmr(flat, None, plan, "sum").reshape((n,) + dshape)

    def _diag_subset_to_dense(self, dims: List[int]) -> Tensor:
        """Diagonal over a SUBSET of the sparse dims (reference SpTensor.py:336-352): the result
        keeps ``dims[0]`` and every non-diagonal dim, ``ret[i, k, ...] = X[i, i, k, ...]``.
        The reference looks each diagonal coordinate up with ONE searchsorted on the hash of the
        diagonal dims only, so it finds a single tuple per diagonal value and fills one ``k``;
        this implements the documented meaning (all tuples whose diagonal dims coincide).  Rarely
        used (3-D tensors only), plain torch indexing: gradients flow through ``index_put``."""
        idx = [i for i in range(self._sd) if i not in dims[1:]]
        on_diag = (self._indices[dims] == self._indices[dims[0]].unsqueeze(0)).all(dim=0)
        out = torch.zeros(tuple(self._shape[i] for i in idx) + self.denseshape,
                          dtype=self._values.dtype, device=self._values.device)
        coords = tuple(self._indices[i][on_diag] for i in idx)
        return out.index_put(coords, self._values[on_diag])

    def _diag_to_sparse(self, dims: List[int]) -> "SparseTensor":
        raise NotImplementedError(
            "diag(return_sparse=True) raises in the reference too (torch.all(dims=), "
            "SpTensor.py:312-313); not part of the hot path")

    def diag(self, dims: Optional[Iterable[int]], return_sparse: bool = False):
        if isinstance(dims, int):
            raise NotImplementedError
        if dims is None:
            dims = list(range(self._sd))
        dims = self._check_dims(sorted(set(dims)))
        return self._diag_to_sparse(dims) if return_sparse else self._diag_to_dense(dims)

    # ---- unpooling (reference SpTensor.py:447-476) -------------------------------------
    def unpooling(self, dims: Union[int, Iterable[int]], tarX: "SparseTensor"):
        """Broadcast ``self`` along the sparse dims ``dims`` of ``tarX``'s pattern; target
        tuples without a source tuple get 0."""
        if isinstance(dims, int):
            dims = [dims]
        dims = list(dims)
        keep = tuple(i for i in range(tarX.sparse_dim) if i not in dims)
        cache = P._cache(tarX.indices)
        # the cache lives on the (long-lived) target pattern; the entry keeps the source index
        # tensor it was built for, so a recycled id() can never return another pattern's plan
        ck = ("unpool_from", id(self._indices), keep)
        hit = cache.get(ck)
        plan = hit[0] if hit is not None and hit[1] is self._indices else None
        if plan is None:
            self_key = P.pack_keys(self._indices, check=__debug__)
            if __debug__:
                assert P.is_sorted(self_key, strict=True), "self is not coalesced"
            pos = P.lookup_sorted(self_key, P.pack_keys(tarX.indices, rows=keep))
            plan = _gather_plan(pos, self.nnz)
            cache[ck] = (plan, self._indices)
        flat, dshape = _flatten_dense(self._values)
        vals = seg_gmr(flat, None, plan, "sum").reshape((tarX.nnz,) + dshape)
        return tarX.tuplewiseapply(lambda _x: vals)

    def unpooling_fromdense1dim(self, dims: int, X: Tensor) -> "SparseTensor":
        """values[t] = X[indices[dims, t]] (dense node features onto the tuples)."""
        assert dims < self._sd, "only unpooling sparse dims"
        assert X.shape[0] == self._shape[dims], "shape not match"
        plan = self._key_plan((int(dims),)).transposed()
        flat, dshape = _flatten_dense(X)
        return self._same_pattern(seg_gmr(flat, None, plan, "sum").reshape((self.nnz,) + dshape))

    # ---- conversions -----------------------------------------------------------------
    @classmethod
    def from_torch_sparse_coo(cls, A: Tensor) -> "SparseTensor":
        assert A.is_sparse, \
            "from_torch_sparse_coo converts a torch.sparse_coo_tensor to SparseTensor"
        return cls(A._indices(), A._values(), A.shape, A.is_coalesced())

    def to_torch_sparse_coo(self) -> Tensor:
        ret = torch.sparse_coo_tensor(self._indices, self._values, size=self._shape)
        return ret._coalesced_(True)

    # ---- tuplewise ops (reference SpTensor.py:491-524) ---------------------------------
    def tuplewiseapply(self, func: Callable[[Tensor], Tensor]) -> "SparseTensor":
        return self._same_pattern(func(self._values))

    def diagonalapply(self, func: Callable[[Tensor, LongTensor], Tensor]) -> "SparseTensor":
        assert self._sd == 2, "only implemented for 2D"
        flag = (self._indices[0] == self._indices[1]).to(torch.long)
        return self._same_pattern(func(self._values, flag))

    def add(self, tarX: "SparseTensor", samesparse: bool) -> "SparseTensor":
        if samesparse:
            return self._same_pattern(self._values + tarX.values)
        return SparseTensor(torch.concat((self._indices, tarX.indices), dim=1),
                            torch.concat((self._values, tarX.values), dim=0), self._shape, False)

    def catvalue(self, tarXs: Union["SparseTensor", Iterable["SparseTensor"]],
                 samesparse: bool) -> "SparseTensor":
        if isinstance(tarXs, SparseTensor):
            tarXs = [tarXs]
        assert samesparse == True, "must have the same sparcity to concat value"  # noqa: E712
        return self._same_pattern(torch.concat([self._values] + [t.values for t in tarXs], dim=-1))


def _gather_plan(pos: Tensor, n_src: int) -> P.TriplePlan:
    """Plan for ``out[t] = src[pos[t]]`` with ``pos[t] == -1`` meaning "no source" (-> 0).

    Built as a segmented reduce whose row t has one entry when pos[t] >= 0 and none
    otherwise, so missing rows are exactly 0 and receive no gradient."""
    cache = P._cache(pos)
    plan = cache.get("gather")
    if plan is None:
        n = pos.numel()
        dev = pos.device
        iota = torch.arange(n, dtype=torch.int32, device=dev)
        oa, oc, od = (torch.empty(n, dtype=torch.int32, device=dev) for _ in range(3))
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        if n:
            ws = P._ws(P.size_query("pgh_compact_ws_bytes", n), dev)
            # keep the rows with pos >= 0: a = pos (>= 0 filter), c = row id t
            P._launch("pgh_compact_triples", P.ptr(pos), None, P.ptr(iota), P.ptr(iota), n,
                      P.ptr(oa), P.ptr(oc), P.ptr(od), P.ptr(cnt), P.ptr(ws), ws.numel(),
                      P.stream_ptr(dev))
        T = int(cnt.item()) if n else 0
        # triples: output row oc[t] <- source row oa[t]; oc is increasing
        plan = P.TriplePlan(T, n, n_src, 0, oc[:T].contiguous(), oa[:T].contiguous(), None,
                            sorted_by="a")
        cache["gather"] = plan
    return plan
