"""Sparse (``SparseTensor``) graph operators -- thin ``nn.Module`` wrappers that pick dims
and precomputed plans and call the backend (reference ``pygho/honn/SpOperator.py``:
``parse_precomputekey`` :15, ``OpNodeMessagePassing`` :47, ``OpMessagePassing`` :88,
``Op2FWL`` :185, ``OpMessagePassingOnSubg2D`` :230, ``...OnSubg3D`` :280,
``...CrossSubg2D`` :330, ``OpDiag`` :375, ``OpDiag2D`` :406, ``OpPooling`` :427 and its
variants :470-545, ``OpUnpooling`` :548-601).  No arithmetic lives here."""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Optional, Union

from torch import Tensor
from torch.nn import Module

from ..backend.Spmm import spmm
from ..backend.Spspmm import spspmm, spspmpnn
from ..backend.SpTensor import SparseTensor

KEYSEP = "___"


def parse_precomputekey(model: Module) -> List[str]:
    """Sorted, de-duplicated plan keys (``op0___op1___dim1___op2___dim2``) that the
    message-passing operators inside ``model`` will look up in ``datadict``."""
    return sorted({m.precomputekey for m in model.modules() if isinstance(m, OpMessagePassing)})


class OpNodeMessagePassing(Module):
    """``X' = A X`` on node features: ``A`` (n, n) sparse, ``X`` (n, d) dense."""

    def __init__(self, aggr: str = "sum") -> None:
        super().__init__()
        self.aggr = aggr

    def forward(self, A: SparseTensor, X: Tensor, tarX: Optional[Tensor] = None) -> Tensor:
        assert A.sparse_dim == 2, "A is adjacency matrix of the whole graph of shape nxn"
        return spmm(A, 1, X, self.aggr)


class OpMessagePassing(Module):
    """Generalised sparse product ``op0 <- aggr(op1 (dim1) x op2 (dim2))`` restricted to the
    pattern of the target.  The plan is taken from ``datadict[f"{key}___acd"]`` (or
    ``___bcd`` / ``___tarind``) where ``key`` is :attr:`precomputekey`."""

    def __init__(self, op0: str = "X", op1: str = "X", dim1: int = 1, op2: str = "A",
                 dim2: int = 0, aggr: str = "sum", message_func: Optional[Callable] = None) -> None:
        super().__init__()
        self.dim1, self.dim2, self.aggr = dim1, dim2, aggr
        self.precomputekey = KEYSEP.join((op0, op1, str(dim1), op2, str(dim2)))
        self.message_func = message_func
        self.use_mpnn = message_func is not None

    def _plan(self, datadict: Dict, suffix: str):
        return datadict.get(f"{self.precomputekey}{KEYSEP}{suffix}", None)

    def forward(self, A: SparseTensor, B: SparseTensor, datadict: Dict,
                tarX: Optional[SparseTensor] = None) -> SparseTensor:
        if self.use_mpnn:
            assert tarX is not None, \
                "target representation is a must when message func is not None"
            return spspmpnn(A, self.dim1, B, self.dim2, tarX, self._plan(datadict, "acd"),
                            self.message_func, self.aggr)
        tar_ind = self._plan(datadict, "tarind") if tarX is None else tarX.indices
        return spspmm(A, self.dim1, B, self.dim2, self.aggr, acd=self._plan(datadict, "acd"),
                      bcd=self._plan(datadict, "bcd"), tar_ind=tar_ind)


class _AdjTupleMessagePassing(OpMessagePassing):
    """Shared forward of the adjacency/tuple variants: subclasses state the expected
    sparse dims of ``A`` and ``X`` and whether the tuple operand comes first."""
    _A_NDIM, _X_NDIM, _TUPLE_FIRST = 2, 2, True
    _A_MSG, _X_MSG = "A should be nxn adjacency matrix ", "X should be 2d representations"

    def forward(self, A: SparseTensor, X: SparseTensor, datadict: Dict,
                tarX: Optional[SparseTensor] = None) -> SparseTensor:
        assert A.sparse_dim == self._A_NDIM, self._A_MSG
        assert X.sparse_dim == self._X_NDIM, self._X_MSG
        first, second = (X, A) if self._TUPLE_FIRST else (A, X)
        return OpMessagePassing.forward(self, first, second, datadict, tarX)


class Op2FWL(OpMessagePassing):
    """2-FWL: ``X[i, j] <- aggr_k X1[i, k] X2[k, j]``."""

    def __init__(self, aggr: str = "sum", optuplefeat: str = "X") -> None:
        super().__init__(optuplefeat, optuplefeat, 1, optuplefeat, 0, aggr)

    def forward(self, X1: SparseTensor, X2: SparseTensor, datadict: Dict,
                tarX: Optional[SparseTensor] = None) -> SparseTensor:
        assert X1.sparse_dim == 2, "X1 should be 2d representations "
        assert X2.sparse_dim == 2, "X2 should be 2d representations"
        return super().forward(X1, X2, datadict, tarX)


class OpMessagePassingOnSubg2D(_AdjTupleMessagePassing):
    """Message passing inside every subgraph: ``X[i, j] <- aggr_k X[i, k] A[k, j]``."""

    def __init__(self, aggr: str = "sum", optuplefeat: str = "X", opadj: str = "A",
                 message_func: Optional[Callable] = None) -> None:
        super().__init__(optuplefeat, optuplefeat, 1, opadj, 0, aggr, message_func)


class OpMessagePassingOnSubg3D(_AdjTupleMessagePassing):
    """3-D tuples: ``X[i, j, k] <- aggr_l X[i, j, l] A[l, k]``."""
    _X_NDIM, _X_MSG = 3, "X should be 3d representations"

    def __init__(self, aggr: str = "sum", optuplefeat: str = "X", opadj: str = "A",
                 message_func: Optional[Callable] = None) -> None:
        super().__init__(optuplefeat, optuplefeat, 2, opadj, 0, aggr, message_func)


class OpMessagePassingCrossSubg2D(_AdjTupleMessagePassing):
    """Message passing across subgraphs: ``X[i, j] <- aggr_k A[i, k] X[k, j]``."""
    _TUPLE_FIRST = False

    def __init__(self, aggr: str = "sum", optuplefeat: str = "X", opadj: str = "A",
                 message_func: Optional[Callable] = None) -> None:
        super().__init__(optuplefeat, opadj, 1, optuplefeat, 0, aggr, message_func)


def _dimlist(dims: Union[int, Iterable[int]]) -> List[int]:
    return sorted({dims} if isinstance(dims, int) else set(dims))


class OpDiag(Module):
    """Diagonal elements over the sparse dims ``dims``."""

    def __init__(self, dims: Iterable[int], return_sparse: bool = False) -> None:
        super().__init__()
        self.dims, self.return_sparse = _dimlist(dims), return_sparse

    def forward(self, A: SparseTensor) -> Union[Tensor, SparseTensor]:
        return A.diag(self.dims, return_sparse=self.return_sparse)


class OpDiag2D(OpDiag):
    def __init__(self) -> None:
        super().__init__([0, 1], False)

    def forward(self, X: SparseTensor) -> Tensor:
        assert X.sparse_dim == 2, "X should be 2d representations"
        return super().forward(X)


class OpPooling(Module):
    """Reduce the sparse dims ``dims`` with ``pool`` in sum / mean / max."""

    def __init__(self, dims: Union[int, Iterable[int]], pool: str = "sum",
                 return_sparse: bool = False) -> None:
        super().__init__()
        self.dims, self.pool, self.return_sparse = _dimlist(dims), pool, return_sparse

    def forward(self, X: SparseTensor) -> Union[SparseTensor, Tensor]:
        return getattr(X, self.pool)(self.dims, return_sparse=self.return_sparse)


class _FixedPooling(OpPooling):
    _DIM, _SPARSE_OUT, _NDIM, _MSG = 1, False, 2, "X should be 2d representations"

    def __init__(self, pool) -> None:
        super().__init__(self._DIM, pool, self._SPARSE_OUT)

    def forward(self, X: SparseTensor):
        assert X.sparse_dim == self._NDIM, self._MSG
        return super().forward(X)


class OpPoolingSubg2D(_FixedPooling):
    """Pool the nodes of each subgraph: (n, n, d) tuples -> (n, d) root nodes."""


class OpPoolingSubg3D(_FixedPooling):
    """Pool the last tuple dim of 3-D representations, sparse output on (i, j)."""
    _DIM, _SPARSE_OUT, _NDIM, _MSG = 2, True, 3, "X should be 3d representations"


class OpPoolingCrossSubg2D(_FixedPooling):
    """Pool the same node across subgraphs: reduce the root dim."""
    _DIM = 0


class OpUnpooling(Module):
    """Broadcast node (or lower-order tuple) features back onto the tuples of ``tarX``
    along ``dims``."""

    def __init__(self, dims: Union[int, Iterable[int]], fromdense1dim: bool = True) -> None:
        super().__init__()
        self.dims, self.fromdense1dim = _dimlist(dims), fromdense1dim

    def forward(self, X: Union[Tensor, SparseTensor], tarX: SparseTensor) -> SparseTensor:
        if not isinstance(X, Tensor):
            return X.unpooling(self.dims, tarX)
        left = [d for d in range(tarX.sparse_dim) if d not in self.dims]
        assert len(left) == 1, "canonly pooling from 1 dim"
        return tarX.unpooling_fromdense1dim(left[0], X)


class OpUnpoolingSubgNodes2D(OpUnpooling):
    """Root-node feature -> every node of its subgraph."""

    def __init__(self) -> None:
        super().__init__(1, True)


class OpUnpoolingRootNodes2D(OpUnpooling):
    """Node feature -> that node in every subgraph."""

    def __init__(self) -> None:
        super().__init__(0, True)
