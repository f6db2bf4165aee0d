"""Dense (``MaskedTensor``) graph operators (reference ``pygho/honn/MaOperator.py``:
``OpNodeMessagePassing`` :14, ``OpSpNodeMessagePassing`` :45, ``OpMessagePassing`` :83,
``Op2FWL`` :126, the Subg2D/3D/CrossSubg2D variants :163-278, sparse-adjacency variants
:281-372, ``OpDiag`` :375-405, ``OpPooling`` :408-464, ``OpUnpooling`` :467-521).
Masked dims are shifted by one w.r.t. the sparse operators because dim 0 is the batch."""
from __future__ import annotations

from typing import Dict, Iterable, List, Union

from torch import Tensor
from torch.nn import Module

from ..backend.Mamamm import mamamm
from ..backend.MaTensor import MaskedTensor
from ..backend.SpTensor import SparseTensor


from ..backend.Spmamm import spmamm


class OpNodeMessagePassing(Module):
    """``X' = A X`` with A (b, n, n, d) and X (b, n, d) masked tensors."""

    def forward(self, A: MaskedTensor, X: MaskedTensor, tarX: MaskedTensor) -> MaskedTensor:
        # the reference drops the unsqueeze result for 2-masked-dim operands (Q3); do what
        # was intended: treat X as (b, n, 1) tuples and squeeze the result
        X3 = MaskedTensor(X.data.unsqueeze(2), X.mask.unsqueeze(2), X.padvalue, True)
        out = mamamm(A, 2, X3, 1, tarX.mask.unsqueeze(2))
        return MaskedTensor(out.data.squeeze(2), tarX.mask, 0.0, True)


class OpSpNodeMessagePassing(Module):
    """``X' = A X`` with a sparse batched adjacency A (b, n, n) and X (b, n, d) masked."""

    def __init__(self, aggr: str = "sum") -> None:
        super().__init__()
        self.aggr = aggr

    def forward(self, A: SparseTensor, X: MaskedTensor, tarX: MaskedTensor) -> MaskedTensor:
        assert A.sparse_dim == 3, "A should be bxnxn adjacency matrix "
        return spmamm(A, 2, X, 1, tarX.mask, self.aggr)


class OpMessagePassing(Module):
    """``mamamm(A, dim1, B, dim2)`` masked by the target's mask."""

    def __init__(self, dim1: int, dim2: int) -> None:
        super().__init__()
        self.dim1, self.dim2 = dim1, dim2

    def forward(self, A: MaskedTensor, B: MaskedTensor, tarX: MaskedTensor) -> MaskedTensor:
        return mamamm(A, self.dim1, B, self.dim2, tarX.mask, True)


class _DenseVariant(OpMessagePassing):
    _DIMS = (2, 1)
    _NDIM1, _NDIM2 = 3, 3
    _MSG1, _MSG2 = "A should be bxnxn adjacency matrix ", "X should be bxnxn 2d representations"
    _SWAP = False      # True: the tuple operand goes first

    def __init__(self) -> None:
        super().__init__(*self._DIMS)

    def forward(self, A: MaskedTensor, X: MaskedTensor, datadict: Dict,
                tarX: MaskedTensor) -> MaskedTensor:
        assert A.masked_dim == self._NDIM1, self._MSG1
        assert X.masked_dim == self._NDIM2, self._MSG2
        first, second = (X, A) if self._SWAP else (A, X)
        return OpMessagePassing.forward(self, first, second, tarX)


class Op2FWL(_DenseVariant):
    """2-FWL on dense tuples: ``X[b,i,j] <- sum_k X1[b,i,k] X2[b,k,j]``."""
    _MSG1, _MSG2 = "X1 should be bxnxn adjacency matrix ", "X2 should be bxnxn 2d representations"


class OpMessagePassingOnSubg2D(_DenseVariant):
    _SWAP = True


class OpMessagePassingOnSubg3D(_DenseVariant):
    _DIMS, _NDIM2, _SWAP = (3, 1), 4, True
    _MSG2 = "X should be bxnxnxn 3d representations"


class OpMessagePassingCrossSubg2D(_DenseVariant):
    _DIMS = (1, 1)


class OpSpMessagePassing(Module):
    def __init__(self, dim1: int, dim2: int, aggr: str = "sum") -> None:
        super().__init__()
        self.dim1, self.dim2, self.aggr = dim1, dim2, aggr

    def forward(self, A: SparseTensor, X: MaskedTensor, tarX: MaskedTensor) -> MaskedTensor:
        assert A.sparse_dim == 3, "A should be bxnxn adjacency matrix "
        return spmamm(A, self.dim1, X, self.dim2, tarX.mask, self.aggr)


class _SpVariant(OpSpMessagePassing):
    _NDIM, _MSG = 3, "X should be bxnxn 2D representation "

    def forward(self, A: SparseTensor, X: MaskedTensor, datadict: Dict,
                tarX: MaskedTensor) -> MaskedTensor:
        assert X.masked_dim == self._NDIM, self._MSG
        return OpSpMessagePassing.forward(self, A, X, tarX)


class OpSpMessagePassingOnSubg2D(_SpVariant):
    def __init__(self, aggr: str = "sum") -> None:
        super().__init__(1, 2, aggr)


class OpSpMessagePassingOnSubg3D(_SpVariant):
    _NDIM, _MSG = 4, "X should be bxnxnxn 3D representation "

    def __init__(self, aggr: str = "sum") -> None:
        super().__init__(1, 3, aggr)


class OpSpMessagePassingCrossSubg2D(_SpVariant):
    def __init__(self, aggr: str = "sum") -> None:
        super().__init__(1, 1, aggr)


def _dimlist(dims: Union[int, Iterable[int]]) -> List[int]:
    return sorted({dims} if isinstance(dims, int) else set(dims))


class OpDiag(Module):
    def __init__(self, dims: Iterable[int]) -> None:
        super().__init__()
        self.dims = _dimlist(dims)

    def forward(self, A: MaskedTensor) -> MaskedTensor:
        return A.diag(self.dims)


class OpDiag2D(OpDiag):
    def __init__(self) -> None:
        super().__init__([1, 2])

    def forward(self, X: MaskedTensor) -> MaskedTensor:
        assert X.masked_dim == 3, "X should be bxnxn 2d representations"
        return super().forward(X)


class OpPooling(Module):
    def __init__(self, dims: Union[int, Iterable[int]], pool: str = "sum") -> None:
        super().__init__()
        self.dims, self.pool = _dimlist(dims), pool

    def forward(self, X: MaskedTensor) -> MaskedTensor:
        return getattr(X, self.pool)(dims=self.dims, keepdim=False)


class _FixedPooling(OpPooling):
    _DIMS, _NDIM, _MSG = [2], 3, "X should be bxnxn 2d representations"

    def __init__(self, pool: str = "sum") -> None:
        super().__init__(self._DIMS, pool)

    def forward(self, X: MaskedTensor) -> MaskedTensor:
        assert X.masked_dim == self._NDIM, self._MSG
        return super().forward(X)


class OpPoolingSubg2D(_FixedPooling):
    pass


class OpPoolingSubg3D(_FixedPooling):
    _DIMS, _NDIM, _MSG = [3], 4, "X should be bxnxnxn 3d representations"


class OpPoolingCrossSubg2D(_FixedPooling):
    _DIMS = [1]


class OpUnpooling(Module):
    def __init__(self, dims: Union[int, Iterable[int]]) -> None:
        super().__init__()
        self.dims = _dimlist(dims)

    def forward(self, X: MaskedTensor, tarX: MaskedTensor) -> MaskedTensor:
        return X.unpooling(self.dims, tarX)


class OpUnpoolingSubgNodes2D(OpUnpooling):
    def __init__(self) -> None:
        super().__init__([2])


class OpUnpoolingRootNodes2D(OpUnpooling):
    def __init__(self) -> None:
        super().__init__([1])
