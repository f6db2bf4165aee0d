"""Representation-agnostic operators: pick the sparse or the dense implementation from a
mode string (reference ``pygho/honn/TensorOp.py:14-500``).  Two-letter modes name the
representation of (A, X): "SS" sparse/sparse, "DD" dense/dense, "SD" sparse adjacency with
dense tuples; one-letter modes name X only."""
from __future__ import annotations

from typing import Callable, Dict, Optional, Union

from torch import Tensor
from torch.nn import Module

from ..backend.MaTensor import MaskedTensor
from ..backend.SpTensor import SparseTensor
from . import MaOperator, SpOperator

AnyTensor = Union[SparseTensor, MaskedTensor]


def _need_sum(aggr: str, what: str) -> None:
    assert aggr == "sum", f"only sum aggragation implemented for {what}"


def _no_message_func(message_func) -> None:
    assert message_func is None, \
        "general message passing with message_func is not implemented for Dense"


class OpNodeMessagePassing(Module):
    def __init__(self, mode: str = "SS", aggr: str = "sum") -> None:
        super().__init__()
        if mode == "SS":
            self.mod = SpOperator.OpNodeMessagePassing(aggr)
        elif mode == "SD":
            self.mod = MaOperator.OpSpNodeMessagePassing(aggr)
        elif mode == "DD":
            assert aggr == "sum", f"aggr {aggr} is not implemented for DD"
            self.mod = MaOperator.OpNodeMessagePassing()
        else:
            raise NotImplementedError

    def forward(self, A: AnyTensor, X: Union[Tensor, MaskedTensor]):
        return self.mod.forward(A, X, X)


class _TupleMessagePassing(Module):
    """forward(A, X, datadict, tarX) dispatching on the mode given at construction."""
    _SP, _MA, _MASP = None, None, None

    def __init__(self, mode: str = "SS", aggr: str = "sum", optuplefeat: str = "X",
                 opadj: str = "A", message_func: Optional[Callable] = None) -> None:
        super().__init__()
        if mode == "SS":
            self.mod = self._make_sparse(aggr, optuplefeat, opadj, message_func)
        elif mode == "SD":
            _no_message_func(message_func)
            self.mod = getattr(MaOperator, self._MASP)(aggr)
        elif mode == "DD":
            _no_message_func(message_func)
            _need_sum(aggr, "Dense adjacency")
            self.mod = getattr(MaOperator, self._MA)()
        else:
            raise NotImplementedError

    def _make_sparse(self, aggr, optuplefeat, opadj, message_func):
        return getattr(SpOperator, self._SP)(aggr, optuplefeat, opadj, message_func)

    def forward(self, A: AnyTensor, X: AnyTensor, datadict: Optional[Dict] = None,
                tarX: Optional[AnyTensor] = None) -> AnyTensor:
        return self.mod.forward(A, X, datadict, tarX)


class OpMessagePassingOnSubg2D(_TupleMessagePassing):
    _SP, _MA, _MASP = ("OpMessagePassingOnSubg2D", "OpMessagePassingOnSubg2D",
                       "OpSpMessagePassingOnSubg2D")


class OpMessagePassingOnSubg3D(_TupleMessagePassing):
    _SP, _MA, _MASP = ("OpMessagePassingOnSubg3D", "OpMessagePassingOnSubg3D",
                       "OpSpMessagePassingOnSubg3D")


class OpMessagePassingCrossSubg2D(_TupleMessagePassing):
    _SP, _MA, _MASP = ("OpMessagePassingCrossSubg2D", "OpMessagePassingCrossSubg2D",
                       "OpSpMessagePassingCrossSubg2D")


class Op2FWL(Module):
    def __init__(self, mode: str = "SS", aggr: str = "sum", optuplefeat: str = "X") -> None:
        super().__init__()
        if mode == "SS":
            self.mod = SpOperator.Op2FWL(aggr, optuplefeat)
        elif mode == "DD":
            _need_sum(aggr, "Dense adjacency")
            self.mod = MaOperator.Op2FWL()
        else:
            raise NotImplementedError

    def forward(self, X1: AnyTensor, X2: AnyTensor, datadict: Optional[Dict] = None,
                tarX: Optional[AnyTensor] = None) -> AnyTensor:
        return self.mod.forward(X1, X2, datadict, tarX)


class _ByRepresentation(Module):
    """One-letter mode: "S" -> SpOperator.<name>, "D" -> MaOperator.<name>."""
    _NAME = ""

    def __init__(self, mode: str = "S", *args) -> None:
        super().__init__()
        if mode not in ("S", "D"):
            raise NotImplementedError
        self.mod = getattr(SpOperator if mode == "S" else MaOperator, self._NAME)(*args)


class OpDiag2D(_ByRepresentation):
    _NAME = "OpDiag2D"

    def forward(self, X: AnyTensor) -> Union[MaskedTensor, Tensor]:
        return self.mod.forward(X)


class _Pooling(_ByRepresentation):
    def __init__(self, mode: str = "S", pool: str = "sum") -> None:
        super().__init__(mode, pool)

    def forward(self, X: AnyTensor) -> Union[MaskedTensor, SparseTensor, Tensor]:
        return self.mod(X)


class OpPoolingSubg2D(_Pooling):
    _NAME = "OpPoolingSubg2D"


class OpPoolingSubg3D(_Pooling):
    _NAME = "OpPoolingSubg3D"


class OpPoolingCrossSubg2D(_Pooling):
    _NAME = "OpPoolingCrossSubg2D"


class _Unpooling(_ByRepresentation):
    def forward(self, X: Union[Tensor, MaskedTensor], tarX: AnyTensor) -> AnyTensor:
        return self.mod.forward(X, tarX)


class OpUnpoolingSubgNodes2D(_Unpooling):
    _NAME = "OpUnpoolingSubgNodes2D"


class OpUnpoolingRootNodes2D(_Unpooling):
    _NAME = "OpUnpoolingRootNodes2D"
