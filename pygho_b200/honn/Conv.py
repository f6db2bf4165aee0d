"""High-order GNN layers assembled from the tensor operators (reference
``pygho/honn/Conv.py``: ``NGNNConv`` :20-58, ``SSWLConv`` :62-103, ``I2Conv`` :107-147,
``DSSGNNConv`` :151-196, ``PPGNConv`` :200-236, ``GNNAKConv`` :240-298).  Every layer has
the signature ``forward(A, X, datadict) -> SparseTensor | MaskedTensor``."""
from __future__ import annotations

from typing import Callable, Optional, Union

from torch.nn import Module

from ..backend.MaTensor import MaskedTensor
from ..backend.SpTensor import SparseTensor
from . import TensorOp
from .utils import MLP

AnyTensor = Union[SparseTensor, MaskedTensor]


class NGNNConv(Module):
    """Nested GNN: an MLP on every tuple, then message passing inside each subgraph."""

    def __init__(self, indim: int, outdim: int, aggr: str = "sum", mode: str = "SS",
                 mlp: dict = {}, optuplefeat: str = "X", opadj: str = "A",
                 message_func: Optional[Callable] = None):
        super().__init__()
        self.aggr = TensorOp.OpMessagePassingOnSubg2D(mode, aggr, optuplefeat, opadj, message_func)
        self.lin = MLP(indim, outdim, **mlp)

    def forward(self, A: AnyTensor, X: AnyTensor, datadict: dict) -> AnyTensor:
        h = X.tuplewiseapply(self.lin)
        return self.aggr.forward(A, h, datadict, h)


class SSWLConv(Module):
    """Subgraph WL (SSWL+): [X, X A (inside subgraphs), A X (across subgraphs)] -> MLP."""

    def __init__(self, indim: int, outdim: int, aggr: str = "sum", mode: str = "SS",
                 mlp: dict = {}, optuplefeat: str = "X", opadj: str = "A"):
        super().__init__()
        self.aggr1 = TensorOp.OpMessagePassingOnSubg2D(mode, aggr, optuplefeat, opadj)
        self.aggr2 = TensorOp.OpMessagePassingCrossSubg2D(mode, aggr, optuplefeat, opadj)
        self.lin = MLP(3 * indim, outdim, **mlp)

    def forward(self, A: AnyTensor, X: AnyTensor, datadict: dict,
                residual: Optional[AnyTensor] = None) -> AnyTensor:
        """``residual`` (not in the reference signature; same sparsity / mask as X) is added to
        the layer output inside the last kernel of the MLP: ``X + conv(X)`` of the reference
        training scripts (example/zinc.py:286) without a separate pass over the tuples."""
        res = None if residual is None else (
            residual.values if isinstance(residual, SparseTensor) else residual.data)
        # residual is X itself (the training scripts' X + conv(X)): tap it inside the aggregate
        # so that its gradient is added to dX by the backward kernel (no extra pass)
        tap = residual is not None and isinstance(X, SparseTensor) and res is X.values
        fused = self._fused_cat(A, X, datadict, tap)
        if fused is not None:
            cat, tapped = fused
            if tapped is not None:
                res = tapped
            return X.tuplewiseapply(lambda _v: self.lin(cat) if res is None else self.lin(cat, res))
        lin = self.lin if res is None else (lambda v: self.lin(v, res))
        inside = self.aggr1.forward(A, X, datadict, X)
        across = self.aggr2.forward(A, X, datadict, X)
        return X.catvalue([inside, across], True).tuplewiseapply(lin)

    def _fused_cat(self, A, X, datadict, tap_residual=False):
        """[X, X(x)A, A(x)X] written into one buffer by the two spspmm launches (sparse mode,
        precomputed plans, sum/mean, 2-D float32 values) -> (buffer, tapped X or None);
        None -> generic path."""
        from ..backend.SpTensor import SparseTensor as _Sp
        from ..honn.SpOperator import KEYSEP
        if not (isinstance(A, _Sp) and isinstance(X, _Sp)):
            return None
        m1, m2 = self.aggr1.mod, self.aggr2.mod
        if m1.use_mpnn or m2.use_mpnn or m1.aggr != m2.aggr or m1.aggr not in ("sum", "mean"):
            return None
        acd1 = datadict.get(m1.precomputekey + KEYSEP + "acd")
        acd2 = datadict.get(m2.precomputekey + KEYSEP + "acd")
        xv, av = X.values, A.values
        if acd1 is None or acd2 is None or xv is None or av is None or xv.ndim != 2 \
                or xv.shape[1:] != av.shape[1:] or xv.dtype != av.dtype or not xv.is_cuda \
                or xv.shape[1] % 4:
            return None
        import torch
        if xv.dtype != torch.float32:
            return None
        from .. import plans as P
        from ..ops import SswlAggregate
        plan_xa = P.plan_from_acd(acd1, X.nnz, X.nnz, A.nnz)
        plan_ax = P.plan_from_acd(acd2, X.nnz, A.nnz, X.nnz)
        if tap_residual and not xv.is_contiguous():
            tap_residual = False
        # sum aggregation: both gradient contributions to X in one segmented reduction
        merged = None
        if m1.aggr == "sum" and X.nnz and A.nnz and xv.requires_grad and torch.is_grad_enabled():
            merged = P.sswl_bwd_group(acd1, acd2, m2.precomputekey + KEYSEP + "acd", X.nnz, A.nnz)
        out = SswlAggregate.apply(xv.contiguous(), av.contiguous(), plan_xa, plan_ax,
                                  0 if m1.aggr == "sum" else 1, tap_residual, merged)
        return out if tap_residual else (out, None)


class I2Conv(Module):
    """I2-GNN layer on 3-D tuples: MLP, then message passing over the last tuple dim."""

    def __init__(self, indim: int, outdim: int, aggr: str = "sum", mode: str = "SS",
                 mlp: dict = {}, optuplefeat: str = "X", opadj: str = "A"):
        super().__init__()
        self.aggr = TensorOp.OpMessagePassingOnSubg3D(mode, aggr, optuplefeat, opadj)
        self.lin = MLP(indim, outdim, **mlp)

    def forward(self, A: AnyTensor, X: AnyTensor, datadict: dict) -> AnyTensor:
        h = X.tuplewiseapply(self.lin)
        return self.aggr.forward(A, h, datadict, h)


class DSSGNNConv(Module):
    """DSS-GNN / ESAN: subgraph message passing plus a global branch (pool across
    subgraphs -> node message passing -> unpool to every subgraph)."""

    def __init__(self, indim: int, outdim: int, aggr_subg: str = "sum", aggr_global: str = "sum",
                 pool: str = "mean", mode: str = "SS", mlp: dict = {}, optuplefeat: str = "X",
                 opadj: str = "A"):
        super().__init__()
        self.aggr_subg = TensorOp.OpMessagePassingOnSubg2D(mode, aggr_subg, optuplefeat, opadj)
        self.pool2global = TensorOp.OpPoolingCrossSubg2D(mode[1], pool)
        self.aggr_global = TensorOp.OpNodeMessagePassing(mode, aggr_global)
        self.unpooling2subg = TensorOp.OpUnpoolingRootNodes2D(mode[1])
        self.lin = MLP(2 * indim, outdim, **mlp)

    def forward(self, A: AnyTensor, X: AnyTensor, datadict: dict) -> AnyTensor:
        shared = self.aggr_global.forward(A, self.pool2global.forward(X))
        glob = self.unpooling2subg.forward(shared, X)
        local = self.aggr_subg.forward(A, X, datadict, X)
        return local.catvalue(glob, True).tuplewiseapply(self.lin)


class PPGNConv(Module):
    """Provably powerful GN / 2-FWL: product of two MLP branches over the middle node."""

    def __init__(self, indim: int, outdim: int, aggr: str = "sum", mode: str = "SS",
                 mlp: dict = {}, optuplefeat: str = "X"):
        super().__init__()
        self.op = TensorOp.Op2FWL(mode, aggr, optuplefeat)
        self.lin1 = MLP(indim, outdim, **mlp)
        self.lin2 = MLP(indim, outdim, **mlp)

    def forward(self, A: AnyTensor, X: AnyTensor, datadict: dict) -> AnyTensor:
        return self.op.forward(X.tuplewiseapply(self.lin1), X.tuplewiseapply(self.lin2),
                               datadict, X)


class GNNAKConv(Module):
    """GNN-AK(+ctx): subgraph message passing, then [pooled subgraph, centroid (diag),
    context (pooled across subgraphs)] broadcast back to the tuples."""

    def __init__(self, indim: int, outdim: int, aggr: str = "sum", pool: str = "mean",
                 mode: str = "SS", mlp0: dict = {}, mlp1: dict = {}, ctx: bool = True,
                 optuplefeat: str = "X", opadj: str = "A"):
        super().__init__()
        self.lin0 = MLP(indim, indim, **mlp0)
        self.aggr = TensorOp.OpMessagePassingOnSubg2D(mode, aggr, optuplefeat, opadj)
        self.diag = TensorOp.OpDiag2D(mode[1])
        self.pool2subg = TensorOp.OpPoolingSubg2D(mode[1], pool)
        self.unpool4subg = TensorOp.OpUnpoolingSubgNodes2D(mode[1])
        self.ctx = ctx
        if ctx:
            self.pool2node = TensorOp.OpPoolingCrossSubg2D(mode[1], pool)
            self.unpool4rootnode = TensorOp.OpUnpoolingRootNodes2D(mode[1])
        self.lin = MLP((3 if ctx else 2) * indim, outdim, **mlp1)

    def forward(self, A: AnyTensor, X: AnyTensor, datadict: dict) -> AnyTensor:
        X = self.aggr.forward(A, X.tuplewiseapply(self.lin0), datadict, X)
        centroid = self.unpool4subg.forward(self.diag.forward(X), X)
        subgraph = self.unpool4subg.forward(self.pool2subg.forward(X), X)
        parts = [centroid]
        if self.ctx:
            parts.append(self.unpool4rootnode.forward(self.pool2node.forward(X), X))
        return subgraph.catvalue(parts, True).tuplewiseapply(self.lin)
