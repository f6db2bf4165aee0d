"""Model assembly of the reference's ZINC example (``example/zinc.py:57-294``,
``example/minimal.py:37-85``) on top of ``pygho_b200``: the benchmark and parity-test
models.  Hyper-parameters that the reference reads from argparse globals are explicit
constructor arguments here (defaults = ``example/work.sh`` line 4: mlplayer 2, outlayer 4,
bn, silu, npool sum, lpool mean)."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn

from pygho_b200 import MaskedTensor, SparseTensor
from pygho_b200.backend.utils import torch_scatter_reduce
from pygho_b200.honn import Conv
from pygho_b200.honn.MaOperator import OpPooling
from pygho_b200.honn.TensorOp import OpPoolingSubg2D, OpPoolingSubg3D
from pygho_b200.honn.utils import MLP, Embedding

CONVS = ("NGNN", "SSWL", "DSSGNN", "PPGN", "I2GNN", "GNNAK")


def make_conv(name: str, dim: int, mode: str, mlp: Dict, aggr: str = "sum", cpool: str = "mean"):
    if name == "SSWL":
        return Conv.SSWLConv(dim, dim, aggr, mode, dict(mlp))
    if name == "NGNN":
        return Conv.NGNNConv(dim, dim, aggr, mode, dict(mlp))
    if name == "DSSGNN":
        return Conv.DSSGNNConv(dim, dim, aggr, aggr, cpool, mode, dict(mlp))
    if name == "PPGN":
        return Conv.PPGNConv(dim, dim, aggr, mode, dict(mlp))
    if name == "I2GNN":
        return Conv.I2Conv(dim, dim, aggr, mode, dict(mlp))
    if name == "GNNAK":
        return Conv.GNNAKConv(dim, dim, aggr, cpool, mode, dict(mlp), dict(mlp))
    raise ValueError(f"unknown conv {name}")


class SpModel(nn.Module):
    """Sparse-representation model (zinc.py:222-294)."""

    def __init__(self, conv: str = "SSWL", num_tasks: int = 1, num_layer: int = 6,
                 hiddim: int = 128, aggr: str = "sum", npool: str = "sum", lpool: str = "mean",
                 cpool: str = "mean", mlplayer: int = 2, outlayer: int = 4, norm: str = "bn",
                 residual: bool = True, normparam: float = 0.1):
        super().__init__()
        base = {"dp": 0.0, "norm": norm, "act": "silu", "normparam": normparam}
        convmlp = dict(base, tailact=True, numlayer=mlplayer)
        self.conv_name, self.residual, self.npool = conv, residual, npool
        # nn.Embedding subclasses (same state dict): gather kernel + deterministic gradient
        self.x_encoder = Embedding(32, hiddim)
        self.ea_encoder = Embedding(16, hiddim)
        self.tuplefeat_encoder = Embedding(16, hiddim)
        if conv == "I2GNN":
            self.tuplefeat_encoder2 = Embedding(16, hiddim)
        self.lin_tupleinit0 = nn.Linear(hiddim, hiddim)
        self.lin_tupleinit1 = nn.Linear(hiddim, hiddim)
        if conv == "I2GNN":
            self.lin_tupleinit2 = nn.Linear(hiddim, hiddim)
        self.subggnns = nn.ModuleList(
            [make_conv(conv, hiddim, "SS", convmlp, aggr, cpool) for _ in range(num_layer)])
        self.lpool = (nn.Sequential(OpPoolingSubg3D("S", lpool), OpPoolingSubg2D("S", lpool))
                      if conv == "I2GNN" else OpPoolingSubg2D("S", lpool))
        self.poolmlp = MLP(hiddim, hiddim, mlplayer, tailact=True, **base)
        self.pred_lin = MLP(hiddim, num_tasks, outlayer, tailact=False, **base)

    def encode(self, datadict: dict):
        x = self.x_encoder(datadict["x"].flatten())
        A = datadict["A"].tuplewiseapply(self.ea_encoder)
        if self.conv_name == "I2GNN":
            X = datadict["X"].tuplewiseapply(
                lambda f: self.tuplefeat_encoder(f[:, 0]) + self.tuplefeat_encoder2(f[:, 1]))
        else:
            X = datadict["X"].tuplewiseapply(self.tuplefeat_encoder)
        return x, A, X

    def tupleinit(self, X: SparseTensor, x: torch.Tensor) -> SparseTensor:
        # zinc.py:270-276 indexes with X.indices[0] / [1]; use the gather kernel
        if self.conv_name == "I2GNN":
            root = X.unpooling_fromdense1dim(0, self.lin_tupleinit0(x)).values
            node = X.unpooling_fromdense1dim(1, self.lin_tupleinit1(x)).values
            # zinc.py:271-273: the third factor is gathered with X.indices[1] too (kept as is)
            third = X.unpooling_fromdense1dim(1, self.lin_tupleinit2(x)).values
            return X.tuplewiseapply(lambda val: root * node * third * val)
        # root[i] * node[j] * val in one gather-multiply launch + one product (ops.GatherProduct)
        return X.gather_product(self.lin_tupleinit0(x), self.lin_tupleinit1(x))

    def forward(self, datadict: dict) -> torch.Tensor:
        x, A, X = self.encode(datadict)
        X = self.tupleinit(X, x)
        for conv in self.subggnns:
            if self.residual and isinstance(conv, Conv.SSWLConv):
                X = conv.forward(A, X, datadict, residual=X)     # X + conv(X), add fused
                continue
            tX = conv.forward(A, X, datadict)
            X = X.add(tX, True) if self.residual else tX
        x = self.poolmlp(self.lpool(X))
        h_graph = torch_scatter_reduce(0, x, datadict["batch"], datadict["num_graphs"], self.npool)
        return self.pred_lin(h_graph)


class MaModel(nn.Module):
    """Dense-representation model (zinc.py:155-219)."""

    def __init__(self, conv: str = "PPGN", num_tasks: int = 1, num_layer: int = 6,
                 hiddim: int = 128, aggr: str = "sum", npool: str = "sum", lpool: str = "mean",
                 cpool: str = "mean", mlplayer: int = 2, outlayer: int = 4, norm: str = "bn",
                 residual: bool = True, normparam: float = 0.1):
        super().__init__()
        base = {"dp": 0.0, "norm": norm, "act": "silu", "normparam": normparam}
        convmlp = dict(base, tailact=True, numlayer=mlplayer)
        self.residual = residual
        self.x_encoder = nn.Embedding(32, hiddim)
        self.ea_encoder = nn.Embedding(16, hiddim, padding_idx=0)
        self.tuplefeat_encoder = nn.Embedding(16, hiddim)
        self.lin_tupleinit0 = nn.Linear(hiddim, hiddim)
        self.lin_tupleinit1 = nn.Linear(hiddim, hiddim)
        self.subggnns = nn.ModuleList(
            [make_conv(conv, hiddim, "DD", convmlp, aggr, cpool) for _ in range(num_layer)])
        self.npool = OpPooling(1, pool=npool)
        self.lpool = OpPoolingSubg2D("D", pool=lpool)
        self.poolmlp = MLP(hiddim, hiddim, mlplayer, tailact=True, **base)
        self.pred_lin = MLP(hiddim, num_tasks, outlayer, tailact=False, **base)

    def forward(self, datadict: dict) -> torch.Tensor:
        x: MaskedTensor = datadict["x"].tuplewiseapply(lambda v: self.x_encoder(v.squeeze(-1)))
        A: MaskedTensor = datadict["A"].tuplewiseapply(self.ea_encoder)
        X: MaskedTensor = datadict["X"].tuplewiseapply(self.tuplefeat_encoder)
        xz = x.fill_masked(0.0)
        X = X.tuplewiseapply(lambda val: self.lin_tupleinit0(xz).unsqueeze(1) *
                             self.lin_tupleinit1(xz).unsqueeze(2) * val)
        for conv in self.subggnns:
            tX = conv.forward(A, X, datadict)
            X = X.add(tX, samesparse=True) if self.residual else tX
        x = self.lpool(X).tuplewiseapply(self.poolmlp)
        return self.pred_lin(self.npool.forward(x).fill_masked(0.0))
