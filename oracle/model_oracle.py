"""CPU restatement of the conv layers and the sparse ZINC model -- TEST INFRASTRUCTURE ONLY.

Used to (1) check the product's layers end to end (forward values, parameter gradients)
and (2) time the reference's CPU path for ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs on the GPU box, where ``/root/reference`` does not exist.

It restates ``pygho/honn/Conv.py`` (``NGNNConv`` :20-58, ``SSWLConv`` :62-103,
``I2Conv`` :107-147, ``DSSGNNConv`` :151-196, ``PPGNConv`` :200-236), ``pygho/honn/utils.py:85-142`` (MLP) and
``example/zinc.py:222-294`` (SpModel) with plain torch CPU ops via
``oracle/torch_oracle.py``.  Tensors are passed as (indices, values) pairs, plans as the
reference's (3, T) LongTensors.  Parameter names equal the product's
(``examples/zinc_models.py``) so a ``state_dict`` can be copied across.  Pinned by
``tests/test_oracle_golden.py::test_oracle_convs_match_reference`` against outputs and
gradients of the real reference layers.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn

from . import torch_oracle as TO

K_XA = "X___X___1___A___0___acd"
K_AX = "X___A___1___X___0___acd"
K_XX = "X___X___1___X___0___acd"
K_I2 = "X___X___2___A___0___acd"


class _BN(nn.Module):
    def __init__(self, dim, momentum):
        super().__init__()
        self.norm = nn.BatchNorm1d(dim, momentum=momentum)

    def forward(self, x):
        return self.norm(x)


class OMLP(nn.Module):
    """[Linear, BatchNorm, SiLU] x numlayer (last block bare when not tailact)."""

    def __init__(self, hiddim, outdim, numlayer, tailact, normparam=0.1, norm="bn"):
        super().__init__()
        layers = []
        for i in range(numlayer):
            last = i == numlayer - 1
            width = outdim if last else hiddim
            layers.append(nn.Linear(hiddim, width))
            if not last or tailact:
                layers.append(_BN(width, normparam) if norm == "bn" else nn.Identity())
                layers.append(nn.SiLU())
        self.lins = nn.Sequential(*layers)

    def forward(self, x):
        return self.lins(x)


def _mlp(indim, outdim, numlayer, normparam):
    """Reference MLP(hiddim=indim, outdim): hidden blocks indim->indim, last indim->outdim."""
    return OMLP(indim, outdim, numlayer, True, normparam)


class ONGNN(nn.Module):
    def __init__(self, dim, aggr, mlplayer, normparam):
        super().__init__()
        self.a, self.lin = aggr, _mlp(dim, dim, mlplayer, normparam)

    def forward(self, A, X, g):
        h = self.lin(X)
        return TO.spspmm(h, A, g[K_XA], X.shape[0], self.a)


class OSSWL(nn.Module):
    def __init__(self, dim, aggr, mlplayer, normparam):
        super().__init__()
        self.a, self.lin = aggr, _mlp(3 * dim, dim, mlplayer, normparam)

    def forward(self, A, X, g):
        inside = TO.spspmm(X, A, g[K_XA], X.shape[0], self.a)
        across = TO.spspmm(A, X, g[K_AX], X.shape[0], self.a)
        return self.lin(torch.cat([X, inside, across], dim=-1))


class ODSSGNN(nn.Module):
    def __init__(self, dim, aggr, mlplayer, normparam, pool="mean"):
        super().__init__()
        self.a, self.pool, self.lin = aggr, pool, _mlp(2 * dim, dim, mlplayer, normparam)

    def forward(self, A, X, g):
        N = g["num_nodes"]
        shared = TO.scatter_reduce(X, g["tupleid"][1], N, self.pool)
        shared = TO.spmm(g["edge_index"], A, (N, N), 1, shared, self.a)
        glob = shared.index_select(0, g["tupleid"][1])
        local = TO.spspmm(X, A, g[K_XA], X.shape[0], self.a)
        return self.lin(torch.cat([local, glob], dim=-1))


class OPPGN(nn.Module):
    def __init__(self, dim, aggr, mlplayer, normparam):
        super().__init__()
        self.a = aggr
        self.lin1, self.lin2 = _mlp(dim, dim, mlplayer, normparam), _mlp(dim, dim, mlplayer, normparam)

    def forward(self, A, X, g):
        return TO.spspmm(self.lin1(X), self.lin2(X), g[K_XX], X.shape[0], self.a)


class OI2(nn.Module):
    """Conv.py:107-147: MLP on every 3-D tuple, then message passing over the last tuple dim."""

    def __init__(self, dim, aggr, mlplayer, normparam):
        super().__init__()
        self.a, self.lin = aggr, _mlp(dim, dim, mlplayer, normparam)

    def forward(self, A, X, g):
        h = self.lin(X)
        return TO.spspmm(h, A, g[K_I2], X.shape[0], self.a)


def pool3d_to_dense(X: torch.Tensor, tid: torch.Tensor, N: int, pool: str) -> torch.Tensor:
    """example/zinc.py:258 read-out of 3-D tuples: OpPoolingSubg3D (pool dim 2 to the sparse
    (i, j) pattern, SpTensor.py:376-380 -> coalesce) then OpPoolingSubg2D (pool dim 1)."""
    key = tid[0] * N + tid[1]
    uniq, inv = torch.unique(key, return_inverse=True)
    pairs = TO.scatter_reduce(X, inv, uniq.numel(), pool)
    return TO.scatter_reduce(pairs, torch.div(uniq, N, rounding_mode="floor"), N, pool)


OCONVS = {"NGNN": ONGNN, "SSWL": OSSWL, "DSSGNN": ODSSGNN, "PPGN": OPPGN, "I2GNN": OI2}


class OSpModel(nn.Module):
    """example/zinc.py:222-294 for the 2-D sparse convs; ``g`` is a dict of CPU tensors:
    x, edge_index, edge_attr, tupleid, tuplefeat, batch, num_graphs, num_nodes, plans."""

    def __init__(self, conv="SSWL", num_layer=6, hiddim=128, aggr="sum", npool="sum",
                 lpool="mean", mlplayer=2, outlayer=4, normparam=0.1, num_tasks=1):
        super().__init__()
        self.npool, self.lpool_name, self.i2 = npool, lpool, conv == "I2GNN"
        self.x_encoder = nn.Embedding(32, hiddim)
        self.ea_encoder = nn.Embedding(16, hiddim)
        self.tuplefeat_encoder = nn.Embedding(16, hiddim)
        if self.i2:
            self.tuplefeat_encoder2 = nn.Embedding(16, hiddim)
        self.lin_tupleinit0 = nn.Linear(hiddim, hiddim)
        self.lin_tupleinit1 = nn.Linear(hiddim, hiddim)
        if self.i2:
            self.lin_tupleinit2 = nn.Linear(hiddim, hiddim)
        self.subggnns = nn.ModuleList(
            [OCONVS[conv](hiddim, aggr, mlplayer, normparam) for _ in range(num_layer)])
        self.poolmlp = OMLP(hiddim, hiddim, mlplayer, True, normparam)
        self.pred_lin = OMLP(hiddim, num_tasks, outlayer, False, normparam)

    def forward(self, g: Dict) -> torch.Tensor:
        x = self.x_encoder(g["x"])
        A = self.ea_encoder(g["edge_attr"])
        tid, N = g["tupleid"], g["num_nodes"]
        if self.i2:
            # zinc.py:104 (two distance labels per 3-D tuple) and :270-273: the third factor is
            # indexed with X.indices[1] as well (reference behaviour, kept)
            tf = g["tuplefeat"]
            X = self.tuplefeat_encoder(tf[:, 0]) + self.tuplefeat_encoder2(tf[:, 1])
            X = self.lin_tupleinit0(x).index_select(0, tid[0]) * \
                self.lin_tupleinit1(x).index_select(0, tid[1]) * \
                self.lin_tupleinit2(x).index_select(0, tid[1]) * X
        else:
            X = self.tuplefeat_encoder(g["tuplefeat"])
            X = self.lin_tupleinit0(x).index_select(0, tid[0]) * \
                self.lin_tupleinit1(x).index_select(0, tid[1]) * X
        for conv in self.subggnns:
            X = X + conv(A, X, g)
        if self.i2:
            h = pool3d_to_dense(X, tid, N, self.lpool_name)
        else:
            h = TO.scatter_reduce(X, tid[0], N, self.lpool_name)
        h = self.poolmlp(h)
        return self.pred_lin(TO.scatter_reduce(h, g["batch"], g["num_graphs"], self.npool))


class OPPGNDense(nn.Module):
    """PPGNConv in DD mode (Conv.py:200-236 with MaOperator.Op2FWL :126-160): two MLP branches
    applied to every (b, i, j) position -- BatchNorm statistics run over ALL positions of the
    padded tensor, pads included, exactly like the reference's tuplewiseapply on the full data
    (MaTensor.py:236-239) -- then the masked 2-FWL contraction."""

    def __init__(self, dim, mlplayer, normparam):
        super().__init__()
        self.lin1, self.lin2 = _mlp(dim, dim, mlplayer, normparam), _mlp(dim, dim, mlplayer, normparam)

    def forward(self, X, mask):
        m = mask.unsqueeze(-1)
        shp = X.shape
        a = self.lin1(X.reshape(-1, shp[-1])).reshape(shp) * m
        b = self.lin2(X.reshape(-1, shp[-1])).reshape(shp) * m
        return TO.mamamm(a, 2, b, 1, mask)


class OMaModel(nn.Module):
    """example/zinc.py:155-219 (dense PPGN model) with the intended MaskedTensor semantics
    (pads hold 0).  ``g``: x (b, n) int64, A (b, n, n) int64 edge labels
    (0 = no edge), X (b, n, n) int64 tuple labels, nmask (b, n) bool."""

    def __init__(self, num_layer=6, hiddim=128, npool="sum", lpool="mean", mlplayer=2,
                 outlayer=4, normparam=0.1, num_tasks=1):
        super().__init__()
        self.npool, self.lpool_name = npool, lpool
        self.x_encoder = nn.Embedding(32, hiddim)
        self.ea_encoder = nn.Embedding(16, hiddim, padding_idx=0)
        self.tuplefeat_encoder = nn.Embedding(16, hiddim)
        self.lin_tupleinit0 = nn.Linear(hiddim, hiddim)
        self.lin_tupleinit1 = nn.Linear(hiddim, hiddim)
        self.subggnns = nn.ModuleList(
            [OPPGNDense(hiddim, mlplayer, normparam) for _ in range(num_layer)])
        self.poolmlp = OMLP(hiddim, hiddim, mlplayer, True, normparam)
        self.pred_lin = OMLP(hiddim, num_tasks, outlayer, False, normparam)

    def forward(self, g: Dict) -> torch.Tensor:
        nm = g["nmask"]
        m2 = nm.unsqueeze(2) & nm.unsqueeze(1)
        x = self.x_encoder(g["x"]) * nm.unsqueeze(-1)
        X = self.tuplefeat_encoder(g["X"]) * m2.unsqueeze(-1)
        X = self.lin_tupleinit0(x).unsqueeze(1) * self.lin_tupleinit1(x).unsqueeze(2) * X
        X = X * m2.unsqueeze(-1)
        for conv in self.subggnns:
            X = X + conv(X, m2)
        h = TO.ma_pool(X, m2, (2,), self.lpool_name)                     # (b, n, d)
        shp = h.shape
        h = self.poolmlp(h.reshape(-1, shp[-1])).reshape(shp) * nm.unsqueeze(-1)
        return self.pred_lin(TO.ma_pool(h, nm, (1,), self.npool))


def host_dense_dict(hb, max_dist: int = 5) -> Dict:
    """CPU tensors of a HostBatch in the dense layout of ``pygho_b200.hodata.device.ma_datadict``
    (reference hodata/MaData.py:109-255): labels padded to the largest graph."""
    import numpy as np
    B = hb.num_graphs
    sizes = np.diff(hb.node_ptr)
    n = int(sizes.max())
    x = np.zeros((B, n), np.int64)
    A = np.zeros((B, n, n), np.int64)
    X = np.zeros((B, n, n), np.int64)
    nmask = np.arange(n)[None, :] < sizes[:, None]
    gb = hb.batch
    x[gb, np.arange(hb.num_nodes) - hb.node_ptr[gb]] = hb.x
    eg = gb[hb.edge_index[0]]
    A[eg, hb.edge_index[0] - hb.node_ptr[eg], hb.edge_index[1] - hb.node_ptr[eg]] = hb.edge_attr
    tg = gb[hb.tupleid[0]]
    X[tg, hb.tupleid[0] - hb.node_ptr[tg], hb.tupleid[1] - hb.node_ptr[tg]] = \
        np.minimum(hb.tuplefeat, max_dist) + 1
    return {"x": torch.from_numpy(x), "A": torch.from_numpy(A), "X": torch.from_numpy(X),
            "nmask": torch.from_numpy(nmask), "y": torch.from_numpy(hb.y), "num_graphs": B}


def host_graph_dict(hb, plans: Dict[str, torch.Tensor]) -> Dict:
    """CPU tensors of a ``pygho_b200.hodata.synthetic.HostBatch`` for :class:`OSpModel`."""
    g = {"x": torch.from_numpy(hb.x), "edge_index": torch.from_numpy(hb.edge_index),
         "edge_attr": torch.from_numpy(hb.edge_attr), "tupleid": torch.from_numpy(hb.tupleid),
         "tuplefeat": torch.from_numpy(hb.tuplefeat), "batch": torch.from_numpy(hb.batch),
         "num_graphs": hb.num_graphs, "num_nodes": hb.num_nodes, "y": torch.from_numpy(hb.y)}
    g.update(plans)
    return g
