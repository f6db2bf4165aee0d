"""Sparse-mode batch containers (reference ``pygho/hodata/SpData.py``): the collate offsets of
``SpHoData.__inc__`` / ``__cat_dim__`` (:56-77) and ``batch2sparse`` (:80-112), PyG-free.

The reference subclasses torch_geometric's ``Data`` so that PyG's ``Batch.from_data_list``
concatenates ``tupleid`` / ``<key>___acd`` along dim 1 and shifts them by per-graph increments;
``collate_sparse`` below does exactly that concatenation for plain per-graph dicts."""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, List, Sequence

import torch

from ..backend.SpTensor import SparseTensor
from ..honn.SpOperator import KEYSEP


def parseop(op: str) -> str:
    """Name of the per-graph size attribute an operator's indices are shifted by
    (reference SpData.py:15-31): ``A`` -> number of edges, ``X<key>`` -> number of tuples."""
    if op == "A":
        return "num_edges"
    if op.startswith("X"):
        return f"num_tuples{op[1:]}"
    raise NotImplementedError(f"operator name {op} not implemented now")


def parsekey(key: str):
    assert len(key.split(KEYSEP)) == 5, "key format not match"
    op0, op1, dim1, op2, dim2 = key.split(KEYSEP)
    parseop(op0), parseop(op1), parseop(op2)
    return op0, op1, int(dim1), op2, int(dim2)


def collate_sparse(graphs: Sequence[Dict], keys: Sequence[str] = ("",)):
    """Concatenate per-graph dicts (``x``, ``edge_index``, ``edge_attr``, ``tupleid<key>``,
    ``tuplefeat<key>``, ``tupleshape<key>``, optional ``<pkey>___acd`` plans, ``y``) with the
    increments of ``SpHoData.__inc__``: node ids by the node count, ``tupleid`` by the graph's
    ``tupleshape``, every row of an ``acd`` plan by the number of tuples / edges of the operator
    it indexes.  Returns a namespace with the reference's batch attribute names."""
    out = SimpleNamespace()
    n_nodes = [int(g["x"].shape[0]) for g in graphs]
    node_off = [0]
    for n in n_nodes:
        node_off.append(node_off[-1] + n)
    out.num_nodes, out.num_graphs = node_off[-1], len(graphs)
    out.ptr = torch.tensor(node_off, dtype=torch.long)
    out.batch = torch.repeat_interleave(torch.arange(len(graphs)), torch.tensor(n_nodes))
    out.x = torch.cat([g["x"] for g in graphs])
    out.edge_index = torch.cat([g["edge_index"] + node_off[i] for i, g in enumerate(graphs)], dim=1)
    out.edge_attr = torch.cat([g["edge_attr"] for g in graphs])
    if "y" in graphs[0]:
        out.y = torch.cat([torch.as_tensor(g["y"]).reshape(1, -1) for g in graphs])
    counts = {"num_edges": [int(g["edge_index"].shape[1]) for g in graphs]}
    for key in keys:
        shapes = torch.stack([torch.as_tensor(g[f"tupleshape{key}"]).reshape(-1) for g in graphs])
        inc = torch.cumsum(shapes, 0) - shapes                       # exclusive prefix of tupleshape
        setattr(out, f"tupleshape{key}", shapes)
        setattr(out, f"tupleid{key}", torch.cat(
            [g[f"tupleid{key}"] + inc[i].reshape(-1, 1) for i, g in enumerate(graphs)], dim=1))
        setattr(out, f"tuplefeat{key}", torch.cat([g[f"tuplefeat{key}"] for g in graphs]))
        counts[f"num_tuples{key}"] = [int(g[f"tupleid{key}"].shape[1]) for g in graphs]
    for name in [k for k in graphs[0] if k.endswith(KEYSEP + "acd")]:
        ops = parsekey(name[:-len(KEYSEP + "acd")])
        rows = [parseop(ops[0]), parseop(ops[1]), parseop(ops[3])]
        offs = [[0] * len(graphs) for _ in range(3)]
        for r, attr in enumerate(rows):
            acc = 0
            for i, c in enumerate(counts[attr]):
                offs[r][i] = acc
                acc += c
        setattr(out, name, torch.cat(
            [g[name] + torch.tensor([[offs[0][i]], [offs[1][i]], [offs[2][i]]], dtype=torch.long)
             for i, g in enumerate(graphs)], dim=1))
    return out


def batch2sparse(batch, keys: List[str] = [""]):
    """Wrap the concatenated arrays of a batch into SparseTensors (reference SpData.py:80-112):
    ``batch.A`` (N x N) and, per key, ``batch.X<key>`` with the summed tuple shape."""
    n = batch.num_nodes
    ea = batch.edge_attr
    batch.A = SparseTensor(batch.edge_index, ea, [n, n] if ea is None else [n, n] + list(ea.shape[1:]),
                           is_coalesced=True)
    for key in keys:
        total = getattr(batch, f"tupleshape{key}").sum(dim=0).tolist()
        tupleid, tuplefeat = getattr(batch, f"tupleid{key}"), getattr(batch, f"tuplefeat{key}")
        X = SparseTensor(tupleid, tuplefeat,
                         shape=total if tuplefeat is None else total + list(tuplefeat.shape[1:]),
                         is_coalesced=True)
        setattr(batch, f"X{key}", X)
    return batch
