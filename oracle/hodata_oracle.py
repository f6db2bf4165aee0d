"""CPU restatement (numpy) of the reference's batch-preparation path -- TEST INFRASTRUCTURE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may import
this module; the product (``pygho_b200``) never does.

Parity status: PINNED.  ``tests/golden/make_golden_hodata.py`` runs the reference's own
``k_hop_subgraph`` (hodata/SpTupleSampler.py:12-88), ``spdsampler``
(hodata/MaTupleSampler.py:11-31, with scipy's shortest_path), ``to_dense_x`` / ``to_dense_adj``
/ ``to_dense_tuplefeat`` (hodata/MaData.py:26-212) on seeded graphs and stores inputs and
outputs in ``tests/golden/hodata.npz``; ``tests/test_oracle_golden.py`` checks every function
below against them.  (``KhopSampler`` itself needs PyG's ``Batch`` to collate its per-node
subgraphs and cannot run here; its per-node content -- subset and dist of ``k_hop_subgraph`` --
is what the golden vectors hold, the collate is the concatenation restated below.)
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def k_hop_subgraph(node: int, num_hops: int, edge_index: np.ndarray,
                   num_nodes: int) -> Tuple[np.ndarray, np.ndarray]:
    """(subset, dist) of reference SpTupleSampler.py:12-88 for a single root with the default
    ``flow='source_to_target'``: every hop collects the SOURCES of the edges whose TARGET is in
    the previous hop's set (:57-64), distances are written from the farthest hop down so the
    nearest wins (:66-67), and ``subset`` is the sorted union (:69-72)."""
    src, dst = edge_index[0], edge_index[1]            # :48-51  col, row = edge_index
    subsets = [np.array([node], dtype=np.int64)]
    for _ in range(num_hops):
        node_mask = np.zeros(num_nodes, dtype=bool)
        node_mask[subsets[-1]] = True
        subsets.append(src[node_mask[dst]])
    dist = np.full(num_nodes, num_nodes + 1, dtype=np.int64)
    for h in range(num_hops, -1, -1):
        dist[subsets[h]] = h
    subset = np.unique(np.concatenate(subsets))
    return subset, dist[subset]


def khop_sampler(edge_index: np.ndarray, num_nodes: int, hop: int) -> Tuple[np.ndarray, np.ndarray]:
    """KhopSampler (SpTupleSampler.py:91-126) of ONE graph: tupleid (2, T) = (root, member)
    for every root in node order, tuplefeat (T,) = hop distance."""
    rows, cols, feats = [], [], []
    for i in range(num_nodes):
        subset, dist = k_hop_subgraph(i, hop, edge_index, num_nodes)
        rows.append(np.full(subset.shape[0], i, dtype=np.int64))
        cols.append(subset)
        feats.append(dist)
    if not rows:
        return np.zeros((2, 0), dtype=np.int64), np.zeros((0,), dtype=np.int64)
    return np.stack([np.concatenate(rows), np.concatenate(cols)]), np.concatenate(feats)


def khop_sampler_batch(edge_index: np.ndarray, node_ptr: np.ndarray,
                       hop: int) -> Tuple[np.ndarray, np.ndarray]:
    """The sampler applied to every graph of a block-diagonal batch (global ids in, global
    ids out): per-graph ``khop_sampler`` + the node offsets of ``SpHoData.__inc__``
    (hodata/SpData.py:60-77)."""
    tids, feats = [], []
    graph_of = np.searchsorted(node_ptr[1:], edge_index[0], side="right")
    for g in range(node_ptr.shape[0] - 1):
        n0, n1 = int(node_ptr[g]), int(node_ptr[g + 1])
        local = edge_index[:, graph_of == g] - n0
        tid, feat = khop_sampler(local, n1 - n0, hop)
        tids.append(tid + n0)
        feats.append(feat)
    if not tids:
        return np.zeros((2, 0), dtype=np.int64), np.zeros((0,), dtype=np.int64)
    return np.concatenate(tids, axis=1), np.concatenate(feats)


def i2_sampler(edge_index: np.ndarray, num_nodes: int, hop: int) -> Tuple[np.ndarray, np.ndarray]:
    """I2Sampler (SpTupleSampler.py:129-174) of ONE graph: for every directed edge (i, j), in
    edge order, the sorted union of the ``hop``-neighbourhoods of i and j (``k_hop_subgraph``
    with the node pair as roots, :153-157) and the two shortest-path distances of every member
    (:164-167, undirected ``shortest_path``).  tupleid (3, T), tuplefeat (T, 2)."""
    full = spd_matrix(edge_index, num_nodes, num_nodes + 1)
    ids, feats = [], []
    for e in range(edge_index.shape[1]):
        i, j = int(edge_index[0, e]), int(edge_index[1, e])
        si, _ = k_hop_subgraph(i, hop, edge_index, num_nodes)
        sj, _ = k_hop_subgraph(j, hop, edge_index, num_nodes)
        subset = np.union1d(si, sj)
        ids.append(np.stack([np.full_like(subset, i), np.full_like(subset, j), subset]))
        feats.append(np.stack([full[i, subset], full[j, subset]], axis=1))
    if not ids:
        return np.zeros((3, 0), dtype=np.int64), np.zeros((0, 2), dtype=np.int64)
    return np.concatenate(ids, axis=1), np.concatenate(feats)


def spd_matrix(edge_index: np.ndarray, num_nodes: int, hop: int) -> np.ndarray:
    """spdsampler (MaTupleSampler.py:11-31): all-pairs unweighted shortest paths of the
    UNDIRECTED graph (``directed=False``), clamped to ``hop + 1``; (n, n).  Unreachable pairs
    are ``hop + 1`` here (the intended value); the reference turns scipy's ``inf`` into
    INT64_MIN before clamping (:29-30), see tests/test_hodata_oracle.py::test_spd_golden."""
    n = num_nodes
    big = n + 1
    d = np.full((n, n), big, dtype=np.int64)
    d[np.arange(n), np.arange(n)] = 0
    d[edge_index[0], edge_index[1]] = np.minimum(d[edge_index[0], edge_index[1]], 1)
    d[edge_index[1], edge_index[0]] = np.minimum(d[edge_index[1], edge_index[0]], 1)
    for k in range(n):                                   # Floyd-Warshall, n <= a few dozen
        d = np.minimum(d, d[:, [k]] + d[[k], :])
    return np.minimum(d, hop + 1)


def to_dense_x(node_x: np.ndarray, xptr: np.ndarray, max_num_nodes: int = None,
               fill=0) -> Tuple[np.ndarray, np.ndarray]:
    """MaData.py:109-149 with the pads filled (what the MaskedTensor constructor is meant to
    do): (b, n, *dense) and the (b, n) mask."""
    b = xptr.shape[0] - 1
    sizes = np.diff(xptr)
    n = int(sizes.max()) if max_num_nodes is None else max_num_nodes
    out = np.full((b, n) + node_x.shape[1:], fill, dtype=node_x.dtype)
    mask = np.arange(n)[None, :] < sizes[:, None]
    for g in range(b):
        out[g, :sizes[g]] = node_x[xptr[g]:xptr[g + 1]]
    return out, mask


def to_dense_adj(edge_index: np.ndarray, edge_batch: np.ndarray, edge_attr: np.ndarray,
                 max_num_nodes: int, batch_size: int, fill=0) -> Tuple[np.ndarray, np.ndarray]:
    """MaData.py:26-72: ``ret[batch, row, col] = edge_attr``; mask marks the edges."""
    out = np.full((batch_size, max_num_nodes, max_num_nodes) + edge_attr.shape[1:], fill,
                  dtype=edge_attr.dtype)
    mask = np.zeros((batch_size, max_num_nodes, max_num_nodes), dtype=bool)
    out[edge_batch, edge_index[0], edge_index[1]] = edge_attr
    mask[edge_batch, edge_index[0], edge_index[1]] = True
    return out, mask


def to_dense_tuplefeat(tuplefeat: np.ndarray, tupleshape: np.ndarray,
                       ptr: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """MaData.py:152-212 for 2-D tuple shapes: graph g's flattened (n1_g x n2_g) features are
    laid out row-major at ``tuplefeat[ptr[g]:ptr[g+1]]`` and padded to the largest shape."""
    b = tupleshape.shape[0]
    n1, n2 = int(tupleshape[:, 0].max()), int(tupleshape[:, 1].max())
    out = np.zeros((b, n1, n2) + tuplefeat.shape[1:], dtype=tuplefeat.dtype)
    mask = np.zeros((b, n1, n2), dtype=bool)
    for g in range(b):
        a, c = int(tupleshape[g, 0]), int(tupleshape[g, 1])
        out[g, :a, :c] = tuplefeat[ptr[g]:ptr[g + 1]].reshape((a, c) + tuplefeat.shape[1:])
        mask[g, :a, :c] = True
    return out, mask
