"""PyG-free synthetic molecule-shaped graph batches (host side, numpy only).

The reference builds its batches with torch_geometric (``hodata/SpData.py:56-112``,
``hodata/SpTupleSampler.py:91-126``); neither PyG nor the ZINC files exist on the
benchmark box, so this module generates graphs with ZINC-like statistics
(about 23 nodes and 50 directed edges per graph, degree <= 4, a few rings) and lays
them out exactly like the reference collate does:

* node ids of graph ``g`` are shifted by the number of nodes before it,
* ``edge_index`` / ``tupleid`` are concatenated along dim 1 (``SpData.py:60-77``),
* tuples are the k-hop pairs ``{(i, j): dist(i, j) <= hop}`` with ``tuplefeat = dist``
  (what ``KhopSampler`` produces, ``SpTupleSampler.py:91-126``), sorted by ``(i, j)``.

Everything here is offline preprocessing (out of the hot path); the index plans
(``<key>___acd``) are built on the device by ``pygho_b200.backend.Spspmm``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

__all__ = [
    "HostGraph", "HostBatch", "zinc_like_graph", "khop_tuples", "i2_tuples",
    "make_graphs", "collate", "make_batch", "graph_from_edges", "sr25_like_graph",
]


@dataclass
class HostGraph:
    """One graph on the host: symmetric, sorted edge list plus its tuple set."""
    num_nodes: int
    x: np.ndarray            # (n,) int64 node labels
    edge_index: np.ndarray   # (2, e) int64, sorted by (row, col), both directions
    edge_attr: np.ndarray    # (e,) int64
    tupleid: np.ndarray      # (sd, t) int64, sorted lexicographically
    tuplefeat: np.ndarray    # (t,) or (t, f) int64
    y: float = 0.0


@dataclass
class HostBatch:
    """Block-diagonal concatenation of graphs (the layout of ``SpData.py:60-112``)."""
    num_graphs: int
    num_nodes: int
    x: np.ndarray
    edge_index: np.ndarray
    edge_attr: np.ndarray
    tupleid: np.ndarray
    tuplefeat: np.ndarray
    batch: np.ndarray        # (N,) graph id of every node
    y: np.ndarray            # (B,) float32
    node_ptr: np.ndarray     # (B+1,)
    edge_ptr: np.ndarray     # (B+1,)
    tuple_ptr: np.ndarray    # (B+1,)
    plans: Dict[str, np.ndarray] = field(default_factory=dict)
    # capacity-padded batches (pygho_b200/static.py): valid (tuples, nodes, graphs), int32
    valid: Optional[np.ndarray] = None

    def nbytes(self) -> int:
        tot = 0
        for v in (self.x, self.edge_index, self.edge_attr, self.tupleid,
                  self.tuplefeat, self.batch, self.y):
            tot += v.nbytes
        for v in self.plans.values():
            tot += v.nbytes
        if self.valid is not None:
            tot += self.valid.nbytes
        return tot


def _bfs_dist(n: int, nbrs: List[List[int]], src: int, cutoff: int) -> np.ndarray:
    dist = np.full(n, -1, dtype=np.int64)
    dist[src] = 0
    frontier = [src]
    d = 0
    while frontier and d < cutoff:
        d += 1
        nxt = []
        for u in frontier:
            for v in nbrs[u]:
                if dist[v] < 0:
                    dist[v] = d
                    nxt.append(v)
        frontier = nxt
    return dist


def _adjacency(n: int, edge_index: np.ndarray) -> List[List[int]]:
    nbrs: List[List[int]] = [[] for _ in range(n)]
    for r, c in edge_index.T.tolist():
        nbrs[r].append(c)
    return nbrs


def graph_from_edges(n: int, und_edges: np.ndarray, rng: np.random.Generator,
                     hop: int = 3, tuples: str = "khop") -> HostGraph:
    """Finish a graph from an undirected edge list ``(m, 2)``."""
    und_edges = np.asarray(und_edges, dtype=np.int64).reshape(-1, 2)
    both = np.concatenate([und_edges, und_edges[:, ::-1]], axis=0)
    key = both[:, 0] * n + both[:, 1]
    key = np.unique(key)
    edge_index = np.stack([key // n, key % n]).astype(np.int64)
    # one attribute per undirected edge, mirrored on both directions
    lo = np.minimum(edge_index[0], edge_index[1])
    hi = np.maximum(edge_index[0], edge_index[1])
    table = rng.integers(1, 4, size=n * n, dtype=np.int64)
    edge_attr = table[lo * n + hi]
    x = rng.integers(0, 28, size=n, dtype=np.int64)
    if tuples == "khop":
        tid, tfeat = khop_tuples(n, edge_index, hop)
    elif tuples == "i2":
        tid, tfeat = i2_tuples(n, edge_index, hop)
    else:
        raise ValueError(f"unknown tuple sampler {tuples}")
    return HostGraph(n, x, edge_index, edge_attr, tid, tfeat,
                     float(rng.standard_normal()))


def zinc_like_graph(rng: np.random.Generator, hop: int = 3,
                    tuples: str = "khop") -> HostGraph:
    """Random molecule-like graph: bounded-degree tree plus 1-3 ring closures."""
    n = int(np.clip(np.rint(rng.normal(23.2, 4.5)), 9, 37))
    deg = np.zeros(n, dtype=np.int64)
    edges: List[Tuple[int, int]] = []
    for v in range(1, n):
        lo = max(0, v - 6)
        cand = [u for u in range(lo, v) if deg[u] < 3]
        if not cand:
            cand = [u for u in range(0, v) if deg[u] < 4] or [v - 1]
        u = int(cand[rng.integers(len(cand))])
        edges.append((u, v))
        deg[u] += 1
        deg[v] += 1
    nbrs: List[List[int]] = [[] for _ in range(n)]
    for u, v in edges:
        nbrs[u].append(v)
        nbrs[v].append(u)
    want = int(rng.integers(1, 4))
    tries = 0
    while want > 0 and tries < 40:
        tries += 1
        u = int(rng.integers(n))
        if deg[u] >= 4:
            continue
        dist = _bfs_dist(n, nbrs, u, 5)
        cand = [v for v in range(n) if 4 <= dist[v] <= 5 and deg[v] < 4]
        if not cand:
            continue
        v = int(cand[rng.integers(len(cand))])
        edges.append((u, v))
        nbrs[u].append(v)
        nbrs[v].append(u)
        deg[u] += 1
        deg[v] += 1
        want -= 1
    return graph_from_edges(n, np.array(edges, dtype=np.int64), rng, hop, tuples)


def khop_tuples(n: int, edge_index: np.ndarray, hop: int) -> Tuple[np.ndarray, np.ndarray]:
    """``{(i, j): dist(i, j) <= hop}`` with the distance as feature (KhopSampler)."""
    nbrs = _adjacency(n, edge_index)
    rows, cols, feats = [], [], []
    for i in range(n):
        dist = _bfs_dist(n, nbrs, i, hop)
        js = np.nonzero(dist >= 0)[0]
        rows.append(np.full(js.shape[0], i, dtype=np.int64))
        cols.append(js.astype(np.int64))
        feats.append(dist[js])
    return (np.stack([np.concatenate(rows), np.concatenate(cols)]),
            np.concatenate(feats))


def i2_tuples(n: int, edge_index: np.ndarray, hop: int) -> Tuple[np.ndarray, np.ndarray]:
    """3-tuples ``(i, j, k)``: for every directed edge (i, j) the union of the
    ``hop``-neighbourhoods of i and j; feature = (dist(i,k), dist(j,k)).
    Mirrors ``I2Sampler`` (``SpTupleSampler.py:129-174``); unreachable distances are
    clamped to ``hop + 1`` so they index an embedding table."""
    nbrs = _adjacency(n, edge_index)
    full = np.stack([_bfs_dist(n, nbrs, i, n) for i in range(n)])
    i0, i1, i2, f = [], [], [], []
    for r, c in edge_index.T.tolist():
        ks = np.nonzero(((full[r] >= 0) & (full[r] <= hop)) |
                        ((full[c] >= 0) & (full[c] <= hop)))[0]
        i0.append(np.full(ks.shape[0], r, dtype=np.int64))
        i1.append(np.full(ks.shape[0], c, dtype=np.int64))
        i2.append(ks.astype(np.int64))
        dr = np.where(full[r][ks] < 0, hop + 1, np.minimum(full[r][ks], hop + 1))
        dc = np.where(full[c][ks] < 0, hop + 1, np.minimum(full[c][ks], hop + 1))
        f.append(np.stack([dr, dc], axis=1))
    tid = np.stack([np.concatenate(i0), np.concatenate(i1), np.concatenate(i2)])
    return tid, np.concatenate(f).astype(np.int64)


# Two Latin squares of order 5 from different main classes: the Cayley table of Z5 and a
# non-group square.  Their Latin-square graphs L3(5) are the strongly regular graphs (25, 12, 5, 6)
# number 1 (= the Paley graph of GF(25)) and number 0 of the reference's sr25 set
# (dataset/sr25/raw/sr251256.g6, 15 graphs; checked with networkx in tests/test_hodata_oracle.py).
_LATIN5 = (
    ((0, 1, 2, 3, 4), (1, 2, 3, 4, 0), (2, 3, 4, 0, 1), (3, 4, 0, 1, 2), (4, 0, 1, 2, 3)),
    ((1, 0, 4, 3, 2), (3, 4, 2, 1, 0), (0, 1, 3, 2, 4), (4, 2, 1, 0, 3), (2, 3, 0, 4, 1)),
)


def sr25_like_graph(rng: np.random.Generator, hop: int = 3, tuples: str = "khop",
                    which: Optional[int] = None) -> HostGraph:
    """A strongly regular graph with the sr25 parameters (25 nodes, 12-regular, lambda 5, mu 6,
    diameter 2) under a random relabelling: cells of a 5 x 5 Latin square, adjacent when they
    share a row, a column or a symbol.  Per graph: 300 directed edges, 625 2-tuples at hop >= 2,
    7500 3-tuples (I2), 7500 / 90 000 plan triples."""
    L = _LATIN5[int(rng.integers(len(_LATIN5))) if which is None else which]
    perm = rng.permutation(25)
    edges = []
    for a in range(25):
        for b in range(a + 1, 25):
            i, j = divmod(a, 5)
            k, l = divmod(b, 5)
            if i == k or j == l or L[i][j] == L[k][l]:
                edges.append((perm[a], perm[b]))
    return graph_from_edges(25, np.array(edges, dtype=np.int64), rng, hop, tuples)


def make_graphs(num: int, seed: int = 0, hop: int = 3, tuples: str = "khop",
                shape: str = "zinc") -> List[HostGraph]:
    rng = np.random.default_rng(seed)
    if shape == "sr25":
        return [sr25_like_graph(rng, hop, tuples) for _ in range(num)]
    if shape != "zinc":
        raise ValueError(f"unknown graph shape {shape}")
    return [zinc_like_graph(rng, hop, tuples) for _ in range(num)]


def collate(graphs: List[HostGraph]) -> HostBatch:
    """Concatenate graphs with the offsets of ``SpHoData.__inc__`` (``SpData.py:60-77``)."""
    B = len(graphs)
    node_ptr = np.zeros(B + 1, dtype=np.int64)
    edge_ptr = np.zeros(B + 1, dtype=np.int64)
    tuple_ptr = np.zeros(B + 1, dtype=np.int64)
    for g, gr in enumerate(graphs):
        node_ptr[g + 1] = node_ptr[g] + gr.num_nodes
        edge_ptr[g + 1] = edge_ptr[g] + gr.edge_index.shape[1]
        tuple_ptr[g + 1] = tuple_ptr[g] + gr.tupleid.shape[1]
    if B == 0:
        raise ValueError("cannot collate an empty list of graphs")
    return HostBatch(
        num_graphs=B,
        num_nodes=int(node_ptr[-1]),
        x=np.concatenate([g.x for g in graphs]),
        edge_index=np.concatenate(
            [g.edge_index + node_ptr[i] for i, g in enumerate(graphs)], axis=1),
        edge_attr=np.concatenate([g.edge_attr for g in graphs]),
        tupleid=np.concatenate(
            [g.tupleid + node_ptr[i] for i, g in enumerate(graphs)], axis=1),
        tuplefeat=np.concatenate([g.tuplefeat for g in graphs]),
        batch=np.repeat(np.arange(B, dtype=np.int64), np.diff(node_ptr)),
        y=np.array([g.y for g in graphs], dtype=np.float32),
        node_ptr=node_ptr, edge_ptr=edge_ptr, tuple_ptr=tuple_ptr,
    )


def make_batch(num_graphs: int, seed: int = 0, hop: int = 3, tuples: str = "khop",
               shape: str = "zinc") -> HostBatch:
    return collate(make_graphs(num_graphs, seed, hop, tuples, shape))
