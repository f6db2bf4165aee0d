"""Golden vectors for the batch-preparation path, generated from the REAL reference.

    python tests/golden/make_golden_hodata.py        # writes tests/golden/hodata.npz

``pygho.hodata`` needs torch_geometric, which is not installed; the functions on this path
(``k_hop_subgraph``, ``spdsampler``, ``to_dense_x``, ``to_dense_adj``, ``to_dense_tuplefeat``)
only use torch / scipy, so the two module FILES are executed with a stub ``torch_geometric``
that provides the three names their bodies touch (``maybe_num_nodes``,
``to_scipy_sparse_matrix``, a ``Data`` holder).  Nothing of the reference is copied."""
import importlib.util
import os
import sys
import types
import warnings

import numpy as np
import scipy.sparse as ssp
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")


class _Data:                      # attribute bag standing in for torch_geometric.data.Data
    def __init__(self, **kw):
        self.__dict__.update(kw)


def _to_scipy(edge_index, num_nodes=None):
    ei = edge_index.numpy()
    return ssp.coo_matrix((np.ones(ei.shape[1]), (ei[0], ei[1])), shape=(num_nodes, num_nodes))


def _maybe_num_nodes(edge_index, num_nodes=None):
    return int(edge_index.max()) + 1 if num_nodes is None else num_nodes


tg = types.ModuleType("torch_geometric")
tgd = types.ModuleType("torch_geometric.data")
tgu = types.ModuleType("torch_geometric.utils")
tgun = types.ModuleType("torch_geometric.utils.num_nodes")
tgnn = types.ModuleType("torch_geometric.nn")
tgd.Data, tgd.Batch = _Data, _Data
tgu.to_scipy_sparse_matrix, tgu.k_hop_subgraph, tgu.coalesce = _to_scipy, None, None
tgun.maybe_num_nodes = _maybe_num_nodes
tgnn.HeteroLinear = type("HeteroLinear", (torch.nn.Module,), {})
tg.data, tg.utils, tg.nn = tgd, tgu, tgnn
sys.modules.update({"torch_geometric": tg, "torch_geometric.data": tgd, "torch_geometric.utils": tgu,
                    "torch_geometric.utils.num_nodes": tgun, "torch_geometric.nn": tgnn})

import pygho.backend  # noqa: E402,F401  (real package: SpTensor / MaTensor for the relative imports)

# run the module files under their real names without executing pygho/hodata/__init__.py
pkg = types.ModuleType("pygho.hodata")
pkg.__path__ = ["/root/reference/pygho/hodata"]
sys.modules["pygho.hodata"] = pkg


def load(name):
    spec = importlib.util.spec_from_file_location(
        f"pygho.hodata.{name}", f"/root/reference/pygho/hodata/{name}.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod


RSamp, RMaSamp, RMaData = load("SpTupleSampler"), load("MaTupleSampler"), load("MaData")

from pygho_b200.hodata.synthetic import make_graphs  # noqa: E402

out = {}
# ---- k_hop_subgraph per root, two hop values, on seeded molecule-like graphs (+1 directed)
graphs = make_graphs(6, seed=77, hop=3)
for hop in (2, 3):
    for gi, g in enumerate(graphs):
        ei = torch.from_numpy(g.edge_index)
        subs, dists, lens = [], [], []
        for i in range(g.num_nodes):
            subset, _, _, _, dist = RSamp.k_hop_subgraph(i, hop, ei, relabel_nodes=True,
                                                         num_nodes=g.num_nodes)
            subs.append(subset.numpy()); dists.append(dist.numpy()); lens.append(subset.shape[0])
        out[f"khop{hop}_g{gi}_subset"] = np.concatenate(subs)
        out[f"khop{hop}_g{gi}_dist"] = np.concatenate(dists)
        out[f"khop{hop}_g{gi}_len"] = np.array(lens)
        out[f"g{gi}_edge_index"] = g.edge_index
        out[f"g{gi}_n"] = np.array(g.num_nodes)
rng = np.random.default_rng(5)
dir_ei = np.unique(rng.integers(0, 12, size=(2, 30)), axis=1)
dir_ei = dir_ei[:, dir_ei[0] != dir_ei[1]]
subs, dists, lens = [], [], []
for i in range(12):
    subset, _, _, _, dist = RSamp.k_hop_subgraph(i, 2, torch.from_numpy(dir_ei), relabel_nodes=True,
                                                 num_nodes=12)
    subs.append(subset.numpy()); dists.append(dist.numpy()); lens.append(subset.shape[0])
out["dir_edge_index"], out["dir_subset"] = dir_ei, np.concatenate(subs)
out["dir_dist"], out["dir_len"] = np.concatenate(dists), np.array(lens)

# ---- spdsampler (scipy shortest_path inside the reference), incl. a disconnected graph
for gi, g in enumerate(graphs[:4]):
    feat, shape = RMaSamp.spdsampler(_Data(edge_index=torch.from_numpy(g.edge_index),
                                           num_nodes=g.num_nodes), hop=3)
    out[f"spd_g{gi}"] = feat.numpy().reshape(shape)
two = np.array([[0, 1, 1, 2, 3, 4], [1, 0, 2, 1, 4, 3]])          # components {0,1,2} and {3,4}, 5 alone
feat, shape = RMaSamp.spdsampler(_Data(edge_index=torch.from_numpy(two), num_nodes=6), hop=2)
out["spd_disc_edge_index"], out["spd_disc"] = two, feat.numpy().reshape(shape)

# ---- I2Sampler content (SpTupleSampler.py:129-174): the reference's own k_hop_subgraph on every
# node PAIR of an edge + the same scipy shortest_path call it makes; the PyG collate is left out
for gi, g in enumerate(graphs[:3]):
    ei = torch.from_numpy(g.edge_index)
    dist_matrix = torch.from_numpy(ssp.csgraph.shortest_path(
        _to_scipy(ei, g.num_nodes), directed=False, unweighted=True,
        return_predecessors=False)).to(torch.long)
    subs, feats, lens = [], [], []
    for e in range(ei.shape[1]):
        pair = ei[:, e]
        subset, _, _, _, _ = RSamp.k_hop_subgraph(pair, 3, ei, relabel_nodes=True,
                                                  num_nodes=g.num_nodes)
        subs.append(subset.numpy()); lens.append(subset.shape[0])
        feats.append(torch.stack((dist_matrix[pair[0].item()][subset],
                                  dist_matrix[pair[1].item()][subset]), dim=-1).numpy())
    out[f"i2_g{gi}_subset"], out[f"i2_g{gi}_len"] = np.concatenate(subs), np.array(lens)
    out[f"i2_g{gi}_feat"] = np.concatenate(feats)

# ---- to_dense_x / to_dense_adj / to_dense_tuplefeat
gen = torch.Generator().manual_seed(3)
sizes = torch.tensor([4, 1, 6, 3])
xptr = torch.cat([torch.zeros(1, dtype=torch.long), sizes.cumsum(0)])
node_x = torch.randint(0, 20, (int(sizes.sum()), 2), generator=gen)
mx = RMaData.to_dense_x(node_x, xptr)
out["dx_x"], out["dx_ptr"] = node_x.numpy(), xptr.numpy()
out["dx_data"] = torch.where(mx.mask.unsqueeze(-1), mx.data, torch.zeros_like(mx.data)).numpy()
out["dx_mask"] = mx.mask.numpy()
eb = torch.tensor([0, 0, 0, 2, 2, 3, 3, 3])
ei = torch.tensor([[0, 1, 3, 0, 5, 0, 1, 2], [1, 0, 2, 5, 0, 1, 2, 0]])
ea = torch.randint(1, 4, (8,), generator=gen)
ma = RMaData.to_dense_adj(ei, eb, ea, 6, 4)
out["da_ei"], out["da_eb"], out["da_ea"] = ei.numpy(), eb.numpy(), ea.numpy()
out["da_data"], out["da_mask"] = ma.data.numpy(), ma.mask.numpy()
eaf = torch.randn((8, 3), generator=gen)
maf = RMaData.to_dense_adj(ei, eb, eaf, 6, 4)
out["da_eaf"], out["da_dataf"] = eaf.numpy(), maf.data.numpy()
tshape = torch.stack([sizes, sizes], 1)
tptr = torch.cat([torch.zeros(1, dtype=torch.long), (sizes * sizes).cumsum(0)])
tfeat = torch.randint(0, 9, (int(tptr[-1]),), generator=gen)
mt = RMaData.to_dense_tuplefeat(tfeat, tshape, tptr)
out["dt_feat"], out["dt_shape"], out["dt_ptr"] = tfeat.numpy(), tshape.numpy(), tptr.numpy()
out["dt_data"] = torch.where(mt.mask, mt.data, torch.zeros_like(mt.data)).numpy()
out["dt_mask"] = mt.mask.numpy()

path = os.path.join(os.environ.get("PYGHO_GOLDEN_OUT", HERE), "hodata.npz")
np.savez_compressed(path, **out)
print(f"hodata: {os.path.getsize(path) / 1024:.1f} KiB, {len(out)} arrays")
