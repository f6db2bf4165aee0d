"""GPU parity of the device-side batch preparation (hodata.cu through the C ABI) against the
numpy oracle and the reference golden vectors: bit-exact (integer / index work)."""
import numpy as np
import pytest
import torch

from oracle import hodata_oracle as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda", 0)


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


@pytest.mark.parametrize("hop", [1, 2, 3, 5])
def test_khop_sampler_matches_oracle(dev, hop):
    from pygho_b200.hodata.SpTupleSampler import KhopSampler
    from pygho_b200.hodata.synthetic import make_batch
    hb = make_batch(48, seed=11, hop=3)
    X = KhopSampler(_t(hb.edge_index, dev), hb.node_ptr, hop)
    tid, feat = H.khop_sampler_batch(hb.edge_index, hb.node_ptr, hop)
    assert np.array_equal(X.indices.cpu().numpy(), tid)
    assert np.array_equal(X.values.cpu().numpy(), feat)
    if hop == 3:                                          # the generator's own host sampler
        assert np.array_equal(tid, hb.tupleid) and np.array_equal(feat, hb.tuplefeat)
    # device ptr, shuffled (ungrouped) edges
    perm = np.random.default_rng(0).permutation(hb.edge_index.shape[1])
    X2 = KhopSampler(_t(hb.edge_index[:, perm], dev), _t(hb.node_ptr, dev), hop, grouped=False)
    assert torch.equal(X2.indices, X.indices) and torch.equal(X2.values, X.values)


def test_khop_reference_golden_and_directed(dev, golden):
    from pygho_b200.hodata.SpTupleSampler import KhopSampler
    g = golden("hodata")
    for hop in (2, 3):
        for gi in range(6):
            n = int(g[f"g{gi}_n"])
            X = KhopSampler(_t(g[f"g{gi}_edge_index"], dev), [0, n], hop)
            assert np.array_equal(X.indices[1].cpu().numpy(), g[f"khop{hop}_g{gi}_subset"])
            assert np.array_equal(X.values.cpu().numpy(), g[f"khop{hop}_g{gi}_dist"])
            assert np.array_equal(X.indices[0].cpu().numpy(),
                                  np.repeat(np.arange(n), g[f"khop{hop}_g{gi}_len"]))
    # directed edges: the search runs target -> source like the reference
    X = KhopSampler(_t(g["dir_edge_index"], dev), [0, 12], 2, grouped=False)
    assert np.array_equal(X.indices[1].cpu().numpy(), g["dir_subset"])
    assert np.array_equal(X.values.cpu().numpy(), g["dir_dist"])


def test_khop_large_and_degenerate_graphs(dev):
    from pygho_b200.hodata.SpTupleSampler import KhopSampler
    rng = np.random.default_rng(1)
    # one 700-node graph (22 bit words per row), one isolated node, one 33-node path
    n_big = 700
    und = np.stack([rng.integers(0, n_big, 1500), rng.integers(0, n_big, 1500)])
    und = und[:, und[0] != und[1]]
    ei_big = np.unique(np.concatenate([und, und[::-1]], axis=1), axis=1)
    path = np.stack([np.arange(32), np.arange(1, 33)])
    ei_path = np.concatenate([path, path[::-1]], axis=1) + n_big + 1
    ei = np.concatenate([ei_big, ei_path], axis=1)
    node_ptr = np.array([0, n_big, n_big + 1, n_big + 34])
    for hop in (2, 40):
        X = KhopSampler(_t(ei, dev), node_ptr, hop, grouped=False)
        tid, feat = H.khop_sampler_batch(ei, node_ptr, hop)
        assert np.array_equal(X.indices.cpu().numpy(), tid)
        assert np.array_equal(X.values.cpu().numpy(), feat)
    # the isolated node is its own (only) tuple
    iso = (X.indices[0] == n_big).nonzero().flatten()
    assert iso.numel() == 1 and int(X.indices[1, iso]) == n_big and int(X.values[iso]) == 0
    # no edges at all
    X = KhopSampler(torch.zeros((2, 0), dtype=torch.int64, device=dev), [0, 3, 5], 2)
    assert np.array_equal(X.indices.cpu().numpy(), np.stack([np.arange(5), np.arange(5)]))


def test_khop_full_size_properties(dev):
    """B=1024 (bench size): sortedness, symmetry, diagonal, counts -- size-independent checks."""
    from pygho_b200.hodata.SpTupleSampler import KhopSampler
    from pygho_b200.hodata.synthetic import make_batch
    hb = make_batch(1024, seed=0, hop=3)
    X = KhopSampler(_t(hb.edge_index, dev), hb.node_ptr, 3)
    assert np.array_equal(X.indices.cpu().numpy(), hb.tupleid)
    assert np.array_equal(X.values.cpu().numpy(), hb.tuplefeat)
    key = X.indices[0] * hb.num_nodes + X.indices[1]
    assert bool((key[1:] > key[:-1]).all())
    keyT = torch.sort(X.indices[1] * hb.num_nodes + X.indices[0]).values
    assert torch.equal(key, keyT)                                   # undirected -> symmetric
    assert int((X.values == 0).sum()) == hb.num_nodes               # exactly the diagonal
    assert int((X.values == 1).sum()) == hb.edge_index.shape[1]     # exactly the edges


def test_i2_sampler_matches_oracle_and_golden(dev, golden):
    from pygho_b200.hodata.SpTupleSampler import I2Sampler
    from pygho_b200.hodata.synthetic import make_batch
    g = golden("hodata")
    for gi in range(3):
        n = int(g[f"g{gi}_n"])
        X = I2Sampler(_t(g[f"g{gi}_edge_index"], dev), [0, n], 3)
        assert np.array_equal(X.indices[2].cpu().numpy(), g[f"i2_g{gi}_subset"])
        assert np.array_equal(X.values.cpu().numpy(), g[f"i2_g{gi}_feat"])
    hb = make_batch(12, seed=8, hop=2, tuples="i2")
    X = I2Sampler(_t(hb.edge_index, dev), hb.node_ptr, 2)
    assert np.array_equal(X.indices.cpu().numpy(), hb.tupleid)
    assert np.array_equal(X.values.cpu().numpy(), hb.tuplefeat)
    assert X.shape[:3] == (hb.num_nodes,) * 3 and X.sparse_dim == 3
    key = (X.indices[0] * hb.num_nodes + X.indices[1]) * hb.num_nodes + X.indices[2]
    assert bool((key[1:] > key[:-1]).all())             # coalesced: strictly increasing


def test_spdsampler_matches_reference_golden(dev, golden):
    from pygho_b200.hodata.MaTupleSampler import spdsampler
    g = golden("hodata")
    eis, ptr = [], [0]
    for gi in range(4):
        eis.append(g[f"g{gi}_edge_index"] + ptr[-1])
        ptr.append(ptr[-1] + int(g[f"g{gi}_n"]))
    X = spdsampler(_t(np.concatenate(eis, axis=1), dev), ptr, 3)
    data, mask = X.data.cpu().numpy(), X.mask.cpu().numpy()
    for gi in range(4):
        n = int(g[f"g{gi}_n"])
        assert np.array_equal(data[gi, :n, :n], g[f"spd_g{gi}"])
        assert mask[gi, :n, :n].all() and mask[gi].sum() == n * n
        assert (data[gi][~mask[gi]] == 0).all()
    X = spdsampler(_t(g["spd_disc_edge_index"], dev), [0, 6], 2, max_num_nodes=8)
    assert X.data.shape == (1, 8, 8)
    got, ref = X.data[0, :6, :6].cpu().numpy(), g["spd_disc"]
    reach = ref >= 0               # unreachable: reference INT64_MIN (bug), ours hop + 1 (Q14)
    assert np.array_equal(got[reach], ref[reach]) and (got[~reach] == 3).all()
    assert np.array_equal(got, H.spd_matrix(g["spd_disc_edge_index"], 6, 2))


def test_dense_layouts_match_reference_golden(dev, golden):
    from pygho_b200.hodata.MaData import to_dense_adj, to_dense_x
    g = golden("hodata")
    mx = to_dense_x(_t(g["dx_x"], dev), _t(g["dx_ptr"], dev))
    assert np.array_equal(mx.data.cpu().numpy(), g["dx_data"])
    assert np.array_equal(mx.mask.cpu().numpy(), g["dx_mask"])
    mx = to_dense_x(_t(g["dx_x"].astype(np.float32), dev), _t(g["dx_ptr"], dev), 7, None, -1.5)
    want, wmask = H.to_dense_x(g["dx_x"].astype(np.float32), g["dx_ptr"], 7, -1.5)
    assert np.array_equal(mx.data.cpu().numpy(), want) and np.array_equal(mx.mask.cpu().numpy(), wmask)
    ma = to_dense_adj(_t(g["da_ei"], dev), _t(g["da_eb"], dev), _t(g["da_ea"], dev), 6, 4)
    assert np.array_equal(ma.data.cpu().numpy(), g["da_data"])
    assert np.array_equal(ma.mask.cpu().numpy(), g["da_mask"])
    ma = to_dense_adj(_t(g["da_ei"], dev), _t(g["da_eb"], dev), _t(g["da_eaf"], dev), 6, 4)
    assert np.array_equal(ma.data.cpu().numpy(), g["da_dataf"])
    # global ids + node_ptr == local ids
    ptr = np.array([0, 6, 12, 18, 24])
    ma2 = to_dense_adj(_t(g["da_ei"] + ptr[g["da_eb"]], dev), _t(g["da_eb"], dev),
                       _t(g["da_ea"], dev), 6, 4, node_ptr=_t(ptr, dev))
    assert np.array_equal(ma2.data.cpu().numpy(), g["da_data"])


def test_ma_datadict_device_padding_matches_host(dev):
    from pygho_b200.hodata.device import ma_datadict
    from pygho_b200.hodata.synthetic import make_batch
    hb = make_batch(16, seed=4)
    dd = ma_datadict(hb, dev)
    sizes = np.diff(hb.node_ptr)
    n = int(sizes.max())
    x, wm = H.to_dense_x(hb.x[:, None], hb.node_ptr, n)
    assert np.array_equal(dd["x"].data.cpu().numpy(), x) and np.array_equal(dd["x"].mask.cpu().numpy(), wm)
    local = lambda idx: idx - hb.node_ptr[hb.batch[idx[0]]]  # noqa: E731
    A, _ = H.to_dense_adj(local(hb.edge_index), hb.batch[hb.edge_index[0]], hb.edge_attr, n, 16)
    assert np.array_equal(dd["A"].data.cpu().numpy(), A)
    Xw, _ = H.to_dense_adj(local(hb.tupleid), hb.batch[hb.tupleid[0]],
                           np.minimum(hb.tuplefeat, 5) + 1, n, 16)
    assert np.array_equal(dd["X"].data.cpu().numpy(), Xw)
    m2 = wm[:, :, None] & wm[:, None, :]
    assert np.array_equal(dd["X"].mask.cpu().numpy(), m2)
    spd = ma_datadict(hb, dev, max_dist=3, tuples="spd")["X"].data.cpu().numpy()
    g0 = H.spd_matrix(hb.edge_index[:, hb.batch[hb.edge_index[0]] == 0], int(sizes[0]), 3)
    assert np.array_equal(spd[0, :sizes[0], :sizes[0]], g0)


def test_to_dense_tuplefeat_batch2dense_and_batch2sparse(dev, golden):
    """``to_dense_tuplefeat`` against the reference's golden output (hodata/MaData.py:152-212);
    ``batch2dense`` / ``batch2sparse`` / ``collate_sparse`` build the reference's batch
    attributes (MaData.py:215-255, SpData.py:56-112) from per-graph arrays."""
    from types import SimpleNamespace
    from pygho_b200.hodata import MaData, SpData
    from pygho_b200.hodata.synthetic import make_graphs, collate
    g = golden("hodata")
    T = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    mt = MaData.to_dense_tuplefeat(T(g["dt_feat"]), T(g["dt_shape"]), T(g["dt_ptr"]))
    assert np.array_equal(mt.mask.cpu().numpy(), g["dt_mask"])
    assert np.array_equal(mt.data.cpu().numpy(), g["dt_data"])
    # batch2dense on the same arrays + the dense-x / adjacency goldens
    batch = SimpleNamespace(x=T(g["dx_x"]), ptr=T(g["dx_ptr"]), edge_index=T(g["da_ei"]),
                            edge_index_batch=T(g["da_eb"]), edge_attr=T(g["da_ea"]),
                            tuplefeat=T(g["dt_feat"]), tupleshape=T(g["dt_shape"]),
                            tuplefeat_ptr=T(g["dt_ptr"]))
    MaData.batch2dense(batch, denseadj=True)
    assert np.array_equal(batch.x.data.cpu().numpy(), g["dx_data"])
    assert np.array_equal(batch.A.data.cpu().numpy(), g["da_data"])
    assert np.array_equal(batch.X.data.cpu().numpy(), g["dt_data"])
    batch2 = SimpleNamespace(x=T(g["dx_x"]), ptr=T(g["dx_ptr"]), edge_index=T(g["da_ei"]),
                             edge_index_batch=T(g["da_eb"]), edge_attr=T(g["da_ea"]),
                             tuplefeat=T(g["dt_feat"]), tupleshape=T(g["dt_shape"]),
                             tuplefeat_ptr=T(g["dt_ptr"]))
    MaData.batch2dense(batch2, denseadj=False)
    A = batch2.A
    assert A.sparse_dim == 3 and tuple(A.shape[:3]) == (4, 6, 6)
    dense = torch.zeros(4, 6, 6, dtype=A.values.dtype, device=dev)
    dense[A.indices[0], A.indices[1], A.indices[2]] = A.values
    assert np.array_equal(dense.cpu().numpy(), g["da_data"])
    # sparse collate: per-graph dicts -> the block-diagonal batch of synthetic.collate
    graphs = make_graphs(4, seed=7)
    hb = collate(graphs)
    dicts = [dict(x=torch.from_numpy(gr.x), edge_index=torch.from_numpy(gr.edge_index),
                  edge_attr=torch.from_numpy(gr.edge_attr), tupleid=torch.from_numpy(gr.tupleid),
                  tuplefeat=torch.from_numpy(gr.tuplefeat),
                  tupleshape=torch.tensor([gr.num_nodes, gr.num_nodes]), y=gr.y) for gr in graphs]
    b = SpData.collate_sparse(dicts)
    assert np.array_equal(b.tupleid.numpy(), hb.tupleid) and np.array_equal(b.edge_index.numpy(), hb.edge_index)
    assert np.array_equal(b.batch.numpy(), hb.batch)
    for k, v in vars(b).items():
        if isinstance(v, torch.Tensor):
            setattr(b, k, v.to(dev))
    SpData.batch2sparse(b)
    assert tuple(b.X.shape) == (hb.num_nodes, hb.num_nodes) and b.X.nnz == hb.tupleid.shape[1]
    assert tuple(b.A.shape) == (hb.num_nodes, hb.num_nodes)


def test_diag_over_a_subset_of_dims(dev):
    """3-D tuples, diagonal over dims (0, 1): ret[i, k] = X[i, i, k] for every k."""
    from pygho_b200 import SparseTensor
    gen = torch.Generator().manual_seed(3)
    n, d = 7, 5
    full = (torch.rand((n, n, n), generator=gen) < 0.4)
    ind = full.nonzero().t().contiguous().to(dev)
    val = torch.randn((ind.shape[1], d), generator=gen).to(dev).requires_grad_(True)
    X = SparseTensor(ind, val, (n, n, n, d), True)
    got = X.diag([0, 1])
    dense = torch.zeros(n, n, n, d, device=dev)
    dense[ind[0], ind[1], ind[2]] = val.detach()
    want = torch.stack([dense[i, i] for i in range(n)])
    assert torch.equal(got, want)
    got.sum().backward()
    on = (ind[0] == ind[1]).float().unsqueeze(1).expand(-1, d)
    assert torch.equal(val.grad, on)
