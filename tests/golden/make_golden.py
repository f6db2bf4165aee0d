"""Generate the golden vectors in this directory from the REAL reference.

Run once in the build container (the reference is Python and importable there):

    python tests/golden/make_golden.py

It imports ``/root/reference/pygho`` (read-only, never copied), runs the reference's
own ``pygho.backend`` / ``pygho.honn`` code on seeded inputs and stores inputs and
outputs as small ``.npz`` files.  The GPU box has no ``/root/reference``; tests only
read the ``.npz`` files.  Nothing in the product imports this script.
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

# pygho.honn.Conv imports torch_geometric.nn.HeteroLinear (used by SUNConv only)
tg, tgnn = types.ModuleType("torch_geometric"), types.ModuleType("torch_geometric.nn")
tgnn.HeteroLinear = type("HeteroLinear", (torch.nn.Module,), {})
tg.nn = tgnn
sys.modules.update({"torch_geometric": tg, "torch_geometric.nn": tgnn})

from pygho import SparseTensor, MaskedTensor  # noqa: E402
from pygho.backend import SpTensor as RSp  # noqa: E402
from pygho.backend import Spspmm as RSS  # noqa: E402
from pygho.backend.Spmm import spmm as rspmm  # noqa: E402
from pygho.backend.Mamamm import mamamm as rmamamm  # noqa: E402
from pygho.backend.utils import torch_scatter_reduce as rscatter  # noqa: E402
from pygho.honn import Conv as RConv  # noqa: E402
from pygho.honn.TensorOp import OpPoolingSubg2D, OpPoolingSubg3D  # noqa: E402

from pygho_b200.hodata.synthetic import make_batch  # noqa: E402

T = torch.from_numpy
torch.set_num_threads(1)      # bit-reproducible reference outputs whichever entry point is used


def npy(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def save(name, **arrs):
    path = os.path.join(os.environ.get("PYGHO_GOLDEN_OUT", HERE), name + ".npz")
    np.savez_compressed(path, **{k: npy(v) for k, v in arrs.items()})
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def rand_sparse(shape, density, gen):
    m = torch.rand(shape, generator=gen) < density
    ind = m.nonzero().t().contiguous()
    return ind


def canon(plan):
    p = npy(plan)
    return p[:, np.lexsort((p[2], p[1], p[0]))]


def g_hash():
    gen = torch.Generator().manual_seed(1)
    out = {}
    for sd, shape in ((2, (300, 200)), (3, (13, 11, 7)), (5, (2, 3, 7, 11, 13))):
        ind = torch.stack([torch.randint(0, s, (41,), generator=gen) for s in shape])
        h = RSp.indicehash(ind)
        out[f"ind{sd}"] = ind
        out[f"hash{sd}"] = h
        out[f"dec{sd}"] = RSp.decodehash(h, sd)
        ds = torch.LongTensor(shape)
        ht = RSp.indicehash_tight(ind, ds)
        out[f"shape{sd}"] = ds
        out[f"tight{sd}"] = ht
        out[f"tdec{sd}"] = RSp.decodehash_tight(ht, ds)
    # the literal known-answer vector of tests/test_backend_sparse.py:94-99
    ptr = torch.tensor([0, 4, 4, 7, 8, 11, 11, 11, 16])
    out["ptr"] = ptr
    out["ptr2batch"] = RSS.ptr2batch(ptr, 16)
    deg = torch.tensor([3, 0, 2, 5, 0, 1])
    out["deg"] = deg
    out["deg2batch"] = RSS.deg2batch(deg, 11)
    save("hash", **out)


def g_coalesce():
    gen = torch.Generator().manual_seed(2)
    out = {}
    n, m, l, nnz, d = 2, 3, 5, 61, 7
    ind = torch.stack([torch.randint(0, s, (nnz,), generator=gen) for s in (n, m, l)])
    val = torch.randn((nnz, d), generator=gen)
    out["ind"], out["val"] = ind, val
    for red in ("sum", "mean", "max", "min"):
        ci, cv = RSp.coalesce(ind, val, red)
        out[f"ind_{red}"], out[f"val_{red}"] = ci, cv
    ival = torch.randint(-9, 9, (nnz,), generator=gen)
    out["ival"] = ival
    for red in ("sum", "mean", "max", "min"):
        _, cv = RSp.coalesce(ind, ival, red)
        out[f"ival_{red}"] = cv
    # scatter with empty rows
    src = torch.randn((37, 3, 2), generator=gen)
    idx = torch.randint(0, 50, (37,), generator=gen)
    out["s_src"], out["s_idx"] = src, idx
    for red in ("sum", "mean", "max", "min"):
        out[f"s_{red}"] = rscatter(0, src, idx, 50, red)
    save("coalesce", **out)


def g_plans():
    gen = torch.Generator().manual_seed(3)
    out = {}
    # 2-D x 2-D, the recipe of tests/test_backend_sparse.py:101-143
    i1 = rand_sparse((30, 20), 0.15, gen)
    i2 = rand_sparse((20, 40), 0.15, gen)
    for tag, (d1, d2, a, b) in {"mm10": (1, 0, i1, i2), "mm01": (0, 1, i2, i1),
                                "mm11": (1, 1, i1, i2.flip(0)),
                                "mm00": (0, 0, i1.flip(0), i2)}.items():
        tar, bcd = RSS.spspmm_ind(a, d1, b, d2)
        out[f"{tag}_i1"], out[f"{tag}_i2"] = a, b
        out[f"{tag}_dims"] = np.array([d1, d2])
        out[f"{tag}_tar"], out[f"{tag}_bcd"] = tar, canon(bcd)
        # random target pattern: half taken from the product, half random
        pick = tar[:, torch.randperm(tar.shape[1], generator=gen)[:tar.shape[1] // 2]]
        rnd = torch.stack([torch.randint(0, int(tar[r].max()) + 1, (25,), generator=gen)
                           for r in range(tar.shape[0])])
        tgt = RSp.decodehash(torch.unique(RSp.indicehash(torch.cat([pick, rnd], 1))),
                             tar.shape[0])
        out[f"{tag}_tgt"] = tgt
        out[f"{tag}_b2a"] = RSS.spsphadamard_ind(tgt, tar)
        out[f"{tag}_acd"] = canon(RSS.filterind(tgt, tar, bcd))
    # 3-D x 3-D (tests/test_backend_sparse.py:162-207) and 3-D x 2-D (I2 key)
    j1 = rand_sparse((13, 11, 5), 0.3, gen)
    j2 = rand_sparse((7, 11, 13), 0.3, gen)
    tar, bcd = RSS.spspmm_ind(j1, 1, j2, 1)
    out["t33_i1"], out["t33_i2"], out["t33_dims"] = j1, j2, np.array([1, 1])
    out["t33_tar"], out["t33_bcd"] = tar, canon(bcd)
    j3 = rand_sparse((9, 9, 9), 0.2, gen)
    a2 = rand_sparse((9, 9), 0.3, gen)
    tar, bcd = RSS.spspmm_ind(j3, 2, a2, 0)
    out["t32_i1"], out["t32_i2"], out["t32_dims"] = j3, a2, np.array([2, 0])
    out["t32_tar"], out["t32_bcd"] = tar, canon(bcd)
    out["t32_acd"] = canon(RSS.filterind(j3, tar, bcd))
    save("plans", **out)


def batch_tensors(hb):
    ei, tid = T(hb.edge_index), T(hb.tupleid)
    return ei, tid


def g_spspmm():
    """Value ops on a small molecule-shaped batch, every aggregation."""
    gen = torch.Generator().manual_seed(4)
    hb = make_batch(4, seed=11)
    ei, tid = batch_tensors(hb)
    N, d = hb.num_nodes, 8
    out = {"edge_index": ei, "tupleid": tid, "N": np.array(N)}
    Av = torch.randn((ei.shape[1], d), generator=gen)
    Xv = torch.randn((tid.shape[1], d), generator=gen)
    out["Av"], out["Xv"] = Av, Xv
    A = SparseTensor(ei, Av, (N, N, d), True)
    X = SparseTensor(tid, Xv, (N, N, d), True)
    keys = {"XA": (X, 1, A, 0), "AX": (A, 1, X, 0), "XX": (X, 1, X, 0)}
    for tag, (P, d1, Q, d2) in keys.items():
        acd = RSS.filterind(tid, *RSS.spspmm_ind(P.indices, d1, Q.indices, d2))
        out[f"{tag}_acd"] = canon(acd)
        for aggr in ("sum", "mean", "max", "min"):
            out[f"{tag}_{aggr}"] = RSS.spspmm(P, d1, Q, d2, aggr, acd=acd, tar_ind=tid).values
    # operands without values count as 1 (Spspmm.py:309-314)
    acd = T(out["XA_acd"])
    Aone = SparseTensor(ei, None, (N, N), True)
    out["XA_sum_noB"] = RSS.spspmm(X, 1, Aone, 0, "sum", acd=acd, tar_ind=tid).values
    # unfiltered product (tar_ind = the product's own pattern)
    tar, bcd = RSS.spspmm_ind(tid, 1, ei, 0)
    out["XA_full_tar"], out["XA_full_bcd"] = tar, canon(bcd)
    out["XA_full_sum"] = RSS.spspmm(X, 1, A, 0, "sum", acd=bcd, tar_ind=tar).values
    # hadamard
    sub = tid[:, torch.rand(tid.shape[1], generator=gen) < 0.6]
    Yv = torch.randn((sub.shape[1], d), generator=gen)
    H = RSS.spsphadamard(X, SparseTensor(sub, Yv, (N, N, d), True))
    out["had_ind2"], out["had_val2"] = sub, Yv
    out["had_ind"], out["had_val"] = H.indices, H.values
    # spmm, both contraction dims, scalar-valued and vector-valued A
    x = torch.randn((N, d), generator=gen)
    out["x"] = x
    A1 = SparseTensor(ei, Av[:, :1].contiguous(), (N, N, 1), True)
    for aggr in ("sum", "mean", "max"):
        out[f"spmm1_{aggr}"] = rspmm(A, 1, x, aggr)
        out[f"spmm0_{aggr}"] = rspmm(A, 0, x, aggr)
    out["spmm1_scalar"] = rspmm(A1, 1, x, "sum")
    out["spmm1_noval"] = rspmm(Aone, 1, x, "sum")
    # pooling / unpooling (SpTensor.py:368-476)
    for aggr in ("sum", "mean", "max"):
        out[f"pool1_{aggr}"] = getattr(X, aggr)([1])
        out[f"pool0_{aggr}"] = getattr(X, aggr)([0])
    out["unpool0"] = X.unpooling_fromdense1dim(0, x).values
    out["unpool1"] = X.unpooling_fromdense1dim(1, x).values
    out["batch"] = T(hb.batch)
    out["readout_sum"] = rscatter(0, x, T(hb.batch), hb.num_graphs, "sum")
    save("spspmm", **out)


def g_3d():
    """3-D tuples (I2 shape): spspmm over dim 2, pooling to sparse."""
    gen = torch.Generator().manual_seed(5)
    hb = make_batch(2, seed=5, tuples="i2")
    ei, tid = batch_tensors(hb)
    N, d = hb.num_nodes, 4
    Av = torch.randn((ei.shape[1], d), generator=gen)
    Xv = torch.randn((tid.shape[1], d), generator=gen)
    A = SparseTensor(ei, Av, (N, N, d), True)
    X = SparseTensor(tid, Xv, (N, N, N, d), True)
    acd = RSS.filterind(tid, *RSS.spspmm_ind(tid, 2, ei, 0))
    out = {"edge_index": ei, "tupleid": tid, "N": np.array(N), "Av": Av, "Xv": Xv,
           "acd": canon(acd)}
    for aggr in ("sum", "max"):
        out[f"mp_{aggr}"] = RSS.spspmm(X, 2, A, 0, aggr, acd=acd, tar_ind=tid).values
    for aggr in ("sum", "mean", "max"):
        P = getattr(X, aggr)([2], return_sparse=True)
        out[f"pool2s_ind_{aggr}"], out[f"pool2s_val_{aggr}"] = P.indices, P.values
        out[f"pool12_{aggr}"] = getattr(X, aggr)([1, 2])
        out[f"pool2_{aggr}"] = getattr(X, aggr)([2])
    P = X.sum([2], return_sparse=True)
    out["unpool_sp"] = P.unpooling([2], X).values
    save("tuples3d", **out)


def g_masked():
    gen = torch.Generator().manual_seed(6)
    b, n, d = 3, 7, 5
    sizes = torch.tensor([7, 4, 5])
    ar = torch.arange(n)
    mask = (ar[None, :, None] < sizes[:, None, None]) & (ar[None, None, :] < sizes[:, None, None])
    A = torch.randn((b, n, n, d), generator=gen) * mask.unsqueeze(-1)
    B = torch.randn((b, n, n, d), generator=gen) * mask.unsqueeze(-1)
    out = {"A": A, "B": B, "mask": mask, "sizes": sizes}
    MA, MB = MaskedTensor(A, mask), MaskedTensor(B, mask)
    for d1 in (1, 2):
        for d2 in (1, 2):
            out[f"mm_{d1}{d2}"] = rmamamm(MA, d1, MB, d2, mask).data * mask.unsqueeze(-1)
    # pooling: pads are zero in the inputs, compare at valid rows (Q1)
    for aggr in ("sum", "mean", "max"):
        for dims in ((1,), (2,), (1, 2)):
            r = getattr(MA, aggr)(list(dims))
            tag = "".join(map(str, dims))
            out[f"pool{tag}_{aggr}"] = r.data
            out[f"pool{tag}_mask"] = r.mask
    # true masked minimum via torch.masked (the reference's min is broken, Q2)
    vd = torch.masked.masked_tensor(A, mask.unsqueeze(-1).expand_as(A).contiguous())
    mn = vd.amin(dim=2)
    out["pool2_min"] = torch.where(mn.get_mask(), mn.get_data(), torch.zeros(()))
    # fill (tests/test_backend_masked.py:50-59) and filterinf (:45-48)
    out["fill1024"] = MaskedTensor(A, mask, padvalue=torch.inf).fill_masked(1024)
    from pygho.backend.MaTensor import filterinf
    fi = torch.tensor([-torch.inf, 0, torch.inf, 1, 2, -torch.inf, 3])
    out["filterinf_in"], out["filterinf_out"] = fi, filterinf(fi)
    save("masked", **out)


def g_conv():
    """One layer of every in-scope conv, forward values and parameter gradients."""
    hb = make_batch(3, seed=21)
    ei, tid = batch_tensors(hb)
    N, d = hb.num_nodes, 8
    gen = torch.Generator().manual_seed(7)
    Av = torch.randn((ei.shape[1], d), generator=gen)
    Xv = torch.randn((tid.shape[1], d), generator=gen)
    out = {"edge_index": ei, "tupleid": tid, "N": np.array(N), "Av": Av, "Xv": Xv}
    mlp = {"numlayer": 2, "tailact": True, "norm": "bn", "act": "silu", "dp": 0.0}
    datadict = {}
    for key, (i1, d1, i2, d2) in {"X___X___1___A___0": (tid, 1, ei, 0),
                                  "X___A___1___X___0": (ei, 1, tid, 0),
                                  "X___X___1___X___0": (tid, 1, tid, 0)}.items():
        datadict[key + "___acd"] = RSS.filterind(tid, *RSS.spspmm_ind(i1, d1, i2, d2))
        out[key + "___acd"] = canon(datadict[key + "___acd"])
    convs = {"NGNN": lambda: RConv.NGNNConv(d, d, "sum", "SS", dict(mlp)),
             "SSWL": lambda: RConv.SSWLConv(d, d, "sum", "SS", dict(mlp)),
             "SSWLmax": lambda: RConv.SSWLConv(d, d, "max", "SS", dict(mlp)),
             "DSSGNN": lambda: RConv.DSSGNNConv(d, d, "sum", "sum", "mean", "SS", dict(mlp)),
             "PPGN": lambda: RConv.PPGNConv(d, d, "sum", "SS", dict(mlp))}
    for name, fn in convs.items():
        torch.manual_seed(100)
        conv = fn()
        A = SparseTensor(ei, Av.clone(), (N, N, d), True)
        xv = Xv.clone().requires_grad_(True)
        X = SparseTensor(tid, xv, (N, N, d), True)
        Y = conv(A, X, datadict)
        loss = (Y.values ** 2).mean()
        loss.backward()
        for k, v in conv.state_dict().items():
            out[f"{name}.sd.{k}"] = v
        for k, p in conv.named_parameters():
            out[f"{name}.grad.{k}"] = p.grad
        out[f"{name}.out"] = Y.values
        out[f"{name}.gradX"] = xv.grad
    save("conv", **out)


def g_conv_i2():
    """I2Conv (Conv.py:107-147) on 3-D tuples with the read-out chain of example/zinc.py:258
    (OpPoolingSubg3D -> OpPoolingSubg2D) and the 3-D tupleinit of zinc.py:270-273 (including its
    use of X.indices[1] for the third factor): forward values, input and parameter gradients."""
    hb = make_batch(3, seed=23, tuples="i2")
    ei, tid = batch_tensors(hb)
    N, d = hb.num_nodes, 8
    gen = torch.Generator().manual_seed(8)
    Av = torch.randn((ei.shape[1], d), generator=gen)
    Xv = torch.randn((tid.shape[1], d), generator=gen)
    xn = torch.randn((N, d), generator=gen)
    out = {"edge_index": ei, "tupleid": tid, "N": np.array(N), "Av": Av, "Xv": Xv, "xn": xn}
    mlp = {"numlayer": 2, "tailact": True, "norm": "bn", "act": "silu", "dp": 0.0}
    key = "X___X___2___A___0"
    acd = RSS.filterind(tid, *RSS.spspmm_ind(tid, 2, ei, 0))
    datadict = {key + "___acd": acd}
    out[key + "___acd"] = canon(acd)
    for name, aggr, pool in (("I2", "sum", "mean"), ("I2max", "max", "max")):
        torch.manual_seed(101)
        conv = RConv.I2Conv(d, d, aggr, "SS", dict(mlp))
        lins = torch.nn.ModuleList([torch.nn.Linear(d, d) for _ in range(3)])
        lpool = torch.nn.Sequential(OpPoolingSubg3D("S", pool), OpPoolingSubg2D("S", pool))
        A = SparseTensor(ei, Av.clone(), (N, N, d), True)
        xv = Xv.clone().requires_grad_(True)
        xnode = xn.clone().requires_grad_(True)
        X = SparseTensor(tid, xv, (N, N, N, d), True)
        X = X.tuplewiseapply(lambda val: lins[0](xnode)[X.indices[0]] * lins[1](xnode)[X.indices[1]]
                             * lins[2](xnode)[X.indices[1]] * val)
        Y = conv(A, X, datadict)
        Z = X.add(Y, True)
        h = lpool(Z)
        loss = (h ** 2).mean() + (Y.values ** 2).mean()
        loss.backward()
        for k, v in conv.state_dict().items():
            out[f"{name}.sd.{k}"] = v
        for k, p in conv.named_parameters():
            out[f"{name}.grad.{k}"] = p.grad
        for i in range(3):
            out[f"{name}.init{i}.weight"], out[f"{name}.init{i}.bias"] = lins[i].weight, lins[i].bias
            out[f"{name}.init{i}.gweight"], out[f"{name}.init{i}.gbias"] = lins[i].weight.grad, lins[i].bias.grad
        out[f"{name}.out"] = Y.values
        out[f"{name}.readout"] = h
        out[f"{name}.gradX"] = xv.grad
        out[f"{name}.gradx"] = xnode.grad
    save("conv_i2", **out)


def g_conv_dd():
    """PPGNConv in DD mode (Conv.py:200-236, mamamm) on graphs of EQUAL size (full masks): the
    one case in which the reference's never-filled pads (MaTensor.py:103-118) cannot
    leak into the result, so its output pins the dense layer's intended semantics."""
    gen = torch.Generator().manual_seed(9)
    b, n, d = 3, 6, 8
    mask = torch.ones((b, n, n), dtype=torch.bool)
    Xd = torch.randn((b, n, n, d), generator=gen)
    mlp = {"numlayer": 2, "tailact": True, "norm": "bn", "act": "silu", "dp": 0.0}
    torch.manual_seed(102)
    conv = RConv.PPGNConv(d, d, "sum", "DD", dict(mlp))
    xd = Xd.clone().requires_grad_(True)
    X = MaskedTensor(xd, mask)
    Y = conv(None, X, {})
    (Y.data ** 2).mean().backward()
    out = {"X": Xd, "mask": mask, "out": Y.data, "gradX": xd.grad}
    for k, v in conv.state_dict().items():
        out[f"sd.{k}"] = v
    for k, p in conv.named_parameters():
        out[f"grad.{k}"] = p.grad
    save("conv_dd", **out)


def golden_spmamm():
    """Reference spmamm (backend/Spmamm.py) on the inputs it can run: its masked_fill call (:60)
    only broadcasts when there are NO dense dims (mask (nnz, k) against mult (nnz, k)), and its
    result is dropped, which is invisible when B's pads already hold the neutral value of the
    aggregation.  Scalar features, one 'other' masked dim."""
    from pygho.backend.Spmamm import spmamm as rspmamm
    gen = torch.Generator().manual_seed(11)
    b, n, k = 3, 7, 5
    sizes = torch.tensor([7, 4, 6])
    adj = (torch.rand((b, n, n), generator=gen) < 0.35)
    valid = torch.arange(n)[None, :] < sizes[:, None]
    adj &= valid[:, :, None] & valid[:, None, :]
    ind = adj.nonzero().t().contiguous()
    aval = torch.rand((ind.shape[1],), generator=gen) + 0.5            # positive: -inf stays -inf
    out = {"ind": ind, "aval": aval, "shape": torch.tensor([b, n, n])}
    for dim2, tag in ((1, "d1"), (2, "d2")):
        bmask = (valid[:, :, None] & (torch.arange(k)[None, None, :] < 4)) if dim2 == 1 else \
            ((torch.arange(k)[None, :, None] < 4) & valid[:, None, :]).expand(b, k, n)
        bmask = bmask.contiguous()
        data = torch.randn(tuple(bmask.shape), generator=gen)
        out[f"{tag}_data"], out[f"{tag}_mask"] = data * bmask, bmask
        for aggr, pad in (("sum", 0.0), ("max", float("-inf"))):
            filled = torch.where(bmask, data, torch.full_like(data, pad))
            for dim1 in (1, 2):
                B = MaskedTensor(filled, bmask, pad, True)
                r = rspmamm(SparseTensor(ind, aval, (b, n, n), True), dim1, B, dim2, None, aggr)
                rd = torch.where(r.mask, r.data, torch.zeros_like(r.data))
                out[f"{tag}_{aggr}_dim{dim1}"] = rd
    save("spmamm", **out)


if __name__ == "__main__" and "spmamm" in sys.argv[1:]:
    golden_spmamm()
    sys.exit(0)

if __name__ == "__main__" and "conv_i2" in sys.argv[1:]:
    g_conv_i2()
    sys.exit(0)

if __name__ == "__main__" and "conv_dd" in sys.argv[1:]:
    g_conv_dd()
    sys.exit(0)


if __name__ == "__main__":
    torch.set_num_threads(1)
    g_hash()
    g_coalesce()
    g_plans()
    g_spspmm()
    g_3d()
    g_masked()
    g_conv()
    g_conv_i2()
    g_conv_dd()
