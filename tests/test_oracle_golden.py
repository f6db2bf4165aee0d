"""Pin the numpy oracle to the real reference (golden vectors made by
tests/golden/make_golden.py from /root/reference) and to the reference's own
known-answer tests."""
import numpy as np
import pytest

from oracle import pygho_oracle as O

RTOL = 1e-5


def close(a, b, rtol=RTOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    scale = max(1.0, float(np.abs(b).max()) if b.size else 1.0)
    assert np.abs(a - b).max(initial=0.0) <= rtol * scale, np.abs(a - b).max()


def test_hash_golden(golden):
    g = golden("hash")
    for sd in (2, 3, 5):
        ind = g[f"ind{sd}"]
        assert np.array_equal(O.indicehash(ind), g[f"hash{sd}"])
        assert np.array_equal(O.decodehash(g[f"hash{sd}"], sd), g[f"dec{sd}"])
        assert np.array_equal(O.decodehash(O.indicehash(ind), sd), ind)
        assert np.array_equal(O.indicehash_tight(ind, g[f"shape{sd}"]), g[f"tight{sd}"])
        assert np.array_equal(O.decodehash_tight(g[f"tight{sd}"], g[f"shape{sd}"]), ind)


def test_hash_tight_horner_kat():
    # reference tests/test_backend_sparse.py:35-48 -- Horner evaluation
    rng = np.random.default_rng(0)
    shape = (2, 3, 7, 11, 13)
    ind = np.stack([rng.integers(0, s, 17) for s in shape])
    horner = (((ind[0] * 3 + ind[1]) * 7 + ind[2]) * 11 + ind[3]) * 13 + ind[4]
    assert np.array_equal(O.indicehash_tight(ind, shape), horner)


def test_hash_keeps_lexicographic_order():
    # reference tests/test_backend_sparse.py:50-60
    rng = np.random.default_rng(1)
    shape = (2, 3, 7, 11, 13)
    ind = np.stack([rng.integers(0, s, 17) for s in shape])
    ind = ind[:, np.lexsort(ind[::-1])]
    assert np.all(np.diff(O.indicehash(ind)) >= 0)


def test_ptr2batch_kat(golden):
    # literal vector of reference tests/test_backend_sparse.py:94-99
    ptr = np.array([0, 4, 4, 7, 8, 11, 11, 11, 16])
    want = np.array([0, 0, 0, 0, 2, 2, 2, 3, 4, 4, 4, 7, 7, 7, 7, 7])
    assert np.array_equal(O.ptr2batch(ptr, 16), want)
    g = golden("hash")
    assert np.array_equal(O.ptr2batch(g["ptr"], 16), g["ptr2batch"])
    assert np.array_equal(O.deg2batch(g["deg"], 11), g["deg2batch"])


def test_coalesce_and_scatter_golden(golden):
    g = golden("coalesce")
    for red in ("sum", "mean", "max", "min"):
        ind, val = O.coalesce(g["ind"], g["val"], red)
        assert np.array_equal(ind, g[f"ind_{red}"])
        close(val, g[f"val_{red}"])
        _, ival = O.coalesce(g["ind"], g["ival"], red)
        assert np.array_equal(ival, g[f"ival_{red}"]), red
        close(O.scatter_reduce(g["s_src"], g["s_idx"], 50, red), g[f"s_{red}"])


def test_create_vs_dense_coalesce():
    # recipe of reference tests/test_backend_sparse.py:62-85 without torch.sparse
    rng = np.random.default_rng(3)
    n, m, l, nnz, d = 2, 3, 5, 23, 7
    ind = np.stack([rng.integers(0, s, nnz) for s in (n, m, l)])
    val = rng.standard_normal((nnz, d)).astype(np.float32)
    dense = np.zeros((n, m, l, d), dtype=np.float64)
    np.add.at(dense, tuple(ind), val)
    cind, cval = O.coalesce(ind, val, "sum")
    assert np.all(np.diff(O.indicehash(cind)) > 0)
    close(cval, dense[tuple(cind)])
    assert cind.shape[1] == int((np.abs(dense).sum(-1) > 0).sum())


def test_plans_golden(golden):
    g = golden("plans")
    for tag in ("mm10", "mm01", "mm11", "mm00", "t33", "t32"):
        d1, d2 = g[f"{tag}_dims"]
        tar, bcd = O.spspmm_ind(g[f"{tag}_i1"], int(d1), g[f"{tag}_i2"], int(d2))
        assert np.array_equal(tar, g[f"{tag}_tar"]), tag
        assert np.array_equal(bcd, g[f"{tag}_bcd"]), tag
        if f"{tag}_tgt" in g:
            assert np.array_equal(O.spsphadamard_ind(g[f"{tag}_tgt"], tar), g[f"{tag}_b2a"])
            assert np.array_equal(O.filterind(g[f"{tag}_tgt"], tar, bcd), g[f"{tag}_acd"])
    tar, bcd = O.spspmm_ind(g["t32_i1"], 2, g["t32_i2"], 0)
    assert np.array_equal(O.filterind(g["t32_i1"], tar, bcd), g["t32_acd"])


def test_spspmm_ind_vs_dense_product():
    # recipe of reference tests/test_backend_sparse.py:101-127: A @ B
    rng = np.random.default_rng(5)
    n, m, l = 30, 20, 40
    A = rng.random((n, m)) * (rng.random((n, m)) < 0.1)
    B = rng.random((m, l)) * (rng.random((m, l)) < 0.1)
    i1, i2 = np.stack(np.nonzero(A)), np.stack(np.nonzero(B))
    tar, bcd = O.spspmm_ind(i1, 1, i2, 0)
    val = O.spspmm(A[tuple(i1)][:, None], B[tuple(i2)][:, None], bcd, tar.shape[1])
    C = A @ B
    assert np.array_equal(tar, np.stack(np.nonzero(C)))
    close(val[:, 0], C[tuple(tar)])


def test_3dmm_vs_einsum():
    # recipe of reference tests/test_backend_sparse.py:162-188
    rng = np.random.default_rng(6)
    n, m, l, k = 13, 5, 7, 11
    A = rng.random((n, k, m)) * (rng.random((n, k, m)) < 0.5)
    B = rng.random((l, k, n)) * (rng.random((l, k, n)) < 0.5)
    i1, i2 = np.stack(np.nonzero(A)), np.stack(np.nonzero(B))
    C = np.einsum("nkm,lkd->nmld", A, B)
    tar, bcd = O.spspmm_ind(i1, 1, i2, 1)
    val = O.spspmm(A[tuple(i1)][:, None], B[tuple(i2)][:, None], bcd, tar.shape[1])
    assert np.array_equal(tar, np.stack(np.nonzero(C)))
    close(val[:, 0], C[tuple(tar)])


def test_value_ops_golden(golden):
    g = golden("spspmm")
    N = int(g["N"])
    ei, tid, Av, Xv, x = g["edge_index"], g["tupleid"], g["Av"], g["Xv"], g["x"]
    ops = {"XA": (Xv, Av, tid, 1, ei, 0), "AX": (Av, Xv, ei, 1, tid, 0), "XX": (Xv, Xv, tid, 1, tid, 0)}
    for tag, (pv, qv, pi, d1, qi, d2) in ops.items():
        acd = O.filterind(tid, *O.spspmm_ind(pi, d1, qi, d2))
        assert np.array_equal(acd, g[f"{tag}_acd"]), tag
        for aggr in ("sum", "mean", "max", "min"):
            close(O.spspmm(pv, qv, acd, tid.shape[1], aggr), g[f"{tag}_{aggr}"])
    close(O.spspmm(Xv, None, g["XA_acd"], tid.shape[1]), g["XA_sum_noB"])
    tar, bcd = O.spspmm_ind(tid, 1, ei, 0)
    assert np.array_equal(tar, g["XA_full_tar"]) and np.array_equal(bcd, g["XA_full_bcd"])
    close(O.spspmm(Xv, Av, bcd, tar.shape[1]), g["XA_full_sum"])
    hi, hv = O.spsphadamard(tid, Xv, g["had_ind2"], g["had_val2"])
    assert np.array_equal(hi, g["had_ind"])
    close(hv, g["had_val"])
    for aggr in ("sum", "mean", "max"):
        close(O.spmm(ei, Av, (N, N), 1, x, aggr), g[f"spmm1_{aggr}"])
        close(O.spmm(ei, Av, (N, N), 0, x, aggr), g[f"spmm0_{aggr}"])
        close(O.sp_pool_dense(tid, Xv, (N, N), [1], aggr), g[f"pool1_{aggr}"])
        close(O.sp_pool_dense(tid, Xv, (N, N), [0], aggr), g[f"pool0_{aggr}"])
    close(O.spmm(ei, Av[:, :1], (N, N), 1, x), g["spmm1_scalar"])
    close(O.spmm(ei, None, (N, N), 1, x), g["spmm1_noval"])
    close(O.sp_unpool_dense(tid, 0, x), g["unpool0"])
    close(O.sp_unpool_dense(tid, 1, x), g["unpool1"])
    close(O.scatter_reduce(x, g["batch"], int(g["batch"].max()) + 1, "sum"), g["readout_sum"])


def test_3d_tuples_golden(golden):
    g = golden("tuples3d")
    N = int(g["N"])
    ei, tid, Av, Xv = g["edge_index"], g["tupleid"], g["Av"], g["Xv"]
    acd = O.filterind(tid, *O.spspmm_ind(tid, 2, ei, 0))
    assert np.array_equal(acd, g["acd"])
    for aggr in ("sum", "max"):
        close(O.spspmm(Xv, Av, acd, tid.shape[1], aggr), g[f"mp_{aggr}"])
    for aggr in ("sum", "mean", "max"):
        pi, pv = O.sp_pool_sparse(tid, Xv, [2], aggr)
        assert np.array_equal(pi, g[f"pool2s_ind_{aggr}"])
        close(pv, g[f"pool2s_val_{aggr}"])
        close(O.sp_pool_dense(tid, Xv, (N, N, N), [1, 2], aggr), g[f"pool12_{aggr}"])
        close(O.sp_pool_dense(tid, Xv, (N, N, N), [2], aggr), g[f"pool2_{aggr}"])
    pi, pv = O.sp_pool_sparse(tid, Xv, [2], "sum")
    close(O.sp_unpool_sparse(pi, pv, tid, [2]), g["unpool_sp"])


def test_masked_golden(golden):
    g = golden("masked")
    A, B, mask = g["A"], g["B"], g["mask"]
    for d1 in (1, 2):
        for d2 in (1, 2):
            close(O.mamamm(A, mask, d1, B, mask, d2, mask), g[f"mm_{d1}{d2}"])
    for aggr in ("sum", "mean", "max"):
        for dims in ((1,), (2,), (1, 2)):
            tag = "".join(map(str, dims))
            data, om = O.ma_pool(A, mask, dims, aggr)
            assert np.array_equal(om, g[f"pool{tag}_mask"])
            sel = om.reshape(om.shape + (1,))
            close(np.where(sel, data, 0), np.where(sel, g[f"pool{tag}_{aggr}"], 0))
    data, om = O.ma_pool(A, mask, (2,), "min")
    close(data, g["pool2_min"])
    close(O.ma_fill(A, mask, 1024), g["fill1024"])
    fi = g["filterinf_in"]
    assert np.array_equal(np.where(np.isinf(fi), 0, fi), g["filterinf_out"])


def test_mamamm_equals_einsum_full_mask():
    # with an all-True mask mamamm(A,2,B,1) == einsum("bijd,bjkd->bikd")
    rng = np.random.default_rng(2)
    A = rng.standard_normal((2, 5, 5, 3)).astype(np.float32)
    B = rng.standard_normal((2, 5, 5, 3)).astype(np.float32)
    m = np.ones((2, 5, 5), dtype=bool)
    close(O.mamamm(A, m, 2, B, m, 1, m), np.einsum("bijd,bjkd->bikd", A, B))


def test_scatter_empty_rows_are_zero():
    # SURVEY Q13
    src = np.array([[-3.0], [-5.0]], dtype=np.float32)
    out = O.scatter_reduce(src, np.array([2, 2]), 4, "max")
    assert out.tolist() == [[0.0], [0.0], [-3.0], [0.0]]
    assert O.scatter_reduce(src[:0], np.zeros(0, dtype=np.int64), 3, "sum").shape == (3, 1)


def test_torch_oracle_matches_golden(golden):
    """The differentiable torch restatement agrees with the real reference's values."""
    import torch
    from oracle import torch_oracle as TO
    g = golden("spspmm")
    Tn = lambda k: torch.from_numpy(g[k])  # noqa: E731
    N = int(g["N"])
    ops = {"XA": ("Xv", "Av"), "AX": ("Av", "Xv"), "XX": ("Xv", "Xv")}
    for tag, (p, q) in ops.items():
        for aggr in ("sum", "mean", "max", "min"):
            out = TO.spspmm(Tn(p), Tn(q), Tn(f"{tag}_acd"), g["tupleid"].shape[1], aggr)
            close(out.numpy(), g[f"{tag}_{aggr}"])
    for aggr in ("sum", "mean", "max"):
        close(TO.spmm(Tn("edge_index"), Tn("Av"), (N, N), 1, Tn("x"), aggr).numpy(), g[f"spmm1_{aggr}"])
        close(TO.spmm(Tn("edge_index"), Tn("Av"), (N, N), 0, Tn("x"), aggr).numpy(), g[f"spmm0_{aggr}"])
        close(TO.sp_pool(Tn("tupleid"), Tn("Xv"), (N, N), 0, aggr).numpy(), g[f"pool1_{aggr}"])
        close(TO.sp_pool(Tn("tupleid"), Tn("Xv"), (N, N), 1, aggr).numpy(), g[f"pool0_{aggr}"])
    m = golden("masked")
    A, B, mask = (torch.from_numpy(m[k]) for k in ("A", "B", "mask"))
    for d1 in (1, 2):
        for d2 in (1, 2):
            close(TO.mamamm(A, d1, B, d2, mask).numpy(), m[f"mm_{d1}{d2}"])
    for aggr in ("sum", "mean", "max"):
        for dims in ((1,), (2,), (1, 2)):
            tag = "".join(map(str, dims))
            sel = m[f"pool{tag}_mask"][..., None]
            close(np.where(sel, TO.ma_pool(A, mask, dims, aggr).numpy(), 0),
                  np.where(sel, m[f"pool{tag}_{aggr}"], 0))


def test_oracle_convs_match_reference(golden):
    """Forward values, input gradient and every parameter gradient of one layer of each
    in-scope conv against the real reference layers (tests/golden/conv.npz)."""
    import torch
    from oracle import model_oracle as MO
    g = golden("conv")
    ei, tid = torch.from_numpy(g["edge_index"]), torch.from_numpy(g["tupleid"])
    gd = {k: torch.from_numpy(v) for k, v in g.items() if k.endswith("___acd")}
    gd.update(edge_index=ei, tupleid=tid, num_nodes=int(g["N"]))
    mk = {"NGNN": lambda: MO.ONGNN(8, "sum", 2, 0.1), "SSWL": lambda: MO.OSSWL(8, "sum", 2, 0.1),
          "SSWLmax": lambda: MO.OSSWL(8, "max", 2, 0.1),
          "DSSGNN": lambda: MO.ODSSGNN(8, "sum", 2, 0.1, "mean"),
          "PPGN": lambda: MO.OPPGN(8, "sum", 2, 0.1)}
    for name, fn in mk.items():
        conv = fn()
        sd = {k[len(name) + 4:]: torch.from_numpy(v) for k, v in g.items()
              if k.startswith(name + ".sd.")}
        conv.load_state_dict(sd)
        Av = torch.from_numpy(g["Av"])
        Xv = torch.from_numpy(g["Xv"]).clone().requires_grad_(True)
        out = conv(Av, Xv, gd)
        (out ** 2).mean().backward()
        close(out.detach().numpy(), g[f"{name}.out"], 2e-5)
        close(Xv.grad.numpy(), g[f"{name}.gradX"], 2e-5)
        for k, p in conv.named_parameters():
            close(p.grad.numpy(), g[f"{name}.grad.{k}"], 5e-5)


def test_oracle_i2conv_matches_reference(golden):
    """I2Conv on 3-D tuples + the zinc.py 3-D tupleinit and read-out chain against the real
    reference (tests/golden/conv_i2.npz): layer output, read-out, input / parameter gradients."""
    import torch
    from oracle import model_oracle as MO
    g = golden("conv_i2")
    ei, tid, N = torch.from_numpy(g["edge_index"]), torch.from_numpy(g["tupleid"]), int(g["N"])
    gd = {k: torch.from_numpy(v) for k, v in g.items() if k.endswith("___acd")}
    for name, aggr, pool in (("I2", "sum", "mean"), ("I2max", "max", "max")):
        conv = MO.OI2(8, aggr, 2, 0.1)
        conv.load_state_dict({k[len(name) + 4:]: torch.from_numpy(v) for k, v in g.items()
                              if k.startswith(name + ".sd.")})
        lins = [torch.nn.Linear(8, 8) for _ in range(3)]
        for i, lin in enumerate(lins):
            lin.load_state_dict({"weight": torch.from_numpy(g[f"{name}.init{i}.weight"]),
                                 "bias": torch.from_numpy(g[f"{name}.init{i}.bias"])})
        Av = torch.from_numpy(g["Av"])
        xv = torch.from_numpy(g["Xv"]).clone().requires_grad_(True)
        xn = torch.from_numpy(g["xn"]).clone().requires_grad_(True)
        X = lins[0](xn)[tid[0]] * lins[1](xn)[tid[1]] * lins[2](xn)[tid[1]] * xv
        Y = conv(Av, X, gd)
        h = MO.pool3d_to_dense(X + Y, tid, N, pool)
        ((h ** 2).mean() + (Y ** 2).mean()).backward()
        close(Y.detach().numpy(), g[f"{name}.out"], 2e-5)
        close(h.detach().numpy(), g[f"{name}.readout"], 2e-5)
        close(xv.grad.numpy(), g[f"{name}.gradX"], 5e-5)
        close(xn.grad.numpy(), g[f"{name}.gradx"], 5e-5)
        for k, p in conv.named_parameters():
            close(p.grad.numpy(), g[f"{name}.grad.{k}"], 5e-5)
        for i, lin in enumerate(lins):
            close(lin.weight.grad.numpy(), g[f"{name}.init{i}.gweight"], 5e-5)
            close(lin.bias.grad.numpy(), g[f"{name}.init{i}.gbias"], 5e-5)


def test_oracle_ppgn_dense_matches_reference(golden):
    """Dense (DD) PPGNConv against the real reference on equal-size graphs
    (tests/golden/conv_dd.npz): output, input gradient, every parameter gradient."""
    import torch
    from oracle import model_oracle as MO
    g = golden("conv_dd")
    conv = MO.OPPGNDense(8, 2, 0.1)
    conv.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd.")})
    X = torch.from_numpy(g["X"]).clone().requires_grad_(True)
    out = conv(X, torch.from_numpy(g["mask"]))
    (out ** 2).mean().backward()
    close(out.detach().numpy(), g["out"], 2e-5)
    close(X.grad.numpy(), g["gradX"], 5e-5)
    for k, p in conv.named_parameters():
        close(p.grad.numpy(), g[f"grad.{k}"], 5e-5)


def test_spmamm_golden(golden):
    """oracle spmamm against the reference's spmamm on the inputs the reference can run
    (scalar features; see tests/golden/make_golden.py::golden_spmamm)."""
    g = golden("spmamm")
    for tag, dim2 in (("d1", 1), ("d2", 2)):
        for aggr in ("sum", "max"):
            for dim1 in (1, 2):
                got, mask = O.spmamm(g["ind"], g["aval"], tuple(g["shape"]), dim1, g[f"{tag}_data"],
                                     g[f"{tag}_mask"], dim2, None, aggr)
                close(got, g[f"{tag}_{aggr}_dim{dim1}"])
                assert np.array_equal(mask, g[f"{tag}_mask"])
