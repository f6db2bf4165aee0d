"""Host-side logic of capacity-padded batches (pygho_b200/static.py): no GPU needed."""
import numpy as np

from oracle import pygho_oracle as O


def _batches():
    from pygho_b200.hodata.synthetic import make_batch
    keys = ["X___X___1___A___0", "X___A___1___X___0"]
    hbs = [make_batch(6, seed=60 + i) for i in range(3)]
    for hb in hbs:
        for key in keys:
            _o0, o1, d1, o2, d2 = key.split("___")
            pick = lambda op: hb.edge_index if op == "A" else hb.tupleid  # noqa: E731
            hb.plans[key] = O.filterind(hb.tupleid, *O.spspmm_ind(pick(o1), int(d1), pick(o2), int(d2)))
    return hbs, keys


def test_pads_are_inert_and_shapes_are_static():
    from pygho_b200 import static as ST
    hbs, keys = _batches()
    caps = ST.capacities(hbs, keys, margin=0.03)
    assert len({caps["N"], caps["A"], caps["X"], caps["B"] + 1}) == 4
    padded = [ST.pad_host_batch(hb, caps, keys) for hb in hbs]
    for hb, pb in zip(hbs, padded):
        nX, N, B = pb.valid.tolist()
        nA = hb.edge_index.shape[1]
        assert (nX, N, B) == (hb.tupleid.shape[1], hb.num_nodes, hb.num_graphs)
        # one shape for every batch
        assert pb.x.shape == (caps["N"],) and pb.edge_index.shape == (2, caps["A"])
        assert pb.tupleid.shape == (2, caps["X"]) and pb.y.shape == (B + 1,) and pb.num_graphs == B + 1
        # valid part untouched, pads at the end and pointing at pad rows only
        assert np.array_equal(pb.tupleid[:, :nX], hb.tupleid) and np.array_equal(pb.x[:N], hb.x)
        assert (pb.tupleid[:, nX:] >= N).all() and (pb.edge_index[:, nA:] >= N).all()
        assert (pb.batch[N:] == B).all() and (pb.batch[:N] < B).all()
        for key in keys:
            _o0, o1, _d1, o2, _d2 = key.split("___")
            acd, T = pb.plans[key], hb.plans[key].shape[1]
            assert acd.shape == (3, caps[key]) and np.array_equal(acd[:, :T], hb.plans[key])
            capn = {"A": caps["A"], "X": caps["X"]}
            # filler triples are "no row" markers: one past the last row of every array
            assert (acd[0, T:] == caps["X"]).all()
            assert (acd[1, T:] == capn["A" if o1 == "A" else "X"]).all()
            assert (acd[2, T:] == capn["A" if o2 == "A" else "X"]).all()
            assert (np.diff(acd[0]) >= 0).all()                    # still grouped by output row
        # spspmm on the padded arrays == spspmm on the exact arrays at every valid row
        rng = np.random.default_rng(1)
        Xv = rng.standard_normal((caps["X"], 4)).astype(np.float32)
        Av = rng.standard_normal((caps["A"], 4)).astype(np.float32)
        T0 = hb.plans[keys[0]].shape[1]
        got = O.spspmm(Xv, Av, pb.plans[keys[0]][:, :T0], caps["X"])   # fillers are never visited
        want = O.spspmm(Xv[:nX], Av[:nA], hb.plans[keys[0]], nX)
        assert np.array_equal(got[:nX], want) and not got[nX:].any()


def test_batch_that_does_not_fit_is_rejected():
    import pytest
    from pygho_b200 import static as ST
    hbs, keys = _batches()
    caps = ST.capacities(hbs[:1], keys)
    big = max(hbs, key=lambda hb: hb.tupleid.shape[1])
    if big is not hbs[0]:
        with pytest.raises(ValueError):
            ST.pad_host_batch(big, caps, keys)
