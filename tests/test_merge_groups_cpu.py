"""CPU test of plans.merge_groups (fixed-shape torch ops, no kernels): the row-wise concatenation
of two CSR groupings used by the one-launch SSWL gradient (ops.SswlAggregate.backward)."""
import numpy as np
import torch

from pygho_b200 import plans as P


def _group(rng, n_rows, T, cap, n_first, n_second):
    """Random CSR grouping with `cap - T` filler entries behind rowptr[n_rows]."""
    rows = np.sort(rng.integers(0, n_rows, T))
    rowptr = np.zeros(n_rows + 1, dtype=np.int32)
    np.add.at(rowptr, rows + 1, 1)
    rowptr = np.cumsum(rowptr).astype(np.int32)
    first = np.concatenate([rng.integers(0, n_first, T), np.full(cap - T, n_first - 1)]).astype(np.int32)
    second = np.concatenate([rng.integers(0, n_second, T), np.full(cap - T, n_second - 1)]).astype(np.int32)
    return P.Group(torch.from_numpy(rowptr), torch.from_numpy(first), torch.from_numpy(second)), rows


def test_merge_groups_rowwise_concatenation_with_fillers():
    rng = np.random.default_rng(3)
    for n_rows, T1, cap1, T2, cap2 in ((7, 20, 20, 13, 13), (50, 300, 340, 0, 5), (9, 0, 0, 4, 4), (33, 100, 128, 90, 96)):
        g1, rows1 = _group(rng, n_rows, T1, cap1, 40, 11)
        g2, rows2 = _group(rng, n_rows, T2, cap2, 40, 11)
        m = P.merge_groups(g1, g2, n_rows, 3, 1, 2)
        rp = m.rowptr.numpy()
        assert rp[0] == 0 and rp[-1] == T1 + T2 and m.first.numel() == cap1 + cap2
        for r in range(n_rows):
            lo1, hi1 = int(g1.rowptr[r]), int(g1.rowptr[r + 1])
            lo2, hi2 = int(g2.rowptr[r]), int(g2.rowptr[r + 1])
            want_f = np.concatenate([3 * g1.first[lo1:hi1].numpy() + 1, 3 * g2.first[lo2:hi2].numpy() + 2])
            want_s = np.concatenate([g1.second[lo1:hi1].numpy(), g2.second[lo2:hi2].numpy()])
            assert np.array_equal(m.first[rp[r]:rp[r + 1]].numpy(), want_f)
            assert np.array_equal(m.second[rp[r]:rp[r + 1]].numpy(), want_s)
