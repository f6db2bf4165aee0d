"""MaskedTensor x MaskedTensor contraction (reference ``pygho/backend/Mamamm.py:7-64``).

The 2-FWL product ``X[b,i,k,:] = sum_j A[b,i,j,:] * B[b,j,k,:]`` is the only dense
contraction on the hot path.  The reference permutes both operands to (dense, b, n, n),
lets ``torch.matmul`` materialise contiguous copies, runs a cuBLAS batched GEMM and
returns a strided view; here the operands stay in their (b, n, n, dense) layout and one
kernel reads each tile once, contracts per channel and writes the masked result.
"""
from __future__ import annotations

import os

import torch
from torch import BoolTensor

from ..ops import MaMaMM
from .MaTensor import MaskedTensor


def default_algo() -> int:
    """0 = exact-fp32 CUDA-core kernel, 1 = tcgen05 TF32 tensor-core kernel with one CTA per
    (graph, channel slab), 2 = the same tiles and MMAs in a persistent warp-specialised
    pipeline (the reference runs this contraction in TF32 too, example/zinc.py:30), 4 = exact fp32
    FMAs from a TMA-fed shared-memory ring (default: the contraction is HBM bound, and this is
    both the fastest kernel -- 45.5 us vs 59 us at b = 128, n <= 40, d = 128 -- and bit-identical
    to algo 0)."""
    return int(os.environ.get("PYGHO_B200_MAMAMM_ALGO", "4"))


def mamamm(A: MaskedTensor, dim1: int, B: MaskedTensor, dim2: int, mask: BoolTensor,
           broadcast_firstdim: bool = True) -> MaskedTensor:
    """Contract masked dim ``dim1`` of ``A`` with masked dim ``dim2`` of ``B``, batched
    over dim 0 and elementwise over the dense dims; the result has masked shape
    ``(batch, *rest of A, *rest of B)`` and is valid where ``mask`` is True.

    Supported operand forms (everything the honn operators use, MaOperator.py:102-278):
    ``A`` is (b, n, n, *) with ``dim1`` in {1, 2}, or has more masked dims with the
    contracted one last; ``B`` is (b, n, n, *) with ``dim2`` in {1, 2}, or has more
    masked dims with the contracted one first (dim 1)."""
    assert broadcast_firstdim, "only batched contraction (broadcast_firstdim=True) is supported"
    assert dim1 > 0, "0 dim of A is batch, need to be broadcasted"
    assert dim2 > 0, "0 dim of B is batch, need to be broadcasted"
    ma, mb = A.masked_dim, B.masked_dim
    assert ma >= 3 and mb >= 3, "operands need (batch, n, n) masked dims"
    tA, tB = A.fill_masked(0.0), B.fill_masked(0.0)
    if tA.dtype != torch.float32 or tB.dtype != torch.float32:
        raise TypeError("mamamm needs float32 data")
    if tuple(A.denseshape) != tuple(B.denseshape):
        dshape = torch.broadcast_shapes(A.denseshape, B.denseshape)
        tA = tA.expand(tuple(A.maskedshape) + tuple(dshape))
        tB = tB.expand(tuple(B.maskedshape) + tuple(dshape))
    else:
        dshape = tuple(A.denseshape)
    b = tA.shape[0]
    dense = 1
    for s in dshape:
        dense *= int(s)
    # bring A to (b, R, n_j) or its transpose
    m_a = m_b = None
    if ma == 3:
        trans_a = dim1 == 1
        a_rest = (tA.shape[2],) if trans_a else (tA.shape[1],)
        a3 = tA.reshape(b, tA.shape[1], tA.shape[2], dense)
        m_a = A.mask
    else:
        assert dim1 == ma - 1, "for >2 tuple dims the contracted dim of A must be the last one"
        trans_a = False
        a_rest = tuple(tA.shape[1:ma - 1])
        a3 = tA.reshape(b, -1, tA.shape[ma - 1], dense)
    if mb == 3:
        trans_b = dim2 == 2
        b_rest = (tB.shape[1],) if trans_b else (tB.shape[2],)
        b3 = tB.reshape(b, tB.shape[1], tB.shape[2], dense)
        m_b = B.mask
    else:
        assert dim2 == 1, "for >2 tuple dims the contracted dim of B must be dim 1"
        trans_b = False
        b_rest = tuple(tB.shape[2:mb])
        b3 = tB.reshape(b, tB.shape[1], -1, dense)
    n_i = a3.shape[2] if trans_a else a3.shape[1]
    n_k = b3.shape[1] if trans_b else b3.shape[2]
    algo = default_algo()     # ops._mm falls back to algo 0 per call when a shape does not fit
    omask = mask if mask.ndim == 3 else mask.reshape(b, n_i, n_k)
    out = MaMaMM.apply(a3, trans_a, m_a, b3, trans_b, m_b, omask, algo)
    out = out.reshape((b,) + a_rest + b_rest + tuple(dshape))
    return MaskedTensor(out, mask, 0.0, is_filled=True)
