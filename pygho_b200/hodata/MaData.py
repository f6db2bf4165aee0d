"""Padded dense layouts of a batch, built on the device (reference ``hodata/MaData.py``:
``to_dense_adj`` :26-72, ``to_dense_x`` :109-149).  Same signatures; the pads are really
filled (the reference leaves clamped-gather garbage in ``to_dense_x`` and relies on the
MaskedTensor constructor)."""
from __future__ import annotations

import struct
from typing import Optional

import torch
from torch import LongTensor, Tensor

from .. import _lib
from .._lib import call, ptr, stream_ptr
from ..backend.MaTensor import MaskedTensor


def _raw(value, dtype: torch.dtype) -> int:
    """bit pattern of ``value`` in ``dtype`` as an unsigned 64-bit integer."""
    if dtype == torch.int64:
        return int(value) & 0xFFFFFFFFFFFFFFFF
    if dtype == torch.int32:
        return int(value) & 0xFFFFFFFF
    if dtype == torch.float32:
        return struct.unpack("<I", struct.pack("<f", float(value)))[0]
    if dtype == torch.float64:
        return struct.unpack("<Q", struct.pack("<d", float(value)))[0]
    raise TypeError(f"dense layouts support int32/int64/float32/float64 features, got {dtype}")


def to_dense_x(nodeX: Tensor, Xptr: LongTensor, max_num_nodes: Optional[int] = None,
               batch_size: Optional[int] = None, filled_value: float = 0) -> MaskedTensor:
    """``ret[g, i] = nodeX[Xptr[g] + i]`` padded to ``(b, n, *dense)`` with its mask."""
    dev = _lib.require_cuda(nodeX, Xptr)
    if batch_size is None:
        batch_size = Xptr.shape[0] - 1
    if max_num_nodes is None:
        max_num_nodes = int(torch.diff(Xptr).max()) if batch_size else 0
    dense = tuple(nodeX.shape[1:])
    width = 1
    for s in dense:
        width *= int(s)
    src = nodeX.contiguous()
    out = torch.empty((batch_size, max_num_nodes) + dense, dtype=nodeX.dtype, device=dev)
    mask = torch.empty((batch_size, max_num_nodes), dtype=torch.bool, device=dev)
    if out.numel():
        call("pgh_pad_rows", ptr(src), ptr(Xptr.to(torch.int64).contiguous()), batch_size,
             max_num_nodes, width, src.element_size(), _raw(filled_value, nodeX.dtype), ptr(out),
             ptr(mask), stream_ptr(dev))
        _lib.count_launch()
    elif mask.numel():
        mask.zero_()
    return MaskedTensor(out, mask, filled_value, True)


def to_dense_adj(edge_index: LongTensor, edge_batch: LongTensor, edge_attr: Optional[Tensor] = None,
                 max_num_nodes: Optional[int] = None, batch_size: Optional[int] = None,
                 filled_value: float = 0, node_ptr: Optional[LongTensor] = None) -> MaskedTensor:
    """Scatter the edges of a batch into ``(b, n, n, *dense)``.  ``edge_index`` holds LOCAL
    node ids like in the reference (``MaHoData.__inc__`` returns 0 for it); pass ``node_ptr``
    when it holds global ids of a block-diagonal batch instead."""
    dev = _lib.require_cuda(edge_index, edge_batch)
    if edge_attr is None:
        edge_attr = torch.ones(edge_batch.shape[0], device=dev)
    if max_num_nodes is None:
        assert node_ptr is None, "max_num_nodes is needed with global indices"
        max_num_nodes = int(edge_index.max()) + 1
    if batch_size is None:
        batch_size = int(edge_batch.max()) + 1
    dense = tuple(edge_attr.shape[1:])
    width = 1
    for s in dense:
        width *= int(s)
    attr = edge_attr.contiguous()
    out = torch.empty((batch_size, max_num_nodes, max_num_nodes) + dense, dtype=attr.dtype,
                      device=dev)
    mask = torch.empty((batch_size, max_num_nodes, max_num_nodes), dtype=torch.bool, device=dev)
    if out.numel():
        call("pgh_dense_adj", ptr(edge_index[0].contiguous()), ptr(edge_index[1].contiguous()),
             ptr(edge_batch.to(torch.int64).contiguous()),
             ptr(node_ptr.to(torch.int64).contiguous()) if node_ptr is not None else None,
             ptr(attr), edge_batch.shape[0], batch_size, max_num_nodes, width, attr.element_size(),
             _raw(filled_value, attr.dtype), ptr(out), ptr(mask), stream_ptr(dev))
        _lib.count_launch(2)
    return MaskedTensor(out, mask, filled_value, True)


def to_sparse_adj(edge_index: LongTensor, edge_batch: LongTensor, edge_attr: Optional[Tensor] = None,
                  max_num_nodes: Optional[int] = None, batch_size: Optional[int] = None):
    """(b, n, n) SparseTensor of a batch's edges with LOCAL node ids (reference
    ``hodata/MaData.py:74-106``): indices = (edge_batch, row, col), not assumed coalesced."""
    from ..backend.SpTensor import SparseTensor
    if max_num_nodes is None:
        max_num_nodes = int(edge_index.max().item()) + 1
    if batch_size is None:
        batch_size = int(torch.max(edge_batch).item()) + 1
    size = [batch_size, max_num_nodes, max_num_nodes]
    if edge_attr is not None:
        size += list(edge_attr.shape[1:])
    ind = torch.cat((edge_batch.unsqueeze(0), edge_index), dim=0)
    return SparseTensor(ind, edge_attr, shape=size, is_coalesced=False)


def to_dense_tuplefeat(tuplefeat: Tensor, tupleshape: LongTensor, tuplefeatptr: LongTensor,
                       max_tupleshape: Optional[LongTensor] = None, batch_size: Optional[int] = None,
                       feat2mask=None) -> MaskedTensor:
    """Pad the row-major tuple features of every subgraph to ``(b, n1, n2, ..., *dense)``
    (reference ``hodata/MaData.py:152-212``): graph g's block of shape ``tupleshape[g]`` starts
    at ``tuplefeat[tuplefeatptr[g]]``.  Built on the device with index arithmetic; positions
    outside a graph's own shape are masked and hold 0 (the reference gathers clamped garbage
    there and relies on the mask)."""
    dev = _lib.require_cuda(tuplefeat, tupleshape, tuplefeatptr)
    if batch_size is None:
        batch_size = tupleshape.shape[0]
    if max_tupleshape is None:
        max_tupleshape = torch.amax(tupleshape, dim=0)
    dims = [int(v) for v in (max_tupleshape.tolist() if isinstance(max_tupleshape, Tensor)
                             else max_tupleshape)]
    ndim = len(dims)
    shape = tupleshape.to(dev)
    src = tuplefeatptr[:-1].to(dev).reshape([batch_size] + [1] * ndim)
    valid = torch.ones([batch_size] + dims, dtype=torch.bool, device=dev)
    stride = torch.ones((batch_size,), dtype=torch.long, device=dev)
    for k in range(ndim - 1, -1, -1):                    # row-major: last dim is contiguous
        view = [1] * (ndim + 1)
        view[k + 1] = dims[k]
        ar = torch.arange(dims[k], device=dev).reshape(view)
        bview = [batch_size] + [1] * ndim
        src = src + ar * stride.reshape(bview)
        valid = valid & (ar < shape[:, k].reshape(bview))
        stride = stride * shape[:, k]
    src = torch.where(valid, src, torch.zeros_like(src)).clamp_(0, max(tuplefeat.shape[0] - 1, 0))
    data = tuplefeat[src] if tuplefeat.shape[0] else tuplefeat.new_zeros(tuple(src.shape) + tuple(tuplefeat.shape[1:]))
    mask = valid if feat2mask is None else (valid & feat2mask(data))
    return MaskedTensor(data, mask, 0, False)


def batch2dense(batch, batch_size: Optional[int] = None, max_num_nodes: Optional[int] = None,
                denseadj: bool = False, keys=("",)):
    """Convert and pad the per-graph arrays of a collated batch object to dense forms, in place
    (reference ``hodata/MaData.py:215-255``).  ``batch`` is any object with the reference's
    attribute names: ``x``, ``ptr``, ``edge_index`` (local ids), ``edge_index_batch``,
    ``edge_attr`` and, per key, ``tuplefeat<key>``, ``tupleshape<key>``, ``tuplefeat<key>_ptr``."""
    batch.x = to_dense_x(batch.x, batch.ptr, max_num_nodes, batch_size)
    batch_size, max_num_nodes = batch.x.shape[0], batch.x.shape[1]
    if denseadj:
        batch.A = to_dense_adj(batch.edge_index, batch.edge_index_batch, batch.edge_attr,
                               max_num_nodes, batch_size)
    else:
        batch.A = to_sparse_adj(batch.edge_index, batch.edge_index_batch, batch.edge_attr,
                                max_num_nodes, batch_size)
    for key in keys:
        X = to_dense_tuplefeat(getattr(batch, f"tuplefeat{key}"), getattr(batch, f"tupleshape{key}"),
                               getattr(batch, f"tuplefeat{key}_ptr"), None, batch_size, None)
        setattr(batch, f"X{key}", X)
    return batch
