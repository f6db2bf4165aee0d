"""``MaskedTensor`` -- dense tensor plus validity mask (reference
``pygho/backend/MaTensor.py``: ``filterinf`` :8-31, class :34-266).

Deviation from the reference, on purpose: the reference records
``padvalue`` *before* calling ``fill_masked_`` so its constructor never fills the masked
entries, and its ``min`` calls ``amax``.  This class implements the documented,
intended semantics: masked entries always hold ``padvalue`` and ``min`` is a minimum.
On inputs whose pads already equal ``padvalue`` both agree at every valid entry.
"""
from __future__ import annotations

from typing import Callable, Iterable, List, Tuple, Union

import torch
from torch import BoolTensor, LongTensor, Tensor

from .._lib import AGGR_CODE
from ..ops import MaskedFill, MaskedPool


def filterinf(X: Tensor, filled_value: float = 0) -> Tensor:
    """Replace +-inf by ``filled_value`` (reference MaTensor.py:8-31)."""
    return X.masked_fill(torch.isinf(X), filled_value)


def _fill(data: Tensor, mask: Tensor, value: float) -> Tensor:
    if data.dtype == torch.float32:
        return MaskedFill.apply(data.contiguous(), mask, float(value))
    # integer feature tensors before the embedding layers: dtype plumbing, not the hot path
    neg = torch.logical_not(mask).reshape(mask.shape + (1,) * (data.ndim - mask.ndim))
    return data.masked_fill(neg, value)


def _prod(xs) -> int:
    r = 1
    for x in xs:
        r *= int(x)
    return r


class MaskedTensor:
    """``data`` (*maskedshape, *denseshape) with ``mask`` (*maskedshape) True = valid."""

    def __init__(self, data: Tensor, mask: BoolTensor, padvalue: float = 0.0,
                 is_filled: bool = False):
        assert data.ndim >= mask.ndim, "data's #dim should be larger than mask "
        assert data.shape[:mask.ndim] == mask.shape, \
            "data and mask's first dimensions should match"
        self._mask = mask
        self._masked_dim = mask.ndim
        self._padvalue = padvalue
        self._data = data if is_filled else _fill(data, mask, padvalue)
        self._neg = None

    # ---- fill ---------------------------------------------------------------------------
    def fill_masked_(self, val: float = 0.0) -> None:
        if self._padvalue == val:
            return
        self._data = _fill(self._data, self._mask, val)
        self._padvalue = val

    def fill_masked(self, val: float = 0.0) -> Tensor:
        if self._padvalue == val:
            return self._data
        return _fill(self._data, self._mask, val)

    def to(self, device, non_blocking: bool = True):
        self._data = self._data.to(device, non_blocking=non_blocking)
        self._mask = self._mask.to(device, non_blocking=non_blocking)
        self._neg = None
        return self

    # ---- properties ----------------------------------------------------------------------
    @property
    def padvalue(self) -> float:
        return self._padvalue

    @property
    def data(self) -> Tensor:
        return self._data

    @property
    def mask(self) -> BoolTensor:
        return self._mask

    @property
    def fullnegmask(self) -> BoolTensor:
        if self._neg is None:
            self._neg = torch.logical_not(self._mask).reshape(
                self._mask.shape + (1,) * self.dense_dim)
        return self._neg

    @property
    def shape(self) -> torch.Size:
        return self._data.shape

    @property
    def masked_dim(self) -> int:
        return self._masked_dim

    @property
    def dense_dim(self) -> int:
        return self._data.ndim - self._masked_dim

    @property
    def maskedshape(self):
        return self.shape[:self._masked_dim]

    @property
    def denseshape(self):
        return self.shape[self._masked_dim:]

    # ---- pooling over masked dims (reference MaTensor.py:175-206) ---------------------------
    def _pool(self, dims: Union[Iterable[int], int], keepdim: bool, aggr: str) -> "MaskedTensor":
        if isinstance(dims, int):
            dims = [dims]
        md = self._masked_dim
        dims = sorted({d % md if d < 0 else d for d in dims})
        assert all(0 <= d < md for d in dims), "can only pool masked dims"
        data, mask = self.fill_masked(0.0), self._mask
        if data.dtype != torch.float32:
            raise TypeError("masked pooling needs float32 data")
        if dims != list(range(dims[0], dims[-1] + 1)):      # not adjacent: bring together
            order = [i for i in range(md) if i not in dims]
            at = sum(1 for i in order if i < dims[0])
            order = order[:at] + dims + order[at:]
            data = data.permute(order + list(range(md, data.ndim)))
            mask = mask.permute(order)
            keptshape = [self.shape[i] for i in order]
            lo = at
        else:
            keptshape = list(self.shape[:md])
            lo = dims[0]
        hi = lo + len(dims)
        outer, red, inner = _prod(keptshape[:lo]), _prod(keptshape[lo:hi]), _prod(keptshape[hi:])
        dense = _prod(self.denseshape)
        d4 = data.reshape(outer, red, inner, dense)
        pad = (-dense) % 4       # the kernels move 128-bit channel groups; rare odd widths are padded
        if pad:
            d4 = torch.nn.functional.pad(d4, (0, pad))
        out, omask = MaskedPool.apply(d4.contiguous(), mask.reshape(outer, red, inner).contiguous(),
                                      1, AGGR_CODE[aggr])
        if pad:
            out = out[..., :dense]
        oshape = keptshape[:lo] + ([1] * len(dims) if keepdim else []) + keptshape[hi:]
        if keepdim and dims != list(range(dims[0], dims[-1] + 1)):
            raise NotImplementedError("keepdim with non-adjacent dims")
        out = out.reshape(tuple(oshape) + tuple(self.denseshape))
        return MaskedTensor(out, omask.reshape(oshape), padvalue=0, is_filled=True)

    def sum(self, dims: Union[Iterable[int], int], keepdim: bool = False) -> "MaskedTensor":
        return self._pool(dims, keepdim, "sum")

    def mean(self, dims: Union[Iterable[int], int], keepdim: bool = False) -> "MaskedTensor":
        return self._pool(dims, keepdim, "mean")

    def max(self, dims: Union[Iterable[int], int], keepdim: bool = False) -> "MaskedTensor":
        return self._pool(dims, keepdim, "max")

    def min(self, dims: Union[Iterable[int], int], keepdim: bool = False) -> "MaskedTensor":
        return self._pool(dims, keepdim, "min")

    # ---- views (reference MaTensor.py:208-234) ---------------------------------------------
    def diag(self, dims: Iterable[int]) -> "MaskedTensor":
        """Diagonal over the listed masked dims, placed at the first of them."""
        assert len(dims) >= 2, "must diag several dims"
        dims = sorted(list(dims))
        if len(dims) != 2:
            raise NotImplementedError("diag over more than two masked dims")
        tdata = torch.diagonal(self._data, 0, dims[0], dims[1])
        tmask = torch.diagonal(self._mask, 0, dims[0], dims[1])
        return MaskedTensor(torch.movedim(tdata, -1, dims[0]), torch.movedim(tmask, -1, dims[0]),
                            self._padvalue, True)

    def unpooling(self, dims: Union[int, Iterable[int]], tarX: "MaskedTensor") -> "MaskedTensor":
        """Insert the masked dims ``dims`` (sizes taken from ``tarX``) by broadcasting."""
        if isinstance(dims, int):
            dims = [dims]
        dims = sorted(list(dims))
        tdata = self._data
        for d in dims:
            tdata = tdata.unsqueeze(d)
        tdata = tdata.expand(*(tarX.shape[i] if i in dims else -1 for i in range(tdata.ndim)))
        return MaskedTensor(tdata, tarX.mask, self._padvalue, False)

    # ---- tuplewise ops (reference MaTensor.py:236-266) ---------------------------------------
    def tuplewiseapply(self, func: Callable[[Tensor], Tensor]) -> "MaskedTensor":
        return MaskedTensor(func(self.fill_masked(0.0)), self._mask)

    def diagonalapply(self, func: Callable[[Tensor, LongTensor], Tensor]) -> "MaskedTensor":
        assert self._masked_dim == 3, "only implemented for 2D"
        eye = torch.eye(self.shape[1], self.shape[2], dtype=torch.long, device=self._data.device)
        return MaskedTensor(func(self._data, eye.unsqueeze(0).expand_as(self._mask)), self._mask)

    def add(self, tarX: "MaskedTensor", samesparse: bool) -> "MaskedTensor":
        assert isinstance(tarX, MaskedTensor)
        if samesparse:
            same_pad = self._padvalue == tarX.padvalue == 0
            return MaskedTensor(tarX.data + self._data, self._mask, self._padvalue,
                                is_filled=same_pad)
        return MaskedTensor(tarX.fill_masked(0.0) + self.fill_masked(0.0),
                            torch.logical_or(self._mask, tarX.mask), 0, True)

    def catvalue(self, tarX: Union["MaskedTensor", Iterable["MaskedTensor"]],
                 samesparse: bool) -> "MaskedTensor":
        assert samesparse == True, "must have the same sparcity to concat value"  # noqa: E712
        if isinstance(tarX, MaskedTensor):
            tarX = [tarX]
        same_pad = all(t.padvalue == self._padvalue for t in tarX)
        cat = torch.concat([self._data] + [t.data for t in tarX], dim=-1)
        return MaskedTensor(cat, self._mask, self._padvalue, is_filled=same_pad)
