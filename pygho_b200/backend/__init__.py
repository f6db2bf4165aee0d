"""B200 implementation of the ``pygho.backend`` tensor-operator layer (same module and
function names as the reference package)."""
from .SpTensor import SparseTensor, indicehash, decodehash, indicehash_tight, decodehash_tight, coalesce
from .MaTensor import MaskedTensor, filterinf
from .Spspmm import (spspmm, spspmpnn, spspmm_ind, filterind, spsphadamard, spsphadamard_ind,
                     ptr2batch, deg2batch)
from .Spmm import spmm
from .Mamamm import mamamm
from .Spmamm import spmamm
from .utils import torch_scatter_reduce
