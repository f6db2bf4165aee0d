"""pygho_b200 -- B200-native (sm_100a) implementation of PygHO's tensor-operator layer.

Drop-in for ``pygho``'s top level (reference ``pygho/__init__.py:1-2``)."""
from .backend.SpTensor import SparseTensor
from .backend.MaTensor import MaskedTensor

__all__ = ["SparseTensor", "MaskedTensor"]
__version__ = "0.1.0"
