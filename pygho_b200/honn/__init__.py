"""Operators and layers of high-order GNNs on top of ``pygho_b200.backend``."""
from . import Conv, MaOperator, SpOperator, TensorOp, utils  # noqa: F401
