"""`Tensor`, its tape record types and the op families - the names `src.tensor` exports in the reference - plus the
device / dtype enums that callers of this package usually want next to them."""

from .device import Device, DType
from .tensor import Tensor
from .types import Leaf, TensorLike, TProps
from .ops import BaseOps, Elementwise, MathOps, OverloadOps, Reduce

_REFERENCE_NAMES = ("Tensor", "TensorLike", "TProps", "Leaf", "BaseOps", "Reduce", "Elementwise", "MathOps", "OverloadOps")
__all__ = _REFERENCE_NAMES + ("Device", "DType")
