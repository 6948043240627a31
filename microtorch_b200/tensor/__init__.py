from .ops import BaseOps, Elementwise, MathOps, OverloadOps, Reduce
from .tensor import Tensor
from .types import Leaf, TensorLike, TProps

__all__ = ["Tensor", "TensorLike", "TProps", "Leaf", "BaseOps", "Reduce", "Elementwise", "MathOps", "OverloadOps"]
