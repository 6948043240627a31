"""Autograd op families behind `Tensor` (device-agnostic: each op asks `backend_for(device)` for its kernels)."""

from .reduce import Reduce
from .overload import OverloadOps
from .math import MathOps
from .elementwise import Elementwise
from .base import BaseOps

__all__ = ("BaseOps", "Elementwise", "MathOps", "OverloadOps", "Reduce")
