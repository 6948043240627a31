"""log / exp / pow / sqrt / tanh (interface of the reference's src/tensor/ops/math.py).
Every backward is ONE fused backend call."""

from __future__ import annotations

from ..types import TensorLike, TProps
from ._tape import be, record


class MathOps:
    @staticmethod
    def log(tensor: TensorLike) -> TProps:
        backend, x = be(tensor), tensor._data
        return record(backend.log(x), tensor, [(tensor, lambda g: backend.log_bwd(g, x))])

    @staticmethod
    def exp(tensor: TensorLike) -> TProps:
        backend = be(tensor)
        y = backend.exp(tensor._data)
        return record(y, tensor, [(tensor, lambda g: backend.exp_bwd(g, y))])

    @staticmethod
    def pow(tensor: TensorLike, pow: float) -> TProps:
        backend, x = be(tensor), tensor._data
        return record(backend.pow(x, pow), tensor, [(tensor, lambda g: backend.pow_bwd(g, x, pow))])

    @staticmethod
    def sqrt(tensor: TensorLike) -> TProps:
        return MathOps.pow(tensor, 0.5)

    @staticmethod
    def tanh(tensor: TensorLike) -> TProps:
        backend = be(tensor)
        y = backend.tanh(tensor._data)          # backward needs the OUTPUT: g * (1 - y^2)
        return record(y, tensor, [(tensor, lambda g: backend.tanh_bwd(g, y))])
