"""where / maximum / minimum / abs (interface of the reference's src/tensor/ops/elementwise.py).

On CUDA these run on the device: comparisons yield float32 0/1 masks that STAY on the operands' device
(the reference evaluates them into CPU tensors even for CUDA operands, so its own CUDA path cannot run
this family - SURVEY.md Q8; fixing that is section 8(f) item 3)."""

from __future__ import annotations

from ..types import TensorLike, TProps
from ._tape import be, record


class Elementwise:
    @staticmethod
    def where(condition: TensorLike, a: TensorLike, b: TensorLike) -> TProps:
        backend = be(a)
        mask = condition.data
        out = backend.where(mask, a.data, b.data)
        return record(out, a, [(a, lambda g: backend.where(mask, g, 0.0)), (b, lambda g: backend.where(mask, 0.0, g))])

    @staticmethod
    def maximum(a: TensorLike, b: TensorLike) -> TProps:
        return Elementwise.where(a > b, a, b)

    @staticmethod
    def minimum(a: TensorLike, b: TensorLike) -> TProps:
        return Elementwise.where(a < b, a, b)

    @staticmethod
    def abs(tensor: TensorLike) -> TProps:
        return Elementwise.where(tensor >= 0, tensor, -tensor)
