"""where / maximum / minimum / abs (interface of the reference's src/tensor/ops/elementwise.py).
CPU only: they are built on comparisons, which the reference evaluates into CPU tensors even
for CUDA operands, so its CUDA path cannot run them either."""

from __future__ import annotations

import numpy as np

from ..device import Device
from ..types import TensorLike, TProps
from ._tape import record


def _cpu_only(*tensors) -> None:
    for t in tensors:
        if getattr(t, "device", Device.CPU) != Device.CPU:
            raise NotImplementedError("where/maximum/minimum/abs are CPU-only (outside the CUDA training path)")


class Elementwise:
    @staticmethod
    def where(condition: TensorLike, a: TensorLike, b: TensorLike) -> TProps:
        _cpu_only(condition, a, b)
        mask = condition.data
        out = np.where(mask, a.data, b.data)
        return record(out, a, [(a, lambda g: np.where(mask, g, 0.0)), (b, lambda g: np.where(mask, 0.0, g))])

    @staticmethod
    def maximum(a: TensorLike, b: TensorLike) -> TProps:
        return Elementwise.where(a > b, a, b)

    @staticmethod
    def minimum(a: TensorLike, b: TensorLike) -> TProps:
        return Elementwise.where(a < b, a, b)

    @staticmethod
    def abs(tensor: TensorLike) -> TProps:
        return Elementwise.where(tensor >= 0, tensor, -tensor)
