"""sum / mean / max / min (interface of the reference's src/tensor/ops/reduce.py)."""

from __future__ import annotations

from typing import Optional

import numpy as np

from ..device import Device
from ..types import Axis, TensorLike, TProps
from ._tape import be, record


class Reduce:
    @staticmethod
    def sum(tensor: TensorLike, axis: int = None, keepdims: bool = False) -> TProps:
        backend, x = be(tensor), tensor.data
        out = backend.sum(x, axis, keepdims)
        return record(out, tensor, [(tensor, lambda g: backend.sum_bwd(g, x, axis, keepdims))])

    @staticmethod
    def mean(tensor: TensorLike, axis: int = None, keepdims: bool = False):
        """sum / count, composed from Tensor ops exactly like the reference (reduce.py:40-43):
        the division is `* count**-1` in float32."""
        count = tensor.data.shape[axis] if axis is not None else tensor.size
        return tensor.sum(axis=axis, keepdims=keepdims) / count

    @staticmethod
    def bkwd_minmax(tensor: TensorLike, output, axis: Optional[Axis], keepdims: bool = False):
        if tensor.device != Device.CPU:
            raise NotImplementedError("max/min are not part of the CUDA training path (SURVEY.md section 8f)")

        def share(grad):
            hit = (tensor.data == output)
            n = np.sum(hit) if axis is None else np.sum(hit, axis=axis, keepdims=True)
            g = grad if (keepdims or axis is None) else np.expand_dims(grad, axis=axis)
            return hit * (g / n)

        return share

    @staticmethod
    def _extreme(fn, tensor: TensorLike, axis, keepdims) -> TProps:
        if tensor.device != Device.CPU:
            raise NotImplementedError("max/min are not part of the CUDA training path (SURVEY.md section 8f)")
        out = fn(tensor.data, axis=axis, keepdims=keepdims)
        return record(out, tensor, [(tensor, Reduce.bkwd_minmax(tensor, out, axis, keepdims))])

    @staticmethod
    def max(tensor: TensorLike, axis: Optional[Axis], keepdims: bool = False) -> TProps:
        return Reduce._extreme(np.max, tensor, axis, keepdims)

    @staticmethod
    def min(tensor: TensorLike, axis: Optional[Axis], keepdims: bool = False) -> TProps:
        return Reduce._extreme(np.min, tensor, axis, keepdims)
