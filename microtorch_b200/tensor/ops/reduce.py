"""sum / mean / max / min (interface of the reference's src/tensor/ops/reduce.py)."""

from __future__ import annotations

from typing import Optional

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
        """The gradient is shared equally between the tied extreme elements (reference reduce.py:46-63).
        CUDA: compare, count, divide and mask are four launches of the existing kernels."""
        backend, x = be(tensor), tensor.data

        def share(grad):
            hit = backend.equal(x, output)
            n = backend.sum(hit, None if axis is None else axis, axis is not None)
            g = grad if (keepdims or axis is None) else backend.expand_dims(grad, axis)
            return backend.mul(hit, backend.div(g, n))

        return share

    @staticmethod
    def _extreme(is_max: bool, tensor: TensorLike, axis, keepdims) -> TProps:
        out = be(tensor).extreme(tensor.data, axis, keepdims, is_max)
        return record(out, tensor, [(tensor, Reduce.bkwd_minmax(tensor, out, axis, keepdims))])

    @staticmethod
    def max(tensor: TensorLike, axis: Optional[Axis], keepdims: bool = False) -> TProps:
        return Reduce._extreme(True, tensor, axis, keepdims)

    @staticmethod
    def min(tensor: TensorLike, axis: Optional[Axis], keepdims: bool = False) -> TProps:
        return Reduce._extreme(False, tensor, axis, keepdims)
