"""Shape ops and the un-broadcast rule (interface of the reference's src/tensor/ops/base.py).

All of these are metadata-only on both backends - strided views, zero bytes moved - and
therefore bit-exact; their backward is the inverse view.  `bkwd_broadcast` is the one real
kernel here: the sum-reduction that folds a gradient back onto a broadcast operand.
"""

from __future__ import annotations

import numpy as np

from ..types import Axis, Shape, TensorLike, TProps
from ._tape import be, record


class BaseOps:
    @staticmethod
    def bkwd_broadcast(tensor: TensorLike):
        """Closure folding a result-shaped gradient onto `tensor`'s shape (base.py:8-58):
        0-d operand -> total sum; 0-d gradient -> passes through; otherwise sum away the
        prepended axes, then (keepdims) the axes where the operand has extent 1."""
        shape, backend = tuple(tensor.shape), be(tensor)

        def fold(grad):
            return backend.unbroadcast(grad, shape)

        return fold

    @staticmethod
    def broadcast_to(tensor: TensorLike, shape: Shape) -> TProps:
        out = be(tensor).broadcast_to(tensor.data, shape)
        return record(out, tensor, [(tensor, BaseOps.bkwd_broadcast(tensor))])

    @staticmethod
    def view(tensor: TensorLike, shape: Shape) -> TProps:
        backend, src_shape = be(tensor), tuple(tensor.shape)
        out = backend.reshape(tensor.data, shape)
        return record(out, tensor, [(tensor, lambda g: backend.reshape(g, src_shape))])

    @staticmethod
    def reshape(tensor: TensorLike, shape: Shape) -> TProps:
        return BaseOps.view(tensor, shape)

    @staticmethod
    def transpose(tensor: TensorLike, axes: Axis = None) -> TProps:
        backend = be(tensor)
        out = backend.transpose(tensor._data, axes)
        inverse = None if axes is None else tuple(int(i) for i in np.argsort(axes))
        return record(out, tensor, [(tensor, lambda g: backend.transpose(g, inverse))])

    @staticmethod
    def squeeze(tensor: TensorLike, axis: Axis) -> TProps:
        backend, src_shape = be(tensor), tuple(tensor.shape)
        out = backend.squeeze(tensor._data, axis)
        if axis is None:
            undo = lambda g: backend.reshape(g, src_shape)
        else:
            undo = lambda g: backend.expand_dims(g, axis)
        return record(out, tensor, [(tensor, undo)])

    @staticmethod
    def unsqueeze(tensor: TensorLike, dim: int) -> TProps:
        backend = be(tensor)
        out = backend.expand_dims(tensor._data, dim)
        return record(out, tensor, [(tensor, lambda g: backend.squeeze(g, dim))])
