"""Arithmetic, indexing and matmul (interface of the reference's src/tensor/ops/overload.py)."""

from __future__ import annotations

import numpy as np

from ..types import Index, TensorLike, TProps
from ._tape import be, record


def normalize_index(index: Index):
    """Tensor/array-like -> raw array; list -> integer array; 1-tuple -> its element."""
    if hasattr(index, "data") and not isinstance(index, (np.ndarray, memoryview)):
        index = index.data
    if isinstance(index, list):
        return np.array(index)
    if isinstance(index, tuple) and len(index) == 1:
        return index[0]
    return index


class OverloadOps:
    @staticmethod
    def get_item(tensor: TensorLike, index: Index) -> TProps:
        """x[index]; backward scatters by ASSIGNMENT into zeros (duplicates do not add up)."""
        backend = be(tensor)
        index = normalize_index(index)
        src = tensor.data
        out = backend.getitem(src, index)
        return record(out, tensor, [(tensor, lambda g: backend.getitem_bwd(g, src, index))])

    @staticmethod
    def neg(tensor: TensorLike) -> TProps:
        backend = be(tensor)
        return record(backend.neg(tensor.data), tensor, [(tensor, lambda g: backend.neg(g))])

    @staticmethod
    def add(a: TensorLike, b: TensorLike) -> TProps:
        backend = be(a)
        sa, sb = tuple(a.shape), tuple(b.shape)
        return record(backend.add(a.data, b.data), a,
                      [(a, lambda g: backend.unbroadcast(g, sa)), (b, lambda g: backend.unbroadcast(g, sb))])

    @staticmethod
    def mul(a: TensorLike, b: TensorLike) -> TProps:
        """d/da = unbroadcast(g * b), d/db = unbroadcast(g * a): product and reduction fused."""
        backend = be(a)
        sa, sb = tuple(a.shape), tuple(b.shape)
        return record(backend.mul(a.data, b.data), a,
                      [(a, lambda g: backend.unbroadcast(g, sa, b.data)),
                       (b, lambda g: backend.unbroadcast(g, sb, a.data))])

    @staticmethod
    def matmul(a: TensorLike, b: TensorLike) -> TProps:
        """a @ b; dA = g @ b^T, dB = a^T @ g (outer products for 1-D operands).  Gradients of a
        batched product are folded back onto lower-rank operands (the reference omits this and
        raises on 3-D Linear inputs)."""
        backend = be(a)
        sa, sb = tuple(a.shape), tuple(b.shape)

        def grad_a(g):
            r = backend.matmul_bwd_a(g, a.data, b.data)
            return r if tuple(r.shape) == sa else backend.unbroadcast(r, sa)

        def grad_b(g):
            r = backend.matmul_bwd_b(g, a.data, b.data)
            return r if tuple(r.shape) == sb else backend.unbroadcast(r, sb)

        return record(backend.matmul(a.data, b.data), a, [(a, grad_a), (b, grad_b)])
