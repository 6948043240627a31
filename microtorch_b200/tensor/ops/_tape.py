"""Shared helper: turn "output array + (source tensor, gradient map) pairs" into a TProps."""

from __future__ import annotations

from typing import Callable, Sequence, Tuple

from ..backend import backend_for
from ..types import Leaf, TensorLike, TProps


def record(out, first: TensorLike, edges: Sequence[Tuple[TensorLike, Callable]]) -> TProps:
    """The tape node of an op: an edge is kept only for sources that require a gradient;
    device and dtype follow the first operand (as in the reference's ops)."""
    deps = [Leaf(value=src, grad_fn=fn) for src, fn in edges if src.requires_grad]
    return TProps(out, any(src.requires_grad for src, _ in edges), deps, device=first.device, dtype=first.dtype)


def be(t: TensorLike):
    return backend_for(t.device)
