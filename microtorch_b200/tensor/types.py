"""Autograd tape records (`Leaf`, `TProps`) and the `TensorLike` protocol.

Same record shapes as the reference's src/tensor/types.py (Leaf: types.py:28-38, TProps:
types.py:44-69) - they are the tape format every op returns - with the protocol cut down to
what `Tensor` really implements.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Protocol, Tuple, Union, runtime_checkable

from .device import Device, DType, Scalar, Vector

Data = Union[Scalar, Vector, list, "TensorLike"]
Index = Union[int, slice, List[int], Tuple[Union[int, slice], ...], "TensorLike", Vector]
Shape = Tuple[int, ...]
Dims = Union[int, Shape]
Axis = Union[int, Tuple[int, ...]]


@runtime_checkable
class TensorLike(Protocol):
    """What the ops layer needs from a tensor."""

    @property
    def data(self) -> Vector: ...
    @property
    def requires_grad(self) -> bool: ...
    @property
    def device(self) -> Device: ...
    @property
    def dtype(self) -> DType: ...
    @property
    def shape(self) -> Tuple[int, ...]: ...
    @property
    def ndim(self) -> int: ...
    @property
    def size(self) -> int: ...
    def props(self) -> Tuple: ...
    def backward(self, grad=None) -> None: ...


@dataclass(frozen=True)
class Leaf:
    """One edge of the tape: the upstream tensor and the map from the node's gradient to
    that tensor's gradient."""
    value: "TensorLike"
    grad_fn: Callable[[Vector], Vector]


DependenciesList = List[Leaf]


@dataclass
class TProps:
    """What an op returns: everything `Tensor(*props())` needs."""
    _data: Vector
    requires_grad: bool
    dependencies: DependenciesList
    device: Device
    dtype: DType

    def props(self) -> Tuple:
        return (self._data, self.requires_grad, self.dependencies, self.device, self.dtype)
