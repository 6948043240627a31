"""Tensor: storage handle + autograd tape + operator overloads.

Public surface and error behaviour mirror the reference's src/tensor/tensor.py (same
constructor argument order - it is what `from_props` splats -, same decorators, same
exceptions), so code and tests written against the reference run unchanged.  What differs is
underneath Device.CUDA:

  * `.data` / `.grad` are `microtorch_b200.cuda.DeviceArray`s (pool blocks driven through the
    C ABI), not cupy arrays;
  * gradients are allocated lazily: the reference zero-fills a grad buffer for EVERY tensor that
    requires grad, intermediates included (tensor.py:76-85 - two passes over 1 GiB per hidden
    activation at B=65536); here `.grad` materialises zeros only when somebody reads it, and the
    first gradient that arrives is adopted instead of added to zeros;
  * `backward()` walks the tape in topological order, once per node, instead of the reference's
    recursive depth-first walk that revisits shared sub-graphs once per edge (tensor.py:281-283).
    Gradient maps are linear, so the sums are the same up to fp32 rounding order.
"""

from __future__ import annotations

import pickle
from functools import wraps
from pathlib import Path
from typing import Callable, Dict, List, Optional, Tuple, TypeGuard, Union, final

import numpy as np

from .backend import backend_for
from .device import Device, DType, Scalar, Vector, _device, _tensor, get_dtype
from .ops import BaseOps, Elementwise, MathOps, OverloadOps, Reduce
from .types import Data, DependenciesList, Dims, Index, Shape, TensorLike, TProps


# --------------------------------------------------------------------------- gates
def is_tensorlike(obj: object) -> TypeGuard[TensorLike]:
    return isinstance(obj, Tensor)


def data_cast(t: Union[Data, "Tensor"], device: Device) -> "Tensor":
    """Anything that is not a Tensor becomes one on `device`.  A Python scalar operand of a CUDA op becomes a
    shared read-only device constant (uploaded once per value) instead of a blocking 4-byte copy per use."""
    if isinstance(t, Tensor):
        return t
    if device == Device.CUDA and isinstance(t, (int, float)) and not isinstance(t, bool):
        from ..cuda.array import device_constant
        return Tensor(device_constant(t), device=device)
    return Tensor(t, device=device)


def data_gate(fn):
    """Coerce the second operand to a Tensor on the first operand's device."""
    @wraps(fn)
    def gated(self: "Tensor", other, *args, **kwargs):
        return fn(self, data_cast(other, self.device), *args, **kwargs)
    return gated


def device_gate(fn):
    """Both operands must live on the same device."""
    @wraps(fn)
    def gated(self: "Tensor", other: "Tensor", *args, **kwargs):
        if self.device != other.device:
            raise ValueError(f"Tensors on different devices: {self.device} vs {other.device}")
        return fn(self, other, *args, **kwargs)
    return gated


def input_gate(fn):
    return data_gate(device_gate(fn))


def from_op(fn):
    """Ops return TProps records; wrap them into a Tensor."""
    @wraps(fn)
    def wrapped(*args, **kwargs):
        return Tensor.from_props(fn(*args, **kwargs))
    return wrapped


def op_gate(fn):
    return input_gate(from_op(fn))


_DEVICE_ARRAY = None


def _device_array_type():
    """microtorch_b200.cuda.DeviceArray, resolved on first use (a CUDA tensor implies the namespace is importable)."""
    global _DEVICE_ARRAY
    if _DEVICE_ARRAY is None:
        from ..cuda.array import DeviceArray
        _DEVICE_ARRAY = DeviceArray
    return _DEVICE_ARRAY


class _LazyZero:
    """Marker: "this tensor's gradient is all zeros and has not been allocated yet"."""
    __slots__ = ()

    def __repr__(self) -> str:
        return "<lazy zero grad>"


_LAZY = _LazyZero()


@final
class Tensor:
    # Satisfies the `TensorLike` protocol structurally (isinstance(t, TensorLike) holds) but does not inherit from it:
    # a Protocol base gives the class typing's metaclass, and then every isinstance(x, Tensor / Parameter) on the
    # per-step host path (Module.parameters, the gates) costs a structural check instead of a pointer compare.
    float64 = DType.FLOAT64
    float32 = DType.FLOAT32
    int64 = DType.INT64
    int32 = DType.INT32
    int16 = DType.INT16
    int8 = DType.INT8

    def __init__(self, data: Data, requires_grad: bool = False, dependencies: Optional[DependenciesList] = None,
                 device: Union[Device, str] = Device.CPU, dtype: Optional[DType] = None) -> None:
        self._device = device if type(device) is Device else _device(device)
        self._dtype = dtype or DType.FLOAT32
        self._dependencies: DependenciesList = dependencies or []
        # per-op host path: the result array of an op already is a native float32 array of this device
        if (self._dtype is DType.FLOAT32 and self._device is Device.CUDA and type(data) is _device_array_type()
                and data.dtype == np.float32):
            self._data_ = data
        else:
            self._data_ = self.build_ndarray(data, self._dtype, self._device)
        self._requires_grad = requires_grad
        self._grad = None
        self._grad_hook: Optional[Callable[["Tensor"], None]] = None
        self._fused_ctx = None
        if requires_grad:
            self.zero_grad()

    @classmethod
    def from_props(cls, prps: Union[TProps, TensorLike]) -> "Tensor":
        return cls(*prps.props())

    # ------------------------------------------------------------------ core fields
    @property
    def _data(self) -> Vector:
        return self._data_

    @_data.setter
    def _data(self, value: Vector) -> None:
        self._data_ = value

    @property
    def data(self) -> Vector:
        return self._data_

    @data.setter
    def data(self, new_data: Data) -> None:
        self._data_ = self.build_ndarray(new_data, self.dtype, self.device)
        self.zero_grad()

    @property
    def requires_grad(self) -> bool:
        return self._requires_grad

    @requires_grad.setter
    def requires_grad(self, flag: bool) -> None:
        self._requires_grad = flag

    @property
    def dependencies(self) -> DependenciesList:
        return self._dependencies

    @dependencies.setter
    def dependencies(self, deps: DependenciesList) -> None:
        self._dependencies = deps

    @property
    def device(self) -> Device:
        return self._device

    @device.setter
    def device(self, dev: Device) -> None:
        self._device = dev

    @property
    def dtype(self) -> DType:
        return self._dtype

    @dtype.setter
    def dtype(self, dt: DType) -> None:
        self._dtype = dt

    @property
    def grad(self) -> Optional[Vector]:
        """The gradient buffer.  A lazily-zero CUDA gradient is materialised on first read, and
        a broadcast (stride-0) gradient is made dense, so callers always see a real array."""
        g = self._grad
        if g is _LAZY:
            g = backend_for(self._device).zeros_like(self._data_, dtype=get_dtype(self._device, self._dtype))
            self._grad = g
        elif self._device == Device.CUDA and g is not None and not g.is_dense:
            g = g.copy()
            self._grad = g
        return g

    @grad.setter
    def grad(self, value: Optional[Vector]) -> None:
        self._grad = value

    @property
    def size(self) -> int:
        return self._data_.size

    @property
    def shape(self) -> Tuple[int, ...]:
        return self._data_.shape

    @property
    def ndim(self) -> int:
        return self._data_.ndim

    def props(self) -> Tuple:
        return (self._data_, self._requires_grad, self._dependencies, self._device, self._dtype)

    def item(self) -> Scalar:
        if self._data_.size != 1:
            raise ValueError(f"Cannot convert tensor with shape {self.shape} to a scalar.")
        return self._data_.item()

    @staticmethod
    def build_ndarray(data: Data, dtype: DType = DType.FLOAT32, device: Device = Device.CPU) -> Vector:
        """Tensor -> its array; a native array of the right dtype passes through untouched;
        anything else is converted (on CUDA: cast on the host, then one H2D copy)."""
        if isinstance(data, Tensor):
            return data.data
        lib = _tensor(device)
        want = get_dtype(device, dtype)
        if isinstance(data, lib.ndarray):
            return data if data.dtype == want else data.astype(want)
        if device == Device.CUDA and isinstance(data, np.ndarray) and data.dtype != want:
            data = data.astype(want)
        return lib.array(data, dtype=want)

    def to(self, device: Union[Device, str], dtype: DType = None) -> "Tensor":
        target = _device(device)
        new_dtype = dtype or self.dtype
        if target == self.device and new_dtype == self.dtype:
            return self
        impl = get_dtype(target, new_dtype)
        if target == self.device:
            moved = self._data_.astype(impl)
        elif self.device == Device.CUDA and target == Device.CPU:
            moved = self._data_.get().astype(impl)                     # D2H
        elif self.device == Device.CPU and target == Device.CUDA:
            moved = _tensor(target).array(self._data_.astype(impl) if self._data_.dtype != impl else self._data_,
                                          dtype=impl)                  # H2D
        else:
            raise RuntimeError(f"Unsupported device transfer: {self.device} -> {target}")
        return Tensor(moved, requires_grad=self.requires_grad, dependencies=self.dependencies, device=target,
                      dtype=new_dtype)

    # ------------------------------------------------------------------ gradients
    def zero_grad(self) -> None:
        """Reset the gradient to zeros.  CPU: in place / eager like the reference.  CUDA: back
        to the unallocated lazy zero - no memset pass, and the next backward adopts its first
        gradient instead of accumulating into zeros."""
        if not self._requires_grad:
            return
        if self._device == Device.CUDA:
            self._grad = _LAZY
        elif self._grad is None or self._grad is _LAZY:
            self._grad = np.zeros_like(self._data_, dtype=get_dtype(self._device, self._dtype))
        else:
            self._grad.fill(0.0)

    def release_grad(self) -> None:
        self._grad = None

    def _receive(self, g: Vector) -> None:
        """Add one finished gradient `g` into this tensor's gradient."""
        if self._device != Device.CUDA:
            self._grad = g if self._grad is None else np.add(self._grad, g)
            return
        cur = self._grad
        if cur is None or cur is _LAZY:
            # adopt; an array already adopted elsewhere is copied so two tensors never share a
            # buffer that zero_grad()/SGD may later write in place
            # A broadcast / strided view (sum().backward()'s stride-0 gradient travelling down a chain of
            # adds) is adopted as is, however many tensors hold it: nothing writes through a view, and
            # the `grad` getter makes a private dense array of it on first read.
            if g.is_dense:
                if g._owned or g._const:
                    g = g.copy()
                g._owned = True
            self._grad = g
        else:
            acc = backend_for(self._device).accumulate(cur, g)
            acc._owned = True
            self._grad = acc

    def backward(self, grad: Optional[Vector] = None) -> None:
        if not self.requires_grad:
            raise ValueError("Backward was called on a non-required-grad tensor!")
        if grad is None:
            if self.shape == ():
                if self.device == Device.CUDA:
                    from ..cuda.array import device_constant
                    grad = device_constant(1.0)
                else:
                    grad = self.build_ndarray(1.0, self.dtype, self.device)
            else:
                raise ValueError(f"Grad must be provided for non-scalar tensors. Tensor shape: {self.shape}")
        if isinstance(grad, np.generic) and self.device == Device.CPU:
            grad = np.asarray(grad)          # 0-d results decay to scalars in NumPy
        if not isinstance(grad, type(self.data)):
            raise ValueError((f"Grad device does not match tensor device: {self.device}. "
                              "Grad must be of same backend type. "
                              f"Expected {type(self.data)}, got {type(grad)}"))
        if grad.shape != self.shape:
            raise ValueError(f"Grad shape {grad.shape} does not match tensor shape {self.shape}")
        _run_backward(self, grad)

    def clip_grad(self, clip_value: float = 1.0) -> None:
        if self.requires_grad and self.grad is not None:
            self.grad = _tensor(self.device).clip(self.grad, -clip_value, clip_value)

    def clip_grad_norm(self, max_norm: float = 1.0, norm_type: float = 2.0, eps: float = 1e-6) -> None:
        if self.requires_grad and self.grad is not None:
            norm = _tensor(self.device).linalg.norm(self.grad, norm_type)
            if norm > max_norm:
                self.grad = self.grad * (max_norm / (norm + eps))

    # ------------------------------------------------------------------ persistence / copies
    def save(self, file_path: Union[str, Path]) -> None:
        host = lambda a: a.get() if (a is not None and self.device == Device.CUDA) else a
        state = {"data": host(self.data), "requires_grad": self.requires_grad, "grad": host(self.grad),
                 "device": self.device.value, "dtype": self.dtype.value}
        with open(file_path, "wb") as f:
            pickle.dump(state, f)

    @classmethod
    def load(cls, file_path: Union[str, Path]) -> "Tensor":
        file_path = Path(file_path)
        if not file_path.exists():
            raise FileNotFoundError(f"No file found at {file_path}")
        try:
            with open(file_path, "rb") as f:
                state = pickle.load(f)
        except (pickle.UnpicklingError, EOFError) as e:
            raise ValueError(f"Invalid tensor file format at {file_path}") from e
        if not isinstance(state, dict) or "data" not in state:
            raise ValueError(f"Invalid tensor file format at {file_path}")
        device = Device(state["device"])
        t = cls(data=state["data"], requires_grad=state["requires_grad"], device=device, dtype=DType(state["dtype"]))
        g = state["grad"]
        t.grad = _tensor(device).array(g) if (g is not None and device == Device.CUDA) else g
        return t

    def detach(self) -> "Tensor":
        return Tensor(self.data, device=self.device, requires_grad=False, dtype=self.dtype)

    def clone(self) -> "Tensor":
        return Tensor(self.data.copy(), device=self.device, requires_grad=self.requires_grad, dtype=self.dtype,
                      dependencies=(self.dependencies.copy() if self.dependencies else None))

    def copy(self) -> "Tensor":
        return self.clone()

    def __repr__(self) -> str:
        return (f"Tensor(data={self.data}, requires_grad={self.requires_grad}, shape={self.shape}, "
                f"dtype={self.dtype}, device={self.device})")

    # ------------------------------------------------------------------ factories: drawn on the tensor's own device,
    # as the reference does (`_tensor(device).random...`, tensor.py:402-434): NumPy's generator on the CPU, the
    # library's Philox stream (microtorch_b200.cuda.random) on CUDA.  Parameter init stays on the host (Q7: nn/param.py).
    @staticmethod
    def randn(dims: Dims = (), requires_grad=False, device: Device = Device.CPU) -> "Tensor":
        if type(dims) is int:
            dims = (dims,)
        return Tensor(_tensor(device).random.randn(*dims), requires_grad=requires_grad, device=device)

    @staticmethod
    def uniform(low: float, high: float, dims: Dims = (), requires_grad=False, device: Device = Device.CPU) -> "Tensor":
        if type(dims) is int:
            dims = (dims,)
        return Tensor(_tensor(device).random.uniform(low, high, dims), requires_grad=requires_grad, device=device)

    # ------------------------------------------------------------------ shape ops
    @from_op
    def view(self, shape: Shape) -> "Tensor":
        return BaseOps.view(self, shape)

    def reshape(self, shape: Shape) -> "Tensor":
        return self.view(shape)

    @from_op
    def broadcast_to(self, shape: Shape) -> "Tensor":
        return BaseOps.broadcast_to(self, shape)

    @from_op
    def transpose(self, axis: Shape = None) -> "Tensor":
        return BaseOps.transpose(self, axis)

    @property
    def T(self) -> "Tensor":
        return self.transpose()

    @from_op
    def squeeze(self, dim: int | Tuple[int] = 0) -> "Tensor":
        return BaseOps.squeeze(self, dim)

    @from_op
    def unsqueeze(self, dim: Tuple[int] = 0, axis=None) -> "Tensor":
        # `axis=` accepted as an alias: the reference's own Linear calls unsqueeze(axis=0) on the
        # 3-D path and crashes on it
        return BaseOps.unsqueeze(self, dim if axis is None else axis)

    # ------------------------------------------------------------------ select ops (CPU only)
    @staticmethod
    @from_op
    def where(condition: "Tensor", a: "Tensor", b: "Tensor") -> "Tensor":
        return Elementwise.where(condition, a, b)

    @staticmethod
    @op_gate
    def maximum(a: "Tensor", b: "Tensor") -> "Tensor":
        return Elementwise.maximum(a, b)

    @staticmethod
    @op_gate
    def minimum(a: "Tensor", b: "Tensor") -> "Tensor":
        return Elementwise.minimum(a, b)

    @from_op
    def abs(self) -> "Tensor":
        return Elementwise.abs(self)

    def _const(self, value) -> "Tensor":
        """A scalar constant on THIS tensor's device (the reference builds these on the CPU, Q8)."""
        return data_cast(value, self.device)

    def threshold(self, threshold: float, value: float) -> "Tensor":
        return Tensor.where(self > threshold, self, self._const(value))

    @input_gate
    def masked_fill(self, mask: "Tensor", value: float) -> "Tensor":
        return Tensor.where(mask, self._const(value), self)

    def sign(self) -> "Tensor":
        return Tensor.where(self > 0, self._const(1), Tensor.where(self < 0, self._const(-1), self._const(0)))

    def clip(self, min_value: Optional[float] = None, max_value: Optional[float] = None) -> "Tensor":
        return Tensor.where(self < min_value, self._const(min_value),
                            Tensor.where(self > max_value, self._const(max_value), self))

    # ------------------------------------------------------------------ indexing / comparisons
    @from_op
    def __getitem__(self, index: Index) -> "Tensor":
        return OverloadOps.get_item(self, index)

    @data_gate
    def __lt__(self, other: Data) -> "Tensor":
        return Tensor(self.data < other.data, device=self.device)

    @data_gate
    def __gt__(self, other: Data) -> "Tensor":
        return Tensor(self.data > other.data, device=self.device)

    @data_gate
    def __eq__(self, other: Data) -> "Tensor":
        return Tensor(self.data == other.data, device=self.device)

    @data_gate
    def __le__(self, other: Data) -> "Tensor":
        return Tensor(self.data <= other.data, device=self.device)

    @data_gate
    def __ge__(self, other: Data) -> "Tensor":
        return Tensor(self.data >= other.data, device=self.device)

    @data_gate
    def __ne__(self, other: Data) -> "Tensor":
        return Tensor(self.data != other.data, device=self.device)

    __hash__ = object.__hash__

    # ------------------------------------------------------------------ arithmetic
    @op_gate
    def __add__(self, other: Data) -> "Tensor":
        return OverloadOps.add(self, other)

    @op_gate
    def __radd__(self, other: Data) -> "Tensor":
        return OverloadOps.add(other, self)

    def __iadd__(self, other: Data) -> "Tensor":
        """In place, outside the tape (this is how Linear's bias escapes autograd, Q1)."""
        _tensor(self.device).add(self.data, Tensor.build_ndarray(other, device=self.device), out=self.data)
        return self

    @from_op
    def __neg__(self) -> "Tensor":
        return OverloadOps.neg(self)

    @input_gate
    def __sub__(self, other: Data) -> "Tensor":
        return self + (-other)

    @input_gate
    def __rsub__(self, other: Data) -> "Tensor":
        return other + (-self)

    def __isub__(self, other: Data) -> "Tensor":
        lib = _tensor(self.device)
        rhs = Tensor.build_ndarray(other, device=self.device)
        if self.device == Device.CUDA:
            lib.subtract(self.data, rhs, out=self.data)        # a + (-b) == a - b exactly
        else:
            lib.add(self.data, -rhs, out=self.data)
        return self

    @op_gate
    def __mul__(self, other: Data) -> "Tensor":
        return OverloadOps.mul(self, other)

    @op_gate
    def __rmul__(self, other: Data) -> "Tensor":
        return OverloadOps.mul(other, self)

    def __imul__(self, other: Data) -> "Tensor":
        _tensor(self.device).multiply(self.data, Tensor.build_ndarray(other, device=self.device), out=self.data)
        return self

    @op_gate
    def __matmul__(self, other: Data) -> "Tensor":
        return OverloadOps.matmul(self, other)

    @op_gate
    def __rmatmul__(self, other: Data) -> "Tensor":
        return OverloadOps.matmul(other, self)

    def __pow__(self, pow: int) -> "Tensor":
        return self.pow(pow)

    @input_gate
    def __truediv__(self, other: Data) -> "Tensor":
        return self * (other ** -1)

    @input_gate
    def __rtruediv__(self, other: Data) -> "Tensor":
        return other * (self ** -1)

    def __itruediv__(self, other: Data) -> "Tensor":
        _tensor(self.device).true_divide(self.data, Tensor.build_ndarray(other, device=self.device), out=self.data)
        return self

    # ------------------------------------------------------------------ math / reductions
    @from_op
    def sum(self, axis: int = None, keepdims: bool = False) -> "Tensor":
        return Reduce.sum(self, axis, keepdims)

    @from_op
    def mean(self) -> "Tensor":
        return Reduce.mean(self)

    @from_op
    def max(self, axis: int = None, keepdims: bool = False) -> "Tensor":
        return Reduce.max(self, axis, keepdims)

    @from_op
    def log(self) -> "Tensor":
        return MathOps.log(self)

    @from_op
    def exp(self) -> "Tensor":
        return MathOps.exp(self)

    @from_op
    def pow(self, pow: int) -> "Tensor":
        return MathOps.pow(self, pow)

    @from_op
    def sqrt(self) -> "Tensor":
        return MathOps.pow(self, 0.5)

    @from_op
    def tanh(self) -> "Tensor":
        return MathOps.tanh(self)


# --------------------------------------------------------------------------- backward engine
def _run_backward(root: Tensor, seed: Vector) -> None:
    """Reverse-topological sweep: every node receives the SUM of its incoming gradients once,
    stores it, and pushes `grad_fn(sum)` to its sources.  Iterative (no recursion limit).

    Interior nodes are visited in reversed depth-first post-order.  A node WITHOUT dependencies (a
    Parameter, an input) is finalised the moment its last incoming edge has been processed rather than
    at its place in that order - in the post-order every parameter sits behind ALL interior nodes, which
    would delay the data-parallel trainer's gradient hooks (one all-reduce per weight, trainer.py) until
    the whole backward pass had been queued.  The order in which a node's incoming gradients are summed is
    the same either way."""
    order: List[Tensor] = []
    seen = set()
    fan_in: Dict[int, int] = {}                     # tape edges pointing at each source tensor
    stack: List[Tuple[Tensor, int]] = [(root, 0)]
    while stack:                                    # post-order DFS without recursion
        node, i = stack.pop()
        if i == 0:
            if id(node) in seen:
                continue
            seen.add(id(node))
            for leaf in node._dependencies:
                key = id(leaf.value)
                fan_in[key] = fan_in.get(key, 0) + 1
        deps = node._dependencies
        if i < len(deps):
            stack.append((node, i + 1))
            child = deps[i].value
            if id(child) not in seen:
                stack.append((child, 0))
        else:
            order.append(node)
    pending: Dict[int, Vector] = {id(root): seed}
    backend = backend_for(root.device)
    for node in reversed(order):
        g = pending.pop(id(node), None)
        if g is None:
            continue
        node._receive(g)
        if node._grad_hook is not None:
            node._grad_hook(node)
        for leaf in node._dependencies:
            up = leaf.grad_fn(g)
            src = leaf.value
            if isinstance(up, np.generic):
                up = np.asarray(up)
            key = id(src)
            if key in pending:
                up = backend.accumulate(pending[key], up)
            left = fan_in[key] = fan_in[key] - 1
            if left == 0 and not src._dependencies:
                pending.pop(key, None)              # a dependency-free tensor is final now
                src._receive(up)
                if src._grad_hook is not None:
                    src._grad_hook(src)
            else:
                pending[key] = up
