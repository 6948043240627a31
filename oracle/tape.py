"""Oracle autograd tape: a NumPy restatement of the reference's Tensor + ops.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Citations are
/root/reference paths.  The restatement keeps the reference's *work* as well as
its results - eager zero-filled grads, accumulate-by-allocation, recursive
depth-first backward without a topological sort - because bench.py times this
module as the "port" CPU baseline.

Repairs applied ("reference-repaired" behaviour):
  S1  a NumPy scalar (np.generic) arriving as a gradient is turned back into a
      0-d ndarray, which is what the CuPy backend the notebook ran on does
      (src/tensor/tensor.py:266-270 rejects it on NumPy).
  S3  batched matmul gradients are un-broadcast onto 2-D operands
      (src/tensor/ops/overload.py:141-153 forgets to), only when `repair_3d`.
"""

from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple, Union

import numpy as np

F32 = np.float32
Array = np.ndarray


def _as_f32(data) -> Array:
    """Tensor.build_ndarray (src/tensor/tensor.py:186-205) for DType.FLOAT32 on CPU:
    a Var passes its array through, an ndarray is cast only when the dtype differs,
    anything else goes through np.array(dtype=float32)."""
    if isinstance(data, Var):
        return data.data
    if isinstance(data, np.ndarray):
        return data if data.dtype == F32 else data.astype(F32)
    return np.array(data, dtype=F32)


def unbroadcast(grad: Array, shape: Tuple[int, ...]) -> Array:
    """BaseOps.bkwd_broadcast (src/tensor/ops/base.py:8-58): fold a gradient of the
    broadcast result back onto an operand of `shape`."""
    ndim = len(shape)
    if ndim == 0:                       # base.py:28-29  scalar operand: everything sums
        return grad.sum()
    if grad.ndim == 0:                  # base.py:32-33  scalar grad passes through
        return grad
    lead = max(0, grad.ndim - ndim)     # base.py:38-41  prepended dims
    if lead > 0:
        grad = grad.sum(axis=tuple(range(lead)), keepdims=False)
    ones = tuple(d for d in range(ndim) if shape[d] == 1 and grad.shape[d] > 1)
    if ones:                            # base.py:44-50  size-1 dims, keepdims
        grad = grad.sum(axis=ones, keepdims=True)
    if grad.shape != tuple(shape):      # base.py:53-54
        grad = grad.reshape(shape)
    return grad


GradFn = Callable[[Array], Array]


class Var:
    """Restatement of src/tensor/tensor.py::Tensor restricted to the hot path."""

    __slots__ = ("data", "requires_grad", "parents", "grad")

    #: statistics the bench/tests read: number of Var constructions, backward visits
    stats = {"ctor": 0, "backward": 0}

    def __init__(self, data, requires_grad: bool = False,
                 parents: Optional[List[Tuple["Var", GradFn]]] = None) -> None:
        # tensor.py:58-85 - storage, then an eagerly allocated AND re-zeroed grad
        self.data = _as_f32(data)
        self.requires_grad = requires_grad
        self.parents = parents or []
        if requires_grad:
            self.grad = np.zeros_like(self.data, dtype=F32)   # tensor.py:77
            self.grad.fill(0.0)                               # tensor.py:84-85 -> :248
        else:
            self.grad = None
        Var.stats["ctor"] += 1

    # -- metadata ---------------------------------------------------------
    @property
    def shape(self):
        return self.data.shape

    @property
    def ndim(self):
        return self.data.ndim

    @property
    def size(self):
        return self.data.size

    def zero_grad(self) -> None:        # tensor.py:236-248
        if self.requires_grad:
            if self.grad is None:
                self.grad = np.zeros_like(self.data, dtype=F32)
            else:
                self.grad.fill(0.0)

    # -- backward ---------------------------------------------------------
    def backward(self, grad: Optional[Array] = None) -> None:
        """tensor.py:256-283.  Depth-first, one recursive call per edge."""
        if not self.requires_grad:
            raise ValueError("Backward was called on a non-required-grad tensor!")
        if grad is None:
            if self.shape == ():
                grad = _as_f32(1.0)     # tensor.py:260-262
            else:
                raise ValueError("Grad must be provided for non-scalar tensors. "
                                 f"Tensor shape: {self.shape}")
        if isinstance(grad, np.generic):            # repair S1
            grad = np.asarray(grad)
        if not isinstance(grad, np.ndarray):        # tensor.py:266-270
            raise ValueError("Grad device does not match tensor device")
        if grad.shape != self.shape:                # tensor.py:273-274
            raise ValueError(f"Grad shape {grad.shape} does not match tensor shape {self.shape}")
        Var.stats["backward"] += 1
        if self.grad is None:
            self.grad = grad
        else:
            self.grad = np.add(self.grad, grad)     # tensor.py:279 (fresh allocation)
        for parent, fn in self.parents:             # tensor.py:281-283
            parent.backward(fn(grad))

    # -- arithmetic (src/tensor/ops/overload.py) -------------------------------
    @staticmethod
    def _lift(x) -> "Var":              # data_cast, tensor.py:14-17
        return x if isinstance(x, Var) else Var(x)

    def __add__(self, other) -> "Var":  # overload.py:68-90
        other = Var._lift(other)
        out = self.data + other.data
        parents = []
        if self.requires_grad:
            parents.append((self, lambda g, s=self.shape: unbroadcast(g, s)))
        if other.requires_grad:
            parents.append((other, lambda g, s=other.shape: unbroadcast(g, s)))
        return Var(out, self.requires_grad or other.requires_grad, parents)

    def __radd__(self, other) -> "Var":  # tensor.py:543-545 -> add(other, self)
        return Var._lift(other) + self

    def __neg__(self) -> "Var":         # overload.py:50-66
        parents = [(self, lambda g: -g)] if self.requires_grad else []
        return Var(-self.data, self.requires_grad, parents)

    def __sub__(self, other) -> "Var":  # tensor.py:556-558   a - b := a + (-b)
        return self + (-Var._lift(other))

    def __rsub__(self, other) -> "Var":  # tensor.py:560-562  b - a := b + (-a)
        return Var._lift(other) + (-self)

    def __mul__(self, other) -> "Var":  # overload.py:92-128
        other = Var._lift(other)
        out = self.data * other.data
        parents = []
        if self.requires_grad:
            parents.append((self, lambda g, o=other, s=self.shape: unbroadcast(g * o.data, s)))
        if other.requires_grad:
            parents.append((other, lambda g, o=self, s=other.shape: unbroadcast(g * o.data, s)))
        return Var(out, self.requires_grad or other.requires_grad, parents)

    def __rmul__(self, other) -> "Var":  # tensor.py:577-579 -> mul(other, self)
        return Var._lift(other) * self

    def __truediv__(self, other) -> "Var":  # tensor.py:601-603  a / b := a * b**-1
        return self * (Var._lift(other) ** -1)

    def __pow__(self, p) -> "Var":      # src/tensor/ops/math.py:44-62
        x = self.data
        parents = []
        if self.requires_grad:
            parents.append((self, lambda g: g * (p * (x ** (p - 1)))))   # math.py:51
        return Var(x ** p, self.requires_grad, parents)

    def log(self) -> "Var":             # math.py:6-23
        x = self.data
        parents = [(self, lambda g: g / x)] if self.requires_grad else []
        return Var(np.log(x), self.requires_grad, parents)

    def tanh(self) -> "Var":            # math.py:68-85 (backward uses the saved OUTPUT)
        y = np.tanh(self.data)
        parents = [(self, lambda g: g * (1 - y ** 2))] if self.requires_grad else []
        return Var(y, self.requires_grad, parents)

    def exp(self) -> "Var":             # math.py:25-42
        y = np.exp(self.data)
        parents = [(self, lambda g: g * y)] if self.requires_grad else []
        return Var(y, self.requires_grad, parents)

    # -- reductions (src/tensor/ops/reduce.py) -------------------------------
    def sum(self, axis=None, keepdims: bool = False) -> "Var":   # reduce.py:8-38
        x = self.data
        out = np.sum(x, axis=axis, keepdims=keepdims)
        parents = []
        if self.requires_grad:
            def bwd(g):
                full = np.ones_like(x)                  # reduce.py:16 (materialised)
                if axis is None:
                    return full * g                     # reduce.py:19
                ge = g if keepdims else np.expand_dims(g, axis=axis)
                return full * ge                        # reduce.py:26
            parents.append((self, bwd))
        return Var(out, self.requires_grad, parents)

    def mean(self) -> "Var":            # reduce.py:40-43 via tensor.py:626-628 (no axis)
        return self.sum(axis=None, keepdims=False) / self.size

    # -- matmul (overload.py:130-164) --------------------------------------
    def __matmul__(self, other: "Var") -> "Var":
        return _matmul(self, other, repair_3d=False)

    # -- shape ops (src/tensor/ops/base.py) ----------------------------------
    def transpose(self, axes=None) -> "Var":    # base.py:107-136
        parents = []
        if self.requires_grad:
            if axes is None:
                parents.append((self, lambda g: np.transpose(g)))
            else:
                inv = tuple(np.argsort(axes))
                parents.append((self, lambda g: np.transpose(g, axes=inv)))
        return Var(np.transpose(self.data, axes=axes), self.requires_grad, parents)

    @property
    def T(self) -> "Var":
        return self.transpose()

    def view(self, shape) -> "Var":     # base.py:82-105
        src = self.shape
        parents = [(self, lambda g: g.reshape(src))] if self.requires_grad else []
        return Var(self.data.reshape(shape), self.requires_grad, parents)

    reshape = view

    def squeeze(self, dim=0) -> "Var":  # base.py:138-161
        src = self.shape
        parents = []
        if self.requires_grad:
            parents.append((self, (lambda g: g.reshape(src)) if dim is None
                            else (lambda g: np.expand_dims(g, axis=dim))))
        return Var(np.squeeze(self.data, axis=dim), self.requires_grad, parents)

    def unsqueeze(self, dim=0) -> "Var":  # base.py:163-187
        parents = [(self, lambda g: np.squeeze(g, axis=dim))] if self.requires_grad else []
        return Var(np.expand_dims(self.data, axis=dim), self.requires_grad, parents)

    def broadcast_to(self, shape) -> "Var":  # base.py:60-80
        parents = ([(self, lambda g, s=self.shape: unbroadcast(g, s))]
                   if self.requires_grad else [])
        return Var(np.broadcast_to(self.data, shape), self.requires_grad, parents)

    def __getitem__(self, index) -> "Var":  # overload.py:9-48
        if hasattr(index, "data") and not isinstance(index, np.ndarray):
            index = index.data
        if isinstance(index, list):
            index = np.array(index)
        if isinstance(index, tuple) and len(index) == 1:
            index = index[0]
        x = self.data
        parents = []
        if self.requires_grad:
            def bwd(g):
                full = np.zeros_like(x)
                full[index] = g          # ASSIGN, duplicates do not accumulate (Q10)
                return full
            parents.append((self, bwd))
        return Var(x[index], self.requires_grad, parents)

    # -- in-place, no tape (tensor.py:547-571) --------------------------------
    def iadd_(self, other) -> "Var":
        np.add(self.data, _as_f32(other), out=self.data)
        return self

    def isub_(self, other) -> "Var":
        neg = -_as_f32(other)            # tensor.py:569 (temp)
        np.add(self.data, neg, out=self.data)
        return self


def _matmul(a: Var, b: Var, repair_3d: bool) -> Var:
    """OverloadOps.matmul (overload.py:130-164); repair S3 when `repair_3d`."""
    out = a.data @ b.data
    parents = []
    if a.requires_grad:
        def bkwd_a(g):
            if b.ndim > 1:
                r = g @ b.data.swapaxes(-1, -2)           # overload.py:142
            else:
                r = np.outer(g, b.data.T).squeeze()       # overload.py:143
            return unbroadcast(r, a.shape) if repair_3d else r
        parents.append((a, bkwd_a))
    if b.requires_grad:
        def bkwd_b(g):
            if a.ndim > 1:
                r = a.data.swapaxes(-1, -2) @ g           # overload.py:153
            else:
                r = np.outer(a.data.T, g).squeeze()       # overload.py:154
            return unbroadcast(r, b.shape) if repair_3d else r
        parents.append((b, bkwd_b))
    return Var(out, a.requires_grad or b.requires_grad, parents)


def matmul(a: Var, b: Var, repair_3d: bool = False) -> Var:
    return _matmul(a, b, repair_3d)
