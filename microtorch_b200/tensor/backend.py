"""Op-granular compute backends behind the autograd ops.

The reference writes every op once against the NumPy API and lets `_tensor(device)` swap
the array module underneath (one library kernel + one allocation per array expression).
Here the ops layer (tensor/ops/*) instead talks to a backend at OP granularity - forward
kernels and *fused* backward kernels - so the CUDA side can answer each op with the minimum
number of launches:

    tanh backward      g*(1-y^2)            1 kernel  (reference math.py:74: 3 + 2 temporaries)
    pow  backward      g*(p*x^(p-1))        1 kernel  (math.py:51: 3)
    mul  backward      unbroadcast(g*other) 1 kernel, the product never reaches HBM
    sum  backward      a stride-0 VIEW of g, no ones_like / multiply (reduce.py:16-26)

`CpuBackend` is the Device.CPU implementation (NumPy, the reference's arithmetic op for
op); `CudaBackend` is Device.CUDA (C ABI only).  They are never mixed: a CUDA tensor that
asks for something the CUDA backend lacks gets NotImplementedError, not a CPU detour.
"""

from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np

from .. import _capi
from .device import Device


# =============================================================================== CPU (NumPy)
class CpuBackend:
    device = Device.CPU
    ndarray = np.ndarray

    # -- forward ------------------------------------------------------------------------
    @staticmethod
    def add(a, b): return a + b
    @staticmethod
    def mul(a, b): return a * b
    @staticmethod
    def neg(a): return -a
    @staticmethod
    def log(x): return np.log(x)
    @staticmethod
    def exp(x): return np.exp(x)
    @staticmethod
    def tanh(x): return np.tanh(x)
    @staticmethod
    def pow(x, p): return x ** p
    @staticmethod
    def matmul(a, b): return a @ b
    @staticmethod
    def sum(x, axis, keepdims): return np.sum(x, axis=axis, keepdims=keepdims)
    @staticmethod
    def extreme(x, axis, keepdims, is_max): return (np.max if is_max else np.min)(x, axis=axis, keepdims=keepdims)
    @staticmethod
    def equal(a, b): return a == b
    @staticmethod
    def div(a, b): return a / b

    # -- backward -----------------------------------------------------------------------
    @staticmethod
    def unbroadcast(grad, shape, other=None):
        """reference base.py:8-58, optionally after `grad * other` (overload.py:109)."""
        if other is not None:
            grad = grad * other
        shape = tuple(shape)
        if len(shape) == 0:
            return grad.sum()
        if grad.ndim == 0:
            return grad
        lead = max(0, grad.ndim - len(shape))
        if lead:
            grad = grad.sum(axis=tuple(range(lead)), keepdims=False)
        axes = tuple(d for d in range(len(shape)) if shape[d] == 1 and grad.shape[d] > 1)
        if axes:
            grad = grad.sum(axis=axes, keepdims=True)
        return grad if grad.shape == shape else grad.reshape(shape)

    @staticmethod
    def log_bwd(g, x): return g / x
    @staticmethod
    def exp_bwd(g, y): return g * y
    @staticmethod
    def tanh_bwd(g, y): return g * (1 - y ** 2)
    @staticmethod
    def pow_bwd(g, x, p): return g * (p * (x ** (p - 1)))

    @staticmethod
    def sum_bwd(g, like, axis, keepdims):
        full = np.ones_like(like)
        if axis is None:
            return full * g
        return full * (g if keepdims else np.expand_dims(g, axis=axis))

    @staticmethod
    def matmul_bwd_a(g, a, b):
        if b.ndim > 1:
            return g @ b.swapaxes(-1, -2)
        return np.outer(g, b.T).squeeze()

    @staticmethod
    def matmul_bwd_b(g, a, b):
        if a.ndim > 1:
            return a.swapaxes(-1, -2) @ g
        return np.outer(a.T, g).squeeze()

    # -- shape / index ------------------------------------------------------------------
    @staticmethod
    def transpose(x, axes=None): return np.transpose(x, axes=axes)
    @staticmethod
    def reshape(x, shape): return x.reshape(shape)
    @staticmethod
    def squeeze(x, axis): return np.squeeze(x, axis=axis)
    @staticmethod
    def expand_dims(x, axis): return np.expand_dims(x, axis=axis)
    @staticmethod
    def broadcast_to(x, shape): return np.broadcast_to(x, shape)
    @staticmethod
    def getitem(x, index): return x[index]

    @staticmethod
    def getitem_bwd(g, like, index):
        full = np.zeros_like(like)
        full[index] = g
        return full

    @staticmethod
    def where(c, a, b): return np.where(c, a, b)

    # -- grads --------------------------------------------------------------------------
    @staticmethod
    def zeros_like(x, dtype=None): return np.zeros_like(x, dtype=dtype)
    @staticmethod
    def accumulate(acc, g): return np.add(acc, g)


# =============================================================================== CUDA (C ABI)
class CudaBackend:
    device = Device.CUDA

    def __init__(self) -> None:
        from ..cuda import kernels, DeviceArray
        self.K = kernels
        self.ndarray = DeviceArray

    # -- forward ------------------------------------------------------------------------
    def add(self, a, b): return self.K.binary(_capi.BIN_ADD, a, b)
    def mul(self, a, b): return self.K.binary(_capi.BIN_MUL, a, b)
    def neg(self, a): return self.K.unary(_capi.UN_NEG, a)
    def log(self, x): return self.K.unary(_capi.UN_LOG, x)
    def exp(self, x): return self.K.unary(_capi.UN_EXP, x)
    def tanh(self, x): return self.K.unary(_capi.UN_TANH, x)
    def pow(self, x, p): return self.K.unary(_capi.UN_POW, x, float(p))
    def matmul(self, a, b): return self.K.matmul(a, b)
    def sum(self, x, axis, keepdims): return self.K.sum(x, axis, keepdims)
    def extreme(self, x, axis, keepdims, is_max): return self.K.extreme(x, axis, keepdims, is_max)
    def equal(self, a, b): return self.K.binary(_capi.BIN_EQ, a, b)
    def div(self, a, b): return self.K.binary(_capi.BIN_DIV, a, b)

    # -- backward (fused) ---------------------------------------------------------------
    def unbroadcast(self, grad, shape, other=None): return self.K.unbroadcast(grad, shape, other)
    def log_bwd(self, g, x): return self.K.binary(_capi.BIN_DIV, g, x)
    def exp_bwd(self, g, y): return self.K.binary(_capi.BIN_EXP_BWD, g, y)
    def tanh_bwd(self, g, y): return self.K.binary(_capi.BIN_TANH_BWD, g, y)
    def pow_bwd(self, g, x, p): return self.K.binary(_capi.BIN_POW_BWD, g, x, float(p))

    def sum_bwd(self, g, like, axis, keepdims):
        # ones_like(x) * g  ==  g broadcast to x.shape: a stride-0 view, zero bytes
        if axis is not None and not keepdims:
            g = g.expand_dims(axis)
        return g.broadcast_to(like.shape)

    def matmul_bwd_a(self, g, a, b):
        if b.ndim > 1:
            return self.K.matmul(g, b.swapaxes(-1, -2))
        from .. import cuda
        return cuda.outer(g, b).squeeze()

    def matmul_bwd_b(self, g, a, b):
        if a.ndim > 1:
            return self.K.matmul(a.swapaxes(-1, -2), g)
        from .. import cuda
        return cuda.outer(a, g).squeeze()

    # -- shape / index (host metadata only) -----------------------------------------------
    def transpose(self, x, axes=None): return x.transpose(axes)
    def reshape(self, x, shape): return x.reshape(shape)
    def squeeze(self, x, axis): return x.squeeze(axis)
    def expand_dims(self, x, axis): return x.expand_dims(axis)
    def broadcast_to(self, x, shape): return x.broadcast_to(shape)
    def getitem(self, x, index): return x[index]

    def getitem_bwd(self, g, like, index):
        full = self.ndarray.zeros(like.shape)
        full[index] = g
        return full

    def where(self, c, a, b): return self.K.where(c, a, b)

    # -- grads --------------------------------------------------------------------------
    def zeros_like(self, x, dtype=None): return self.ndarray.zeros(x.shape)
    def accumulate(self, acc, g): return self.K.binary(_capi.BIN_ADD, acc, g)


_CPU = CpuBackend()
_CUDA: Optional[CudaBackend] = None


def backend_for(device: Device):
    """The compute backend of `device`.  Device.CUDA needs the shared library and a GPU."""
    global _CUDA
    if device == Device.CUDA:
        if _CUDA is None:
            _capi.require_device()
            _CUDA = CudaBackend()
        return _CUDA
    return _CPU
