"""`microtorch_b200.cuda` - the array namespace `_tensor(Device.CUDA)` returns.

It stands where the `cupy` module stands in the reference (src/tensor/device.py:74-84):
the NumPy-API names the reference calls on the module object, all
backed by libmicrotorch_b200.so (reductions incl. max / min, comparisons, where, clip, abs and norm run
on the device).
"""

from __future__ import annotations

import numpy as _np

from .. import _capi
from . import kernels
from .array import DeviceArray, _as_device

ndarray = DeviceArray

float64 = _np.float64
float32 = _np.float32
int64 = _np.int64
int32 = _np.int32
int16 = _np.int16
int8 = _np.int8


def set_float32_matmul_precision(level: str) -> None:
    """"high" | "higher" (default) | "highest": fp32 emulation level of the tensor-core GEMMs (see _capi)."""
    _capi.set_float32_matmul_precision(level)


def is_available() -> bool:
    """True when the shared library loads and at least one CUDA device is visible."""
    return _capi.device_count() > 0


def array(data, dtype=float32) -> DeviceArray:
    if isinstance(data, DeviceArray):
        if _np.dtype(dtype) != data.dtype:
            return data.astype(dtype)
        return data.copy()
    if _np.dtype(dtype) != _np.dtype(_np.float32):
        raise NotImplementedError(f"device arrays are float32; got dtype {_np.dtype(dtype)}")
    return DeviceArray.from_host(data, _np.float32)


asarray = array


def pinned_empty(shape, dtype=float32) -> "_np.ndarray":
    """A NumPy array backed by page-locked host memory (mt_host_alloc): the staging buffer of an input pipeline.
    The memory is released when the array is garbage-collected."""
    import ctypes as C
    import weakref
    shape = (shape,) if isinstance(shape, int) else tuple(shape)
    nbytes = int(_np.prod(shape, dtype=_np.int64)) * _np.dtype(dtype).itemsize
    _capi.require_device()
    p = C.c_void_p()
    nalloc = nbytes if nbytes > 0 else 1        # (`max`/`min` in this namespace are the NumPy-API stubs below)
    _capi.check(_capi.lib().mt_host_alloc(C.byref(p), nalloc), "mt_host_alloc")
    buf = (C.c_char * nalloc).from_address(p.value)
    arr = _np.frombuffer(buf, dtype=dtype, count=int(_np.prod(shape, dtype=_np.int64))).reshape(shape)
    weakref.finalize(buf, _capi.lib().mt_host_free, p.value)
    return arr


def prefetch(host_pinned) -> DeviceArray:
    """Start uploading a pinned float32 host array on the copy stream and return its DeviceArray at once.
    Call `wait_prefetch()` before the first kernel that reads it.  Issue the prefetch of batch i+1 BEFORE launching
    step i: the copy then overlaps step i's compute (see mt_h2d_prefetch in the C header)."""
    host = _np.ascontiguousarray(host_pinned, dtype=_np.float32)
    out = DeviceArray.empty(host.shape)
    if host.nbytes:
        _capi.check(_capi.lib().mt_h2d_prefetch(out.ptr, host.ctypes.data, host.nbytes), "mt_h2d_prefetch")
    return out


def wait_prefetch() -> None:
    _capi.check(_capi.lib().mt_prefetch_wait(), "mt_prefetch_wait")


class random:                      # noqa: N801  (NumPy / CuPy name)
    """Device generator behind Tensor.randn / Tensor.uniform on device="cuda" - where the reference calls
    cupy.random.randn / cupy.random.uniform (src/tensor/tensor.py:402-434).  Philox4x32-10 keyed by `seed`; every call
    consumes ceil(n / 4) blocks of the stream (mt_rng_*_f32 in the C header has the exact definition), so a seeded
    program is reproducible.  Never seeded: the seed is drawn from NumPy's global generator at first use, so
    `np.random.seed(k)` fixes the device stream too."""
    _seed = None
    _offset = 0

    @classmethod
    def seed(cls, seed=None) -> None:
        cls._seed = None if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF
        cls._offset = 0

    @classmethod
    def _state(cls, n):
        if cls._seed is None:
            cls._seed = int(_np.random.randint(0, 2 ** 31 - 1)) | (int(_np.random.randint(0, 2 ** 31 - 1)) << 32)
        off = cls._offset
        cls._offset += (n + 3) // 4
        return cls._seed, off

    @classmethod
    def randn(cls, *dims) -> DeviceArray:
        out = DeviceArray.empty(tuple(int(d) for d in dims))
        seed, off = cls._state(out.size)
        _capi.check(_capi.lib().mt_rng_normal_f32(out.ptr, out.size, seed, off), "mt_rng_normal_f32")
        return out

    @classmethod
    def uniform(cls, low=0.0, high=1.0, size=None) -> DeviceArray:
        shape = () if size is None else ((int(size),) if _np.isscalar(size) else tuple(int(d) for d in size))
        out = DeviceArray.empty(shape)
        seed, off = cls._state(out.size)
        _capi.check(_capi.lib().mt_rng_uniform_f32(out.ptr, out.size, seed, off, float(low), float(high)), "mt_rng_uniform_f32")
        return out


def zeros_like(x, dtype=None) -> DeviceArray:
    return DeviceArray.zeros(x.shape)


def ones_like(x, dtype=None) -> DeviceArray:
    out = DeviceArray.empty(x.shape)
    out.fill(1.0)
    return out


def zeros(shape, dtype=float32) -> DeviceArray:
    return DeviceArray.zeros((shape,) if isinstance(shape, int) else shape)


def _out(op, a, b, out):
    if out is None:
        return kernels.binary(op, a, b)
    if out is a:
        return kernels.inplace(op, out, b)
    return kernels.binary(op, a, b, out=out)


def add(a, b, out=None):
    return _out(_capi.BIN_ADD, a, b, out)


def multiply(a, b, out=None):
    return _out(_capi.BIN_MUL, a, b, out)


def true_divide(a, b, out=None):
    return _out(_capi.BIN_DIV, a, b, out)


def subtract(a, b, out=None):
    return _out(_capi.BIN_SUB, a, b, out)


def log(x):
    return kernels.unary(_capi.UN_LOG, x)


def exp(x):
    return kernels.unary(_capi.UN_EXP, x)


def tanh(x):
    return kernels.unary(_capi.UN_TANH, x)


def sum(x, axis=None, keepdims=False):      # noqa: A001
    return kernels.sum(x, axis, keepdims)


def transpose(x, axes=None):
    return _as_device(x).transpose(axes)


def squeeze(x, axis=None):
    return _as_device(x).squeeze(axis)


def expand_dims(x, axis):
    return _as_device(x).expand_dims(axis)


def broadcast_to(x, shape):
    return _as_device(x).broadcast_to(shape)


def outer(a, b):
    a, b = _as_device(a), _as_device(b)
    return kernels.binary(_capi.BIN_MUL, a.reshape((-1, 1)), b.reshape((1, -1)))


def matmul(a, b):
    return kernels.matmul(a, b)


def _unsupported(name):
    def fn(*_a, **_k):
        raise NotImplementedError(
            f"{name} is not implemented on the B200 backend (outside the training-step path; the "
            "reference's CuPy path cannot run it either, SURVEY.md Q8).  There is no CPU fallback.")
    fn.__name__ = name
    return fn


def where(cond, a, b):
    return kernels.where(cond, a, b)


def maximum(a, b):
    return kernels.binary(_capi.BIN_MAX, a, b)


def minimum(a, b):
    return kernels.binary(_capi.BIN_MIN, a, b)


def abs(x):                        # noqa: A001
    return kernels.unary(_capi.UN_ABS, x)


def clip(x, lo, hi):
    out = x
    if lo is not None:
        out = kernels.binary(_capi.BIN_MAX, out, float(lo))
    if hi is not None:
        out = kernels.binary(_capi.BIN_MIN, out, float(hi))
    return out


class linalg:                      # noqa: N801  (NumPy name)
    @staticmethod
    def norm(x, ord=2.0):          # noqa: A002
        """np.linalg.norm(x, ord) as a host float (the reference compares it on the host, tensor.py:308-309), with
        NumPy's meaning of `ord` per rank (Tensor.clip_grad_norm, reference src/tensor/tensor.py:297-310):
          1-D  vector p-norm                                             - device reductions
          2-D  ord=1 / inf: largest absolute column / row sum; 'fro'     - device reductions
               ord=2: the SPECTRAL norm (largest singular value).  There is no SVD on the device: the gradient is
               downloaded and NumPy's own SVD gives the value, so the scale factor is the reference's bit for bit.
               Only this one scalar is computed on the host; the clipping multiply stays on the device.
          other ranks raise like NumPy does when `ord` is given."""
        x = _as_device(x)
        if x.ndim not in (1, 2):
            raise ValueError("Improper number of dimensions to norm.")
        axes = tuple(range(x.ndim))
        if x.ndim == 2:
            if ord == 2 or ord == -2:
                return float(_np.linalg.norm(x.get(), ord))
            if ord in (1, -1, _np.inf, -_np.inf):
                sums = kernels.sum(kernels.unary(_capi.UN_ABS, x), axis=0 if ord in (1, -1) else 1)
                return float(kernels.extreme(sums, None, False, ord in (1, _np.inf)).item())
            if ord in ("fro", None):
                return float(kernels.reduce_sum(x, axes, x).item()) ** 0.5
            raise ValueError("Invalid norm order for matrices.")
        if ord in (2, None):
            return float(kernels.reduce_sum(x, axes, x).item()) ** 0.5        # fused x*x reduction
        if ord == 1:
            return float(kernels.reduce_sum(kernels.unary(_capi.UN_ABS, x), axes).item())
        if ord in (_np.inf, -_np.inf):
            return float(kernels.extreme(kernels.unary(_capi.UN_ABS, x), None, False, ord == _np.inf).item())
        p = float(ord)
        powed = kernels.unary(_capi.UN_POW, kernels.unary(_capi.UN_ABS, x), p)
        return float(kernels.reduce_sum(powed, axes).item()) ** (1.0 / p)


def max(x, axis=None, keepdims=False):          # noqa: A001 (NumPy name)
    return kernels.extreme(x, axis, keepdims, True)


def min(x, axis=None, keepdims=False):          # noqa: A001 (NumPy name)
    return kernels.extreme(x, axis, keepdims, False)
