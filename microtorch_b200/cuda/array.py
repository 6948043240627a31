"""DeviceArray: the ndarray of the B200 backend (what cupy.ndarray is to the reference).

A DeviceArray is a strided fp32 view (pointer + shape + element strides) of a block owned
by the library's caching pool (`mt_malloc`).  Metadata operations - reshape of a dense
array, transpose, squeeze, expand_dims, basic slicing, broadcast_to - are done on the host
and cost zero bytes (bit-exact by construction).  Everything that moves
bytes is a C-ABI call; nothing here computes on the CPU.

Surface = the array-object API the reference calls on `tensor.data` / `grad`
(SURVEY.md Appendix D): shape/ndim/size/dtype/T, reshape, sum, swapaxes, squeeze, astype,
copy, fill, item, get, __getitem__/__setitem__, + * @ ** / unary -, and comparisons (float32
0/1 masks that stay on the device - the reference moves them to the CPU, Q8).
"""

from __future__ import annotations

import ctypes as C
import functools
import math
from typing import Optional, Sequence, Tuple, Union

import numpy as np

from .. import _capi

F32 = np.dtype(np.float32)
I64 = np.dtype(np.int64)


class _Block:
    """Owner of one pool allocation; returned to the pool when the last view dies."""

    __slots__ = ("ptr", "nbytes")

    def __init__(self, nbytes: int) -> None:
        self.ptr = None
        if not _capi._DEVICE_READY:
            _capi.require_device()
        p = C.c_void_p()
        status = _capi._LIB.mt_malloc(C.byref(p), nbytes if nbytes > 0 else 1)
        if status:
            _capi.check(status, "mt_malloc")
        self.ptr = p.value
        self.nbytes = nbytes

    def __del__(self) -> None:
        ptr, self.ptr = getattr(self, "ptr", None), None
        if ptr:
            try:
                _capi._LIB.mt_free(ptr)
            except Exception:       # interpreter teardown
                pass


@functools.lru_cache(maxsize=8192)
def _dense_strides_cached(shape: tuple) -> Tuple[int, ...]:
    st, acc = [], 1
    for d in reversed(shape):
        st.append(acc)
        acc *= max(int(d), 1)
    return tuple(reversed(st))


def _dense_strides(shape: Sequence[int]) -> Tuple[int, ...]:
    return _dense_strides_cached(shape if type(shape) is tuple else tuple(shape))


def _int_tuple(seq) -> Tuple[int, ...]:
    """Shape / stride tuple of Python ints (no work when it already is one: the per-launch host path)."""
    if type(seq) is tuple:
        for d in seq:
            if type(d) is not int:
                break
        else:
            return seq
    return tuple(int(d) for d in seq)


def _numel(shape: Sequence[int]) -> int:
    n = 1
    for d in shape:
        n *= int(d)
    return n


class DeviceArray:
    __slots__ = ("_block", "ptr", "shape", "strides", "dtype", "_host_value", "_owned", "_const", "_dense", "__weakref__")
    __array_priority__ = 1000       # numpy scalars defer to our reflected operators

    def __init__(self, block: _Block, ptr: int, shape: Sequence[int], strides: Sequence[int],
                 dtype: np.dtype = F32) -> None:
        self._block = block
        self.ptr = ptr
        self.shape = _int_tuple(shape)
        self.strides = _int_tuple(strides)                  # in ELEMENTS
        self.dtype = dtype if dtype is F32 else np.dtype(dtype)
        self._host_value = None     # host mirror of a 0-d constant built from a Python scalar
        self._owned = False         # adopted as some tensor's .grad (see Tensor._receive)
        self._const = False         # shared read-only constant (device_constant): never adopted, never written
        self._dense = None          # cached is_dense (shape and strides never change after construction)

    # ------------------------------------------------------------------ constructors
    @classmethod
    def empty(cls, shape: Sequence[int], dtype: np.dtype = F32) -> "DeviceArray":
        shape = _int_tuple(shape)
        if dtype is not F32:
            dtype = np.dtype(dtype)
        blk = _Block(_numel(shape) * dtype.itemsize)
        return cls(blk, blk.ptr, shape, _dense_strides_cached(shape), dtype)

    @classmethod
    def zeros(cls, shape: Sequence[int], dtype: np.dtype = F32) -> "DeviceArray":
        out = cls.empty(shape, dtype)
        if out.nbytes:
            _capi.check(_capi.lib().mt_memset_zero(out.ptr, out.nbytes), "mt_memset_zero")
        return out

    @classmethod
    def from_host(cls, data, dtype: np.dtype = F32) -> "DeviceArray":
        """H2D upload (cp.array(np_array), reference tensor.py:205,221)."""
        dtype = np.dtype(dtype)
        if dtype not in (F32, I64):
            raise NotImplementedError(f"device arrays are float32 (int64 for indices); got {dtype}")
        scalar = None
        if isinstance(data, (int, float)) and not isinstance(data, bool):
            scalar = float(np.float32(data)) if dtype == F32 else None
        host = np.asarray(data, dtype=dtype)
        if not host.flags.c_contiguous:      # (np.ascontiguousarray would turn a 0-d scalar into shape (1,))
            host = np.array(host, dtype=dtype, order="C")
        out = cls.empty(host.shape, dtype)
        if host.nbytes:
            _capi.check(_capi.lib().mt_h2d(out.ptr, host.ctypes.data, host.nbytes), "mt_h2d")
        out._host_value = scalar
        return out

    # ------------------------------------------------------------------ metadata
    @property
    def ndim(self) -> int:
        return len(self.shape)

    @property
    def size(self) -> int:
        return _numel(self.shape)

    @property
    def nbytes(self) -> int:
        return self.size * self.dtype.itemsize

    @property
    def itemsize(self) -> int:
        return self.dtype.itemsize

    @property
    def is_dense(self) -> bool:
        """C-contiguous (extent-1 dims are free)."""
        dense = self._dense
        if dense is None:
            dense, acc = True, 1
            for d, s in zip(reversed(self.shape), reversed(self.strides)):
                if d != 1 and s != acc:
                    dense = False
                    break
                acc *= d
            self._dense = dense
        return dense

    def _view(self, shape, strides, ptr=None) -> "DeviceArray":
        v = DeviceArray(self._block, self.ptr if ptr is None else ptr, shape, strides, self.dtype)
        v._const = self._const          # a view of a shared read-only constant is read-only too
        return v

    def dense(self) -> "DeviceArray":
        """self if already C-contiguous, else a materialised copy (one strided-gather kernel)."""
        return self if self.is_dense else self.copy()

    # ------------------------------------------------------------------ host transfer
    def get(self) -> np.ndarray:
        """D2H download (cupy `.get()`, reference tensor.py:218).  Blocks."""
        src = self.dense()
        out = np.empty(self.shape, self.dtype)
        if out.nbytes:
            _capi.check(_capi.lib().mt_d2h(out.ctypes.data, src.ptr, out.nbytes), "mt_d2h")
        return out

    def __array__(self, dtype=None, copy=None):
        host = self.get()
        return host if dtype is None else host.astype(dtype)

    def item(self):
        if self.size != 1:
            raise ValueError("can only convert an array of size 1 to a Python scalar")
        return self.get().reshape(()).item()

    def __float__(self) -> float:
        return float(self.item())

    def __repr__(self) -> str:
        return f"DeviceArray({np.array2string(self.get(), separator=', ')}, dtype={self.dtype.name})"

    def __str__(self) -> str:
        """Like the host array (the notebook prints `data_loss.data` inside an f-string, demo.ipynb:5186)."""
        return str(self.get())

    def __format__(self, spec: str) -> str:
        return format(self.get(), spec)

    def __len__(self) -> int:
        if not self.shape:
            raise TypeError("len() of unsized object")
        return self.shape[0]

    # ------------------------------------------------------------------ copies / fills
    def copy(self) -> "DeviceArray":
        out = DeviceArray.empty(self.shape, self.dtype)
        if not out.nbytes:
            return out
        if self.is_dense:
            _capi.check(_capi.lib().mt_d2d(out.ptr, self.ptr, out.nbytes), "mt_d2d")
        elif self.dtype == F32 and any(s == 0 and d > 1 for d, s in zip(self.shape, self.strides)):
            # broadcast view: x * 1 through the broadcasting kernel (vectorised row/column/scalar paths, exact)
            from . import kernels as K
            K.binary(_capi.BIN_MUL, self, _one(), out=out)
        else:
            self._require_f32("copy of a strided view")
            _capi.check(_capi.lib().mt_copy_strided_f32(out.ptr, self.ptr, self.ndim, _capi.i64(self.shape),
                                                        _capi.i64(self.strides)), "mt_copy_strided_f32")
        return out

    def astype(self, dtype) -> "DeviceArray":
        if np.dtype(dtype) == self.dtype:
            return self.copy()
        raise NotImplementedError(f"astype({np.dtype(dtype)}) on the device: only float32 is computed on (cast on the host)")

    def fill(self, value: float) -> None:
        self._require_f32("fill")
        if self._const:
            raise ValueError("this array is a shared read-only constant")
        if not self.size:
            return
        if not self.is_dense:
            raise NotImplementedError("fill() on a strided view")
        if value == 0:
            _capi.check(_capi.lib().mt_memset_zero(self.ptr, self.nbytes), "mt_memset_zero")
        else:
            _capi.check(_capi.lib().mt_fill_f32(self.ptr, self.size, float(value)), "mt_fill_f32")
        self._host_value = None

    def _require_f32(self, what: str) -> None:
        if self.dtype != F32:
            raise NotImplementedError(f"{what}: only float32 device arrays are computed on")

    # ------------------------------------------------------------------ shape ops (zero bytes)
    def reshape(self, *shape) -> "DeviceArray":
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        elif len(shape) == 1 and shape[0] is None:
            shape = ()
        shape = _int_tuple(shape)
        if shape == self.shape:
            return self
        n = self.size
        if -1 in shape:
            known = -_numel(shape)
            if known == 0 or n % known:
                raise ValueError(f"cannot reshape array of size {n} into shape {shape}")
            shape = tuple(n // known if d == -1 else d for d in shape)
        if _numel(shape) != n:
            raise ValueError(f"cannot reshape array of size {n} into shape {shape}")
        if shape == self.shape:
            return self
        src = self.dense()                      # NumPy copies too when a view cannot express it
        return src._view(shape, _dense_strides(shape))

    def transpose(self, *axes) -> "DeviceArray":
        if len(axes) == 1 and (axes[0] is None or isinstance(axes[0], (tuple, list))):
            axes = axes[0]
        if axes is None or len(axes) == 0:
            axes = tuple(reversed(range(self.ndim)))
        axes = tuple(int(a) % self.ndim if self.ndim else 0 for a in axes)
        if sorted(axes) != list(range(self.ndim)):
            raise ValueError("axes don't match array")
        return self._view([self.shape[a] for a in axes], [self.strides[a] for a in axes])

    @property
    def T(self) -> "DeviceArray":
        return self.transpose()

    def swapaxes(self, a: int, b: int) -> "DeviceArray":
        ax = list(range(self.ndim))
        ax[a], ax[b] = ax[b], ax[a]
        return self.transpose(ax)

    def squeeze(self, axis=None) -> "DeviceArray":
        if axis is None:
            keep = [i for i, d in enumerate(self.shape) if d != 1]
        else:
            axes = (axis,) if isinstance(axis, int) else tuple(axis)
            axes = tuple(a % self.ndim for a in axes)
            for a in axes:
                if self.shape[a] != 1:
                    raise ValueError("cannot select an axis to squeeze out which has size not equal to one")
            keep = [i for i in range(self.ndim) if i not in axes]
        return self._view([self.shape[i] for i in keep], [self.strides[i] for i in keep])

    def expand_dims(self, axis) -> "DeviceArray":
        axes = (axis,) if isinstance(axis, int) else tuple(axis)
        nd = self.ndim + len(axes)
        axes = sorted(a % nd for a in axes)
        shape, strides = list(self.shape), list(self.strides)
        for a in axes:
            shape.insert(a, 1)
            strides.insert(a, 0)
        return self._view(shape, strides)

    def broadcast_to(self, shape) -> "DeviceArray":
        shape = _int_tuple(shape)
        if shape == self.shape:
            return self
        if len(shape) < self.ndim:
            raise ValueError("input operand has more dimensions than allowed by the axis remapping")
        lead = len(shape) - self.ndim
        strides = [0] * lead
        for d, s, t in zip(self.shape, self.strides, shape[lead:]):
            if d == t:
                strides.append(s)
            elif d == 1:
                strides.append(0)
            else:
                raise ValueError(f"operands could not be broadcast together with remapped shapes {self.shape} -> {shape}")
        return self._view(shape, strides)

    # ------------------------------------------------------------------ indexing
    def _basic_index(self, index):
        """(ptr, shape, strides) of a basic (int / slice / None / Ellipsis) index, else None."""
        if not isinstance(index, tuple):
            index = (index,)
        if any(not (isinstance(i, (int, np.integer, slice)) or i is None or i is Ellipsis) for i in index):
            return None
        n_real = sum(1 for i in index if i is not None and i is not Ellipsis)
        if n_real > self.ndim:
            raise IndexError("too many indices for array")
        if Ellipsis in index:
            k = index.index(Ellipsis)
            index = index[:k] + (slice(None),) * (self.ndim - n_real) + index[k + 1:]
        else:
            index = index + (slice(None),) * (self.ndim - n_real)
        off, shape, strides, dim = 0, [], [], 0
        for i in index:
            if i is None:
                shape.append(1)
                strides.append(0)
                continue
            d, s = self.shape[dim], self.strides[dim]
            if isinstance(i, slice):
                start, stop, step = i.indices(d)
                n = max(0, (stop - start + (step - (1 if step > 0 else -1))) // step)
                off += start * s
                shape.append(n)
                strides.append(s * step)
            else:
                j = int(i)
                if j < -d or j >= d:
                    raise IndexError(f"index {j} is out of bounds for axis {dim} with size {d}")
                off += (j % d) * s
            dim += 1
        return self.ptr + off * self.itemsize, shape, strides

    @staticmethod
    def _index_array(index) -> Optional[np.ndarray]:
        if isinstance(index, DeviceArray):
            index = index.get()
        if isinstance(index, (list, np.ndarray)):
            arr = np.asarray(index)
            if arr.dtype.kind in "iu" and arr.ndim == 1:
                return arr.astype(np.int64)
        return None

    def __getitem__(self, index) -> "DeviceArray":
        basic = self._basic_index(index)
        if basic is not None:
            ptr, shape, strides = basic
            return self._view(shape, strides, ptr)
        idx = self._index_array(index)
        if idx is None or self.ndim < 1:
            raise NotImplementedError("device indexing supports basic slices and a 1-D integer array on axis 0")
        n_rows = self.shape[0]
        if ((idx < -n_rows) | (idx >= n_rows)).any():
            raise IndexError("index out of bounds")
        self._require_f32("integer-array indexing")
        src = self.dense()
        row = _numel(self.shape[1:])
        out = DeviceArray.empty((len(idx),) + self.shape[1:])
        didx = DeviceArray.from_host(idx, I64)
        _capi.check(_capi.lib().mt_gather_rows_f32(out.ptr, src.ptr, didx.ptr, len(idx), row, n_rows), "mt_gather_rows_f32")
        return out

    def __setitem__(self, index, value) -> None:
        """Assignment (full_grad[index] = grad, reference overload.py:36-38): duplicates ASSIGN."""
        self._require_f32("__setitem__")
        basic = self._basic_index(index)
        if basic is not None:
            ptr, shape, strides = basic
            value = _as_device(value)
            src = value.broadcast_to(shape).dense() if value.shape != tuple(shape) else value.dense()
            _capi.check(_capi.lib().mt_scatter_strided_f32(ptr, _capi.i64(strides), src.ptr, len(shape),
                                                           _capi.i64(shape)), "mt_scatter_strided_f32")
            self._host_value = None
            return
        idx = self._index_array(index)
        if idx is None or not self.is_dense:
            raise NotImplementedError("device assignment supports basic slices and a 1-D integer array on axis 0")
        row = _numel(self.shape[1:])
        value = _as_device(value)
        want = (len(idx),) + self.shape[1:]
        src = value.broadcast_to(want).dense() if value.shape != want else value.dense()
        didx = DeviceArray.from_host(idx, I64)
        _capi.check(_capi.lib().mt_scatter_rows_assign_f32(self.ptr, src.ptr, didx.ptr, len(idx), row, self.shape[0]),
                    "mt_scatter_rows_assign_f32")

    # ------------------------------------------------------------------ arithmetic (delegates to kernels.py)
    def __add__(self, o): from . import kernels as K; return K.binary(_capi.BIN_ADD, self, o)
    def __radd__(self, o): from . import kernels as K; return K.binary(_capi.BIN_ADD, o, self)
    def __sub__(self, o): from . import kernels as K; return K.binary(_capi.BIN_SUB, self, o)
    def __rsub__(self, o): from . import kernels as K; return K.binary(_capi.BIN_SUB, o, self)
    def __mul__(self, o): from . import kernels as K; return K.binary(_capi.BIN_MUL, self, o)
    def __rmul__(self, o): from . import kernels as K; return K.binary(_capi.BIN_MUL, o, self)
    def __truediv__(self, o): from . import kernels as K; return K.binary(_capi.BIN_DIV, self, o)
    def __rtruediv__(self, o): from . import kernels as K; return K.binary(_capi.BIN_DIV, o, self)
    def __neg__(self): from . import kernels as K; return K.unary(_capi.UN_NEG, self)
    def __pow__(self, p): from . import kernels as K; return K.unary(_capi.UN_POW, self, float(p))
    def __matmul__(self, o): from . import kernels as K; return K.matmul(self, o)
    def __rmatmul__(self, o): from . import kernels as K; return K.matmul(o, self)

    def sum(self, axis=None, keepdims: bool = False):
        from . import kernels as K
        return K.sum(self, axis, keepdims)

    # comparisons: float32 masks (1.0 / 0.0) on the device
    def __lt__(self, o): from . import kernels as K; return K.binary(_capi.BIN_LT, self, o)
    def __gt__(self, o): from . import kernels as K; return K.binary(_capi.BIN_GT, self, o)
    def __le__(self, o): from . import kernels as K; return K.binary(_capi.BIN_LE, self, o)
    def __ge__(self, o): from . import kernels as K; return K.binary(_capi.BIN_GE, self, o)
    def __eq__(self, o): from . import kernels as K; return K.binary(_capi.BIN_EQ, self, o)
    def __ne__(self, o): from . import kernels as K; return K.binary(_capi.BIN_NE, self, o)
    __hash__ = object.__hash__


_CONSTS: dict = {}


def device_constant(value: float) -> "DeviceArray":
    """A shared, read-only 0-d device array for a Python scalar operand (`x * 0.49`, `/ count`, the backward seed).
    Uploading a scalar is a blocking 4-byte H2D copy; operands coerced inside ops are never written, so they
    are uploaded once per distinct value."""
    key = float(np.float32(value))
    arr = _CONSTS.get(key)
    if arr is None:
        if len(_CONSTS) > 4096:
            _CONSTS.clear()
        arr = DeviceArray.from_host(key)
        arr._const = True
        _CONSTS[key] = arr
    return arr


_ONE = None


def _one() -> DeviceArray:
    global _ONE
    if _ONE is None:
        _ONE = DeviceArray.from_host(1.0)
    return _ONE


def _as_device(x) -> DeviceArray:
    if isinstance(x, DeviceArray):
        return x
    if isinstance(x, np.ndarray) and x.dtype != np.float32:
        x = x.astype(np.float32)
    return DeviceArray.from_host(x)
