"""Tiny helpers to drive the C ABI directly from tests/tools with NumPy host buffers."""
import ctypes as C

import numpy as np

from microtorch_b200 import _capi


class Dev:
    """A device buffer owned by the test (mt_malloc / mt_free)."""

    def __init__(self, host=None, shape=None):
        self.lib = _capi.lib()
        if host is not None:
            host = np.ascontiguousarray(host, dtype=np.float32 if host.dtype.kind == "f" else host.dtype)
            self.shape, self.dtype = host.shape, host.dtype
        else:
            self.shape, self.dtype = tuple(shape), np.dtype(np.float32)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        p = C.c_void_p()
        _capi.check(self.lib.mt_malloc(C.byref(p), max(self.nbytes, 1)), "mt_malloc")
        self.ptr = p.value
        if host is not None and self.nbytes:
            _capi.check(self.lib.mt_h2d(self.ptr, host.ctypes.data, self.nbytes), "mt_h2d")

    def get(self):
        out = np.empty(self.shape, self.dtype)
        if self.nbytes:
            _capi.check(self.lib.mt_d2h(out.ctypes.data, self.ptr, self.nbytes), "mt_d2h")
        return out

    def free(self):
        if self.ptr:
            self.lib.mt_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def strides_of(shape, bshape):
    """Element strides of a contiguous array of `shape` broadcast to `bshape` (0 where broadcast)."""
    shape = (1,) * (len(bshape) - len(shape)) + tuple(shape)
    st, acc = [], 1
    for d in reversed(shape):
        st.append(acc)
        acc *= d
    st = st[::-1]
    return [0 if shape[i] == 1 and bshape[i] != 1 else st[i] for i in range(len(bshape))]


def timed(fn, iters=10, warmup=3, flush=True):
    """Average ms per call of fn() measured with CUDA events on the library's stream."""
    lib = _capi.lib()
    for _ in range(warmup):
        fn()
    e0, e1 = C.c_void_p(), C.c_void_p()
    lib.mt_event_create(C.byref(e0))
    lib.mt_event_create(C.byref(e1))
    total = 0.0
    for _ in range(iters):
        if flush:
            lib.mt_l2_flush()
        lib.mt_event_record(e0)
        fn()
        lib.mt_event_record(e1)
        lib.mt_event_sync(e1)
        ms = C.c_float()
        lib.mt_event_elapsed_ms(e0, e1, C.byref(ms))
        total += ms.value
    lib.mt_event_destroy(e0)
    lib.mt_event_destroy(e1)
    return total / iters
