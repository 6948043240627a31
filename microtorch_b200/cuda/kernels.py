"""Kernel wrappers of the B200 backend: DeviceArray in, DeviceArray out, every byte moved by
libmicrotorch_b200.so through the C ABI.  Host-side logic here is shape bookkeeping only
(broadcast strides, the canonical [outer, red, inner] plan of a reduction)."""

from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .. import _capi
from .array import DeviceArray, _as_device, _dense_strides, _numel


def _lib():
    return _capi.lib()


# ----------------------------------------------------------------------------- elementwise
def broadcast_shape(a: Sequence[int], b: Sequence[int]) -> Tuple[int, ...]:
    if a == b:
        return a if type(a) is tuple else tuple(a)
    if len(b) == 0:
        return a if type(a) is tuple else tuple(a)
    if len(a) == 0:
        return b if type(b) is tuple else tuple(b)
    return tuple(np.broadcast_shapes(tuple(a), tuple(b)))


def binary(op: int, a, b, scalar: float = 0.0, out: Optional[DeviceArray] = None) -> DeviceArray:
    """out = op(a, b) with NumPy broadcasting; operands may be any strided views."""
    a, b = _as_device(a), _as_device(b)
    shape = broadcast_shape(a.shape, b.shape)
    av, bv = a.broadcast_to(shape), b.broadcast_to(shape)
    if out is None:
        out = DeviceArray.empty(shape)
    else:
        if out.shape != shape or not out.is_dense:
            raise ValueError("binary(out=): output must be a dense array of the broadcast shape")
        if out._const:
            raise ValueError("binary(out=): the output is a shared read-only constant")
        out._host_value = None
    if out.size:
        status = _lib().mt_ew_binary_f32(op, out.ptr, av.ptr, bv.ptr, len(shape), _capi.i64(shape), _capi.i64(av.strides),
                                         _capi.i64(bv.strides), float(scalar))
        if status:
            _capi.check(status, "mt_ew_binary_f32")
    return out


def where(cond, a, b) -> DeviceArray:
    """cond != 0 ? a : b with NumPy broadcasting (Python scalars allowed for a / b)."""
    cond, a, b = _as_device(cond), _as_device(a), _as_device(b)
    shape = broadcast_shape(broadcast_shape(cond.shape, a.shape), b.shape)
    cv, av, bv = cond.broadcast_to(shape), a.broadcast_to(shape), b.broadcast_to(shape)
    out = DeviceArray.empty(shape)
    if out.size:
        _capi.check(_lib().mt_ew_where_f32(out.ptr, cv.ptr, av.ptr, bv.ptr, len(shape), _capi.i64(shape),
                                           _capi.i64(cv.strides), _capi.i64(av.strides), _capi.i64(bv.strides)),
                    "mt_ew_where_f32")
    return out


def inplace(op: int, dst: DeviceArray, src) -> DeviceArray:
    """dst (op)= src, broadcasting src; dst may be a strided view (np.add(a, b, out=a))."""
    src = _as_device(src)
    if dst._const:
        raise ValueError("in-place update of a shared read-only constant")
    if broadcast_shape(dst.shape, src.shape) != dst.shape:
        raise ValueError(f"non-broadcastable output operand with shape {dst.shape}")
    sv = src.broadcast_to(dst.shape)
    if dst.size:
        _capi.check(_lib().mt_ew_inplace_f32(op, dst.ptr, _capi.i64(dst.strides), sv.ptr, _capi.i64(sv.strides),
                                             dst.ndim, _capi.i64(dst.shape)), "mt_ew_inplace_f32")
    dst._host_value = None
    return dst


def _unbroadcast_view(x: DeviceArray) -> DeviceArray:
    """The distinct elements of a broadcast (stride-0) view: extent 1 wherever the stride is 0."""
    idx = tuple(slice(0, 1) if (s == 0 and d > 1) else slice(None) for d, s in zip(x.shape, x.strides))
    return x[idx]


def unary(op: int, x: DeviceArray, scalar: float = 0.0) -> DeviceArray:
    x = _as_device(x)
    if x.size and any(s == 0 and d > 1 for d, s in zip(x.shape, x.strides)):
        # f(broadcast(v)) == broadcast(f(v)): touch only the distinct elements (e.g. -g for g = sum().backward())
        return unary(op, _unbroadcast_view(x), scalar).broadcast_to(x.shape)
    src = x.dense()
    out = DeviceArray.empty(x.shape)
    if out.size:
        _capi.check(_lib().mt_ew_unary_f32(op, out.ptr, src.ptr, out.size, float(scalar)), "mt_ew_unary_f32")
    return out


# ----------------------------------------------------------------------------- reductions
def plan_reduction(shape: Sequence[int], reduce_axes: Sequence[int]) -> List[Tuple[int, int, int]]:
    """Decompose "sum over `reduce_axes` of a dense array of `shape`" into passes of the
    [outer, red, inner] primitive.  Adjacent axes with the same role are merged; each pass
    removes the LAST reduced group, so every pass reads a dense array.  Pure host logic.

    Returns [(outer, red, inner), ...]; the array after pass i has shape outer x inner
    (flattened) and feeds pass i+1.  Extent-1 axes never matter."""
    red = {int(a) % max(len(shape), 1) for a in reduce_axes}
    groups: List[List[int]] = []           # [is_reduced, extent]
    for i, d in enumerate(shape):
        if d == 1:
            continue
        r = int(i in red)
        if groups and groups[-1][0] == r:
            groups[-1][1] *= d
        else:
            groups.append([r, int(d)])
    passes = []
    while any(g[0] for g in groups):
        k = max(i for i, g in enumerate(groups) if g[0])
        outer = _numel([g[1] for g in groups[:k]])
        inner = _numel([g[1] for g in groups[k + 1:]])
        passes.append((outer, groups[k][1], inner))
        del groups[k]
        merged: List[List[int]] = []
        for g in groups:                   # neighbours of the removed group may now merge
            if merged and merged[-1][0] == g[0]:
                merged[-1][1] *= g[1]
            else:
                merged.append(list(g))
        groups = merged
    return passes


def _other_strides_3d(other: DeviceArray, shape: Sequence[int], reduce_axes: Sequence[int],
                      plan: Tuple[int, int, int]) -> Optional[Tuple[int, int, int]]:
    """Strides (so, sr, si) of `other` (already broadcast to `shape`) over the single-pass
    [outer, red, inner] view, or None when its layout cannot be expressed that way."""
    nd = len(shape)
    red = sorted(int(a) % nd for a in reduce_axes)
    live = [i for i in range(nd) if shape[i] != 1]
    red_live = [i for i in live if i in red]
    if not red_live:
        return None
    lo, hi = red_live[0], red_live[-1]
    if any(i not in red for i in live if lo <= i <= hi):
        return None                         # reduced axes are not one contiguous group
    groups = ([i for i in live if i < lo], red_live, [i for i in live if i > hi])
    out = []
    for axes in groups:
        if not axes:
            out.append(0)
            continue
        # the group must be collapsible for `other`: stride[i] == stride[i+1]*shape[i+1]
        for x, y in zip(axes[:-1], axes[1:]):
            if other.strides[x] != other.strides[y] * shape[y]:
                return None
        out.append(other.strides[axes[-1]])
    return tuple(out)


def reduce_sum(x: DeviceArray, reduce_axes: Sequence[int], other: Optional[DeviceArray] = None) -> DeviceArray:
    """sum(x * other, axis=reduce_axes) with the reduced axes REMOVED (flat result of the kept
    extents, caller reshapes).  `other` (optional, broadcastable to x) is fused into the first
    pass when its layout allows, otherwise multiplied first."""
    x = _as_device(x)
    shape = x.shape
    red_set = {int(a) % max(len(shape), 1) for a in reduce_axes}
    kept = [d for i, d in enumerate(shape) if i not in red_set]
    if other is None and x.size and all(x.strides[a] == 0 or shape[a] == 1 for a in red_set) \
            and any(shape[a] > 1 for a in red_set):
        # the view is constant along every reduced axis (a broadcast gradient): sum == value * count
        count = _numel([shape[a] for a in red_set])
        idx = tuple(slice(0, 1) if i in red_set else slice(None) for i in range(len(shape)))
        return unary(_capi.UN_SCALE, x[idx], float(count)).dense().reshape(kept)
    if other is not None:
        other = _as_device(other).broadcast_to(shape)
        bcast = lambda a: all(st == 0 or d == 1 for d, st in zip(a.shape, a.strides))
        if x.size and bcast(other):         # sum(x * c) == c * sum(x): c is one broadcast scalar
            c = other[(0,) * len(shape)]
            return binary(_capi.BIN_MUL, reduce_sum(x, reduce_axes), c)
        if x.size and bcast(x):             # sum(c * other) == c * sum(other)  (c = sum().backward()'s gradient)
            c = x[(0,) * len(shape)]
            return binary(_capi.BIN_MUL, reduce_sum(other, reduce_axes), c)
        if not x.is_dense and other.is_dense:
            x, other = other, x             # the kernel streams its first operand densely, the second may be strided
    src = x.dense()
    passes = plan_reduction(shape, reduce_axes)
    if not passes:                          # nothing to add up (all reduced extents are 1)
        res = src if other is None else binary(_capi.BIN_MUL, src, other)
        return res.reshape(kept)
    fused = None
    if other is not None:
        if len(passes) == 1:
            fused = _other_strides_3d(other, shape, reduce_axes, passes[0])
        if fused is None:
            src = binary(_capi.BIN_MUL, src, other)
            other = None
    cur = src
    for i, (outer, red, inner) in enumerate(passes):
        out = DeviceArray.empty((outer * inner,))
        if i == 0 and fused is not None:
            so, sr, si = fused
            optr = other.ptr
        else:
            so = sr = si = 0
            optr = None
        _capi.check(_lib().mt_reduce_sum_f32(out.ptr, cur.ptr, outer, red, inner, optr, so, sr, si), "mt_reduce_sum_f32")
        cur = out
    return cur.reshape(kept)


def sum(x: DeviceArray, axis=None, keepdims: bool = False) -> DeviceArray:      # noqa: A001 (NumPy name)
    x = _as_device(x)
    nd = x.ndim
    if axis is None:
        axes = tuple(range(nd))
    elif isinstance(axis, (tuple, list)):
        axes = tuple(int(a) % nd for a in axis)
    else:
        axes = (int(axis) % nd,)
    res = reduce_sum(x, axes)
    if keepdims:
        return res.reshape([1 if i in axes else d for i, d in enumerate(x.shape)])
    return res


def extreme(x: DeviceArray, axis=None, keepdims: bool = False, is_max: bool = True) -> DeviceArray:
    """np.max / np.min(x, axis, keepdims) (Reduce.max / Reduce.min forward, reference reduce.py:66,87)."""
    x = _as_device(x)
    nd = x.ndim
    if axis is None:
        axes = tuple(range(nd))
    elif isinstance(axis, (tuple, list)):
        axes = tuple(int(a) % nd for a in axis)
    else:
        axes = (int(axis) % nd,)
    if x.size == 0:
        raise ValueError("zero-size array to reduction operation which has no identity")
    kept = [d for i, d in enumerate(x.shape) if i not in axes]
    cur = x.dense()
    for outer, red, inner in plan_reduction(x.shape, axes):
        out = DeviceArray.empty((outer * inner,))
        _capi.check(_lib().mt_reduce_extreme_f32(1 if is_max else 0, out.ptr, cur.ptr, outer, red, inner),
                    "mt_reduce_extreme_f32")
        cur = out
    res = cur.reshape(kept) if cur is not x else cur.copy().reshape(kept)
    if keepdims:
        return res.reshape([1 if i in axes else d for i, d in enumerate(x.shape)])
    return res


def unbroadcast(grad: DeviceArray, shape: Sequence[int], other: Optional[DeviceArray] = None) -> DeviceArray:
    """BaseOps.bkwd_broadcast (reference base.py:8-58) with an optional fused `grad * other`
    (mul backward, overload.py:109-112): fold `grad` onto an operand of `shape`."""
    shape = tuple(shape)
    if len(shape) == 0:
        if grad.ndim == 0 and other is None:
            return grad
        return reduce_sum(grad, tuple(range(grad.ndim)), other).reshape(())
    if grad.ndim == 0:
        return grad if other is None else binary(_capi.BIN_MUL, grad, other)
    lead = max(0, grad.ndim - len(shape))
    axes = list(range(lead)) + [lead + d for d in range(len(shape)) if shape[d] == 1 and grad.shape[lead + d] > 1]
    if not axes:
        res = grad if other is None else binary(_capi.BIN_MUL, grad, other)
    else:
        res = reduce_sum(grad, axes, other)
    return res if res.shape == shape else res.reshape(shape)


# ----------------------------------------------------------------------------- GEMM
def matmul(a, b) -> DeviceArray:
    """a @ b with NumPy semantics for 1-D/2-D/batched operands, any strides (no copies)."""
    a, b = _as_device(a), _as_device(b)
    a1, b1 = a.ndim == 1, b.ndim == 1
    if a.ndim == 0 or b.ndim == 0:
        raise ValueError("matmul: Input operand does not have enough dimensions")
    av = a.expand_dims(0) if a1 else a
    bv = b.expand_dims(1) if b1 else b
    M, K = av.shape[-2], av.shape[-1]
    K2, N = bv.shape[-2], bv.shape[-1]
    if K != K2:
        raise ValueError(f"matmul: Input operand 1 has a mismatch in its core dimension 0 (size {K2} is different from {K})")
    batch_shape = broadcast_shape(av.shape[:-2], bv.shape[:-2])
    if len(batch_shape) > 1:
        av = av.broadcast_to(batch_shape + (M, K)).dense().reshape((-1, M, K)) if av.ndim > 2 else av
        bv = bv.broadcast_to(batch_shape + (K, N)).dense().reshape((-1, K, N)) if bv.ndim > 2 else bv
    nb = _numel(batch_shape)
    a_sb = av.strides[0] if av.ndim == 3 and av.shape[0] != 1 else 0
    b_sb = bv.strides[0] if bv.ndim == 3 and bv.shape[0] != 1 else 0
    out = DeviceArray.empty(batch_shape + (M, N))
    if out.size:
        if K == 0:
            out.fill(0.0)
        else:
            _capi.check(_lib().mt_matmul_f32(out.ptr, av.ptr, bv.ptr, nb, M, N, K, a_sb, av.strides[-2], av.strides[-1],
                                             b_sb, bv.strides[-2], bv.strides[-1]), "mt_matmul_f32")
    if a1 and b1:
        return out.reshape(batch_shape)
    if a1:
        return out.reshape(batch_shape + (N,))
    if b1:
        return out.reshape(batch_shape + (M,))
    return out


def linear_forward(x2d: DeviceArray, weight: DeviceArray, bias: Optional[DeviceArray], act: int) -> DeviceArray:
    """Y = act(X W^T + b): one fused launch (mt_linear_fwd_f32)."""
    x2d, weight = x2d.dense(), weight.dense()
    M, K = x2d.shape
    N = weight.shape[0]
    out = DeviceArray.empty((M, N))
    b = bias.dense() if bias is not None else None
    _capi.check(_lib().mt_linear_fwd_f32(out.ptr, x2d.ptr, weight.ptr, b.ptr if b is not None else None, M, N, K, act),
                "mt_linear_fwd_f32")
    return out


LINEAR_ACCUMULATE_DW = 1
LINEAR_DX_TIMES_TANH_GRAD_OF_X = 2


def linear_backward(grad_out: DeviceArray, y_act: Optional[DeviceArray], x2d: DeviceArray, weight: DeviceArray,
                    need_dx: bool, dx_preact: bool = False) -> Tuple[DeviceArray, Optional[DeviceArray]]:
    """(dW, dX) of the fused Linear(+Tanh) node.  `y_act` given: grad_out is dL/dY and tanh' is applied first
    (GEMM prologue or one explicit pass, the library decides); None: grad_out already is dL/dZ.
    `dx_preact`: x2d is itself a tanh output, so dX is returned multiplied by (1 - x^2) in the GEMM epilogue -
    i.e. as the pre-activation gradient of the layer that produced x."""
    g = grad_out.dense()
    x2d, weight = x2d.dense(), weight.dense()
    M, K = x2d.shape
    N = weight.shape[0]
    dW = DeviceArray.empty((N, K))
    dX = DeviceArray.empty((M, K)) if need_dx else None
    ya = y_act.dense() if y_act is not None else None
    _capi.check(_lib().mt_linear_bwd_f32(dX.ptr if dX is not None else None, dW.ptr, g.ptr,
                                         ya.ptr if ya is not None else None, x2d.ptr, weight.ptr, M, N, K,
                                         LINEAR_DX_TIMES_TANH_GRAD_OF_X if (dx_preact and need_dx) else 0),
                "mt_linear_bwd_f32")
    return dW, dX


# ----------------------------------------------------------------------------- loss / SGD
def loss_forward_backward(kind: int, pred: DeviceArray, target: DeviceArray, count: int,
                          want_grad: bool) -> Tuple[DeviceArray, Optional[DeviceArray]]:
    p, t = pred.dense(), target.dense()
    loss = DeviceArray.empty(())
    dpred = DeviceArray.empty(pred.shape) if want_grad else None
    scale = float(np.float32(count) ** -1)              # Tensor(count) ** -1 in float32 (reduce.py:41-43)
    _capi.check(_lib().mt_loss_fwd_bwd_f32(kind, loss.ptr, dpred.ptr if dpred is not None else None, p.ptr, t.ptr,
                                           p.size, scale), "mt_loss_fwd_bwd_f32")
    return loss, dpred


def sgd_multi(params: Sequence[DeviceArray], grads: Sequence[DeviceArray], lr: float, zero_grads: bool = False) -> None:
    n = len(params)
    if not n:
        return
    for p, g in zip(params, grads):
        if not (p.is_dense and g.is_dense and p.size == g.size):
            raise ValueError("sgd_multi: parameters and gradients must be dense and of equal size")
    pa = (C.c_void_p * n)(*[p.ptr for p in params])
    ga = (C.c_void_p * n)(*[g.ptr for g in grads])
    sa = (C.c_size_t * n)(*[p.size for p in params])
    _capi.check(_lib().mt_sgd_multi_f32(pa, ga, sa, n, float(lr), 1 if zero_grads else 0), "mt_sgd_multi_f32")
    for p in params:
        p._host_value = None


def zero_multi(arrays: Sequence[DeviceArray]) -> None:
    arrays = [a for a in arrays if a.size]
    n = len(arrays)
    if not n:
        return
    pa = (C.c_void_p * n)(*[a.ptr for a in arrays])
    ba = (C.c_size_t * n)(*[a.nbytes for a in arrays])
    _capi.check(_lib().mt_memset_multi(pa, ba, n), "mt_memset_multi")


def synchronize() -> None:
    _capi.check(_lib().mt_sync(), "mt_sync")
