"""Kernel-level parity through the raw C ABI (no Tensor layer): every entry point against
NumPy on the same seeded inputs.  Needs a B200: run with `-m gpu` under gpurun.

Tolerances (north_star): shape/index work bit-exact; elementwise + reductions 1e-5 relative;
GEMM paths 1e-4 relative (norm-wise: max|err| / max|ref|)."""
import ctypes as C

import numpy as np
import pytest

from microtorch_b200 import _capi
from capi_util import Dev, strides_of

pytestmark = pytest.mark.gpu

RT = dict(rtol=1e-5, atol=1e-7)


@pytest.fixture(scope="module")
def lib():
    _capi.require_device()
    return _capi.lib()


def binary(lib, op, a, b, scalar=0.0):
    bshape = np.broadcast_shapes(a.shape, b.shape)
    da, db, out = Dev(a), Dev(b), Dev(shape=bshape)
    _capi.check(lib.mt_ew_binary_f32(op, out.ptr, da.ptr, db.ptr, len(bshape), _capi.i64(bshape),
                                     _capi.i64(strides_of(a.shape, bshape)), _capi.i64(strides_of(b.shape, bshape)),
                                     scalar))
    return out.get()


SHAPES = [((6, 8), (6, 8)), ((64, 4096), (1, 4096)), ((64, 4096), (64, 1)), ((37, 129), ()), ((3, 6, 8), (8,)),
          ((3, 6, 8), (1, 6, 1)), ((1 << 20,), (1 << 20,)), ((1000003,), (1000003,)), ((5, 1, 7), (1, 4, 1)),
          ((0, 4), (1, 4)), ((129, 260), (129, 1)), ((1, 260), (129, 1))]


@pytest.mark.parametrize("sa,sb", SHAPES)
def test_binary_broadcast_bit_exact(lib, sa, sb):
    rng = np.random.default_rng(0)
    a = rng.uniform(0.5, 1.5, sa).astype(np.float32)
    b = rng.uniform(0.5, 1.5, sb).astype(np.float32)
    np.testing.assert_array_equal(binary(lib, _capi.BIN_ADD, a, b), a + b)
    np.testing.assert_array_equal(binary(lib, _capi.BIN_MUL, a, b), a * b)
    np.testing.assert_array_equal(binary(lib, _capi.BIN_SUB, a, b), a - b)
    np.testing.assert_array_equal(binary(lib, _capi.BIN_DIV, a, b), a / b)
    np.testing.assert_array_equal(binary(lib, _capi.BIN_ADD, b, a), b + a)      # broadcast on the left


def test_fused_backward_ops(lib):
    rng = np.random.default_rng(1)
    g = rng.standard_normal((129, 257)).astype(np.float32)
    x = rng.uniform(0.5, 1.5, (129, 257)).astype(np.float32)
    y = np.tanh(x)
    np.testing.assert_array_equal(binary(lib, _capi.BIN_TANH_BWD, g, y), g * (1 - y ** 2))
    for p in (2.0, 3.0, -1.0, 0.5, 2.5):
        want = g * (np.float32(p) * x ** np.float32(p - 1))
        np.testing.assert_allclose(binary(lib, _capi.BIN_POW_BWD, g, x, p), want, **RT)
    # a 0-d upstream gradient broadcast over the saved tensor (sum().backward())
    g0 = np.float32(0.25).reshape(())
    np.testing.assert_array_equal(binary(lib, _capi.BIN_TANH_BWD, g0, y), g0 * (1 - y ** 2))


@pytest.mark.parametrize("n", [1, 3, 1024, 1 << 20, (1 << 20) + 7])
def test_unary(lib, n):
    rng = np.random.default_rng(2)
    x = rng.uniform(0.5, 1.5, n).astype(np.float32)
    dx, out = Dev(x), Dev(shape=(n,))
    cases = [(_capi.UN_NEG, 0, -x, True), (_capi.UN_LOG, 0, np.log(x), False), (_capi.UN_TANH, 0, np.tanh(x), False),
             (_capi.UN_EXP, 0, np.exp(x), False), (_capi.UN_POW, 2.0, x ** 2, True), (_capi.UN_POW, -1.0, x ** -1, True),
             (_capi.UN_POW, 0.5, x ** 0.5, True), (_capi.UN_POW, 3.0, x ** 3, False), (_capi.UN_POW, 1.0, x, True),
             (_capi.UN_POW, 0.0, np.ones_like(x), True), (_capi.UN_SCALE, 0.37, x * np.float32(0.37), True),
             (_capi.UN_ADDS, 0.5, x + np.float32(0.5), True)]
    for op, s, want, exact in cases:
        _capi.check(lib.mt_ew_unary_f32(op, out.ptr, dx.ptr, n, s))
        got = out.get()
        if exact:
            np.testing.assert_array_equal(got, want.astype(np.float32))
        else:
            np.testing.assert_allclose(got, want, **RT)


def reduce(lib, x, outer, red, inner, other=None, ostr=(0, 0, 0)):
    dx, out = Dev(x), Dev(shape=(outer, inner))
    do = Dev(other) if other is not None else None
    _capi.check(lib.mt_reduce_sum_f32(out.ptr, dx.ptr, outer, red, inner, do.ptr if do else None, *ostr))
    return out.get()


@pytest.mark.parametrize("outer,red,inner", [(1, 1 << 22, 1), (1, 1000003, 1), (1024, 4096, 1), (5000, 10, 1),
                                             (7, 33, 1), (1, 8192, 4096), (1, 131, 2048), (3, 1000, 36), (256, 3, 2),
                                             (1, 1, 5), (1, 70000, 12)])
def test_reduce_sum(lib, outer, red, inner):
    rng = np.random.default_rng(3)
    x = rng.uniform(0.5, 1.5, (outer, red, inner)).astype(np.float32)
    got = reduce(lib, x, outer, red, inner)
    want = x.astype(np.float64).sum(axis=1)
    np.testing.assert_allclose(got, want, rtol=1e-5)
    np.testing.assert_allclose(got, x.sum(axis=1), rtol=1e-5)          # and against NumPy's fp32 pairwise sum


@pytest.mark.parametrize("outer,red,inner", [(1, 1 << 22, 1), (1, 1000003, 1), (1024, 4096, 1), (5000, 10, 1), (7, 33, 1),
                                             (1, 8192, 4096), (1, 131, 2048), (3, 1000, 36), (256, 3, 2), (1, 1, 5),
                                             (1, 70000, 12), (2, 5, 3)])
def test_reduce_extreme_bit_exact(lib, outer, red, inner):
    """Reduce.max / Reduce.min forward: selection, so bit-exact against np.max / np.min, NaN included."""
    rng = np.random.default_rng(31)
    x = rng.standard_normal((outer, red, inner)).astype(np.float32)
    dx, out = Dev(x), Dev(shape=(outer, inner))
    for is_max, fn in ((1, np.max), (0, np.min)):
        _capi.check(lib.mt_reduce_extreme_f32(is_max, out.ptr, dx.ptr, outer, red, inner))
        np.testing.assert_array_equal(out.get(), fn(x, axis=1))
    x[0, red // 2, 0] = np.nan                                          # NaN propagates like NumPy's max / min
    dx = Dev(x)
    for is_max, fn in ((1, np.max), (0, np.min)):
        _capi.check(lib.mt_reduce_extreme_f32(is_max, out.ptr, dx.ptr, outer, red, inner))
        np.testing.assert_array_equal(out.get(), fn(x, axis=1))


def test_reduce_sum_fused_multiply(lib):
    """unbroadcast(grad * other): the mul-backward pattern, column, row and full reductions."""
    rng = np.random.default_rng(4)
    g = rng.standard_normal((512, 1024)).astype(np.float32)
    x = rng.uniform(0.5, 1.5, (512, 1024)).astype(np.float32)
    want = (g.astype(np.float64) * x).sum(axis=0)
    got = reduce(lib, g, 1, 512, 1024, x, (0, 1024, 1))
    np.testing.assert_allclose(got[0], want, rtol=1e-5, atol=1e-4)
    got = reduce(lib, g, 512, 1024, 1, x, (1024, 1, 0))
    np.testing.assert_allclose(got[:, 0], (g.astype(np.float64) * x).sum(axis=1), rtol=1e-5, atol=1e-4)
    row = rng.uniform(0.5, 1.5, (1024,)).astype(np.float32)            # other broadcast along the reduced axis
    got = reduce(lib, g, 1, 512, 1024, row, (0, 0, 1))
    np.testing.assert_allclose(got[0], (g.astype(np.float64) * row).sum(axis=0), rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("kind,name", [(0, "mse"), (1, "ce"), (2, "bce")])
@pytest.mark.parametrize("n", [160, 65536 * 10, 1000003])
def test_loss_fwd_bwd(lib, kind, name, n):
    rng = np.random.default_rng(5)
    p = rng.uniform(0.05, 0.95, n).astype(np.float32)
    t = (rng.uniform(0, 1, n) > 0.5).astype(np.float32)
    inv = np.float32(1.0) / np.float32(n)
    p64, t64 = p.astype(np.float64), t.astype(np.float64)
    if name == "mse":
        val, grad = ((p64 - t64) ** 2).mean(), 2 * (p64 - t64) / n
    elif name == "ce":
        val, grad = (np.log(p64) - t64).mean(), 1.0 / (n * p64)
    else:
        val = (-(t64 * np.log(p64) + (1 - t64) * np.log(1 - p64))).mean()
        grad = -(t64 / p64 - (1 - t64) / (1 - p64)) / n
    dp_, dt_, loss, dpred = Dev(p), Dev(t), Dev(shape=(1,)), Dev(shape=(n,))
    _capi.check(lib.mt_loss_fwd_bwd_f32(kind, loss.ptr, dpred.ptr, dp_.ptr, dt_.ptr, n, inv))
    np.testing.assert_allclose(loss.get()[0], val, rtol=1e-5)
    np.testing.assert_allclose(dpred.get(), grad, rtol=1e-5, atol=1e-12)
    loss2 = Dev(shape=(1,))
    _capi.check(lib.mt_loss_fwd_bwd_f32(kind, loss2.ptr, None, dp_.ptr, dt_.ptr, n, inv))   # value only, re-armed ticket
    assert loss2.get()[0] == loss.get()[0]


def test_raw_ce_nan_propagates(lib):
    p = np.float32([0.5, -0.3, 0.9, 0.2])
    t = np.float32([1, 0, 0, 1])
    dp_, dt_, loss, dpred = Dev(p), Dev(t), Dev(shape=(1,)), Dev(shape=(4,))
    _capi.check(lib.mt_loss_fwd_bwd_f32(1, loss.ptr, dpred.ptr, dp_.ptr, dt_.ptr, 4, np.float32(0.25)))
    assert np.isnan(loss.get()[0])
    np.testing.assert_allclose(dpred.get(), np.float32(0.25) / p, rtol=1e-6)


def test_sgd_multi_bit_exact(lib):
    rng = np.random.default_rng(6)
    sizes = [4096 * 1024, 4096, 10 * 4096, 10, 3, 0, 1000003]
    ps = [rng.standard_normal(s).astype(np.float32) for s in sizes]
    gs = [rng.standard_normal(s).astype(np.float32) for s in sizes]
    dps, dgs = [Dev(p) for p in ps], [Dev(g) for g in gs]
    n = len(sizes)
    pa = (C.c_void_p * n)(*[d.ptr for d in dps])
    ga = (C.c_void_p * n)(*[d.ptr for d in dgs])
    sa = (C.c_size_t * n)(*sizes)
    lr = np.float32(0.01)
    _capi.check(lib.mt_sgd_multi_f32(pa, ga, sa, n, lr, 0))
    for p, g, d, dg in zip(ps, gs, dps, dgs):
        np.testing.assert_array_equal(d.get(), p + (-(lr * g)))        # optimizer.py:39 / tensor.py:569-570
        np.testing.assert_array_equal(dg.get(), g)
    _capi.check(lib.mt_sgd_multi_f32(pa, ga, sa, n, lr, 1))
    for dg in dgs:
        assert not dg.get().any()


def test_memset_multi_and_fill(lib):
    sizes = [1024, 12, 4096 * 4096, 7]
    bufs = [Dev(np.ones(s, np.float32)) for s in sizes]
    n = len(sizes)
    pa = (C.c_void_p * n)(*[b.ptr for b in bufs])
    ba = (C.c_size_t * n)(*[4 * s for s in sizes])
    _capi.check(lib.mt_memset_multi(pa, ba, n))
    for b in bufs:
        assert not b.get().any()
    _capi.check(lib.mt_fill_f32(bufs[3].ptr, 7, 2.5))
    assert (bufs[3].get() == 2.5).all()


def test_strided_copy_and_scatter_bit_exact(lib):
    rng = np.random.default_rng(7)
    x = rng.standard_normal((5, 6, 7)).astype(np.float32)
    dx = Dev(x)
    view = x.transpose(2, 0, 1)
    out = Dev(shape=view.shape)
    st = [s // 4 for s in view.strides]
    _capi.check(lib.mt_copy_strided_f32(out.ptr, dx.ptr, 3, _capi.i64(view.shape), _capi.i64(st)))
    np.testing.assert_array_equal(out.get(), np.ascontiguousarray(view))
    back = Dev(np.zeros((5, 6, 7), np.float32))
    _capi.check(lib.mt_scatter_strided_f32(back.ptr, _capi.i64(st), out.ptr, 3, _capi.i64(view.shape)))
    np.testing.assert_array_equal(back.get(), x)


def test_gather_scatter_rows_bit_exact(lib):
    rng = np.random.default_rng(8)
    x = rng.standard_normal((50, 33)).astype(np.float32)
    idx = np.array([0, 7, 7, 49, -1, 3, 7], dtype=np.int64)
    dx, di = Dev(x), Dev(idx)
    out = Dev(shape=(len(idx), 33))
    _capi.check(lib.mt_gather_rows_f32(out.ptr, dx.ptr, di.ptr, len(idx), 33, 50))
    np.testing.assert_array_equal(out.get(), x[idx])
    g = rng.standard_normal((len(idx), 33)).astype(np.float32)
    full = Dev(np.zeros((50, 33), np.float32))
    dg = Dev(g)
    _capi.check(lib.mt_scatter_rows_assign_f32(full.ptr, dg.ptr, di.ptr, len(idx), 33, 50))
    want = np.zeros((50, 33), np.float32)
    want[idx] = g                                                         # duplicates: last one wins (Q10)
    np.testing.assert_array_equal(full.get(), want)


def gemm_err(got, ref):
    return float(np.abs(got.astype(np.float64) - ref).max() / (np.abs(ref).max() + 1e-30))


@pytest.mark.parametrize("M,N,K", [(200, 3, 2), (200, 100, 2), (96, 10, 64), (65, 33, 17), (512, 256, 128),
                                   (4096, 10, 512), (300, 132, 100)])
@pytest.mark.parametrize("engine", [0, -1])
def test_linear_fwd_bwd_small(lib, M, N, K, engine):
    rng = np.random.default_rng(9)
    X = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) * 0.1).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    dY = rng.standard_normal((M, N)).astype(np.float32)
    lib.mt_gemm_set_engine(engine)
    try:
        Xd, Wd, bd, dYd = Dev(X), Dev(W), Dev(b), Dev(dY)
        Y = Dev(shape=(M, N))
        _capi.check(lib.mt_linear_fwd_f32(Y.ptr, Xd.ptr, Wd.ptr, bd.ptr, M, N, K, 1))
        ref = np.tanh(X.astype(np.float64) @ W.astype(np.float64).T + b)
        assert gemm_err(Y.get(), ref) < 1e-4
        Yact = Y.get()
        dX, dW = Dev(shape=(M, K)), Dev(np.ones((N, K), np.float32))
        _capi.check(lib.mt_linear_bwd_f32(dX.ptr, dW.ptr, dYd.ptr, Y.ptr, Xd.ptr, Wd.ptr, M, N, K, 1))
        dz = dY.astype(np.float64) * (1 - Yact.astype(np.float64) ** 2)
        assert gemm_err(dW.get(), dz.T @ X + 1.0) < 1e-4
        assert gemm_err(dX.get(), dz @ W) < 1e-4
    finally:
        lib.mt_gemm_set_engine(-1)


def test_tcgen05_engine_is_used_and_accurate(lib):
    """The Linear GEMMs of an MLP-sized layer must run on the tensor cores (engine 1) and match
    float64 to 1e-4 (the error-compensated split keeps ~fp32 accuracy; plain TF32 would be ~1e-3)."""
    rng = np.random.default_rng(10)
    M, N, K = 1024, 512, 768
    X = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) * 0.05).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    dY = rng.standard_normal((M, N)).astype(np.float32)
    Xd, Wd, bd, dYd, Y = Dev(X), Dev(W), Dev(b), Dev(dY), Dev(shape=(M, N))
    eng = C.c_int(-5)
    _capi.check(lib.mt_linear_fwd_f32(Y.ptr, Xd.ptr, Wd.ptr, bd.ptr, M, N, K, 1))
    lib.mt_gemm_last_engine(C.byref(eng))
    assert eng.value == 1
    ref = np.tanh(X.astype(np.float64) @ W.astype(np.float64).T + b)
    e = gemm_err(Y.get(), ref)
    assert e < 2e-5, e            # fp32-level (plain TF32 would be ~1e-3), far below the 1e-4 bar
    dX, dW = Dev(shape=(M, K)), Dev(shape=(N, K))
    _capi.check(lib.mt_linear_bwd_f32(dX.ptr, dW.ptr, dYd.ptr, Y.ptr, Xd.ptr, Wd.ptr, M, N, K, 0))
    lib.mt_gemm_last_engine(C.byref(eng))
    assert eng.value == 1
    dz = dY.astype(np.float64) * (1 - Y.get().astype(np.float64) ** 2)
    assert gemm_err(dW.get(), dz.T @ X) < 1e-5
    assert gemm_err(dX.get(), dz @ W) < 1e-5


RAGGED = [(1000, 260, 100), (129, 68, 4100), (131, 516, 1028), (2049, 257, 260), (128, 64, 2052), (257, 4100, 36),
          (4099, 132, 132), (640, 1100, 516), (128, 4096, 8196), (8200, 200, 64)]


@pytest.mark.parametrize("M,N,K", RAGGED)
def test_linear_ragged_shapes_on_tensor_cores(lib, M, N, K):
    """Tile-ragged extents (M, N not multiples of 128/256, K not a multiple of 32, split-K and paired-CTA shapes) through
    the tcgen05 path - forward with bias+tanh, dW with accumulation, dX with the producer's tanh' epilogue - vs float64."""
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    X = np.tanh(rng.standard_normal((M, K))).astype(np.float32)
    W = (rng.standard_normal((N, K)) * (1.0 / np.sqrt(K))).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    dY = rng.standard_normal((M, N)).astype(np.float32)
    Xd, Wd, bd, dYd, Y = Dev(X), Dev(W), Dev(b), Dev(dY), Dev(shape=(M, N))
    eng = C.c_int(-5)
    # the shapes are below the size at which the dispatcher prefers tcgen05, so the engine is forced where the kernel can
    # take the shape at all (it stores 128-bit rows: N % 4 != 0 is served by the SIMT tile - still checked below)
    lib.mt_gemm_set_engine(1 if N % 4 == 0 else -1)
    try:
        _capi.check(lib.mt_linear_fwd_f32(Y.ptr, Xd.ptr, Wd.ptr, bd.ptr, M, N, K, 1))
        lib.mt_gemm_last_engine(C.byref(eng))
        assert eng.value == (1 if N % 4 == 0 else 0), "unexpected GEMM engine for this shape"
        Yh = Y.get()
        assert gemm_err(Yh, np.tanh(X.astype(np.float64) @ W.astype(np.float64).T + b)) < 2e-5
        dX, dW = Dev(shape=(M, K)), Dev(np.full((N, K), 0.5, np.float32))
        _capi.check(lib.mt_linear_bwd_f32(dX.ptr, dW.ptr, dYd.ptr, Y.ptr, Xd.ptr, Wd.ptr, M, N, K, 1 | 2))
        lib.mt_gemm_last_engine(C.byref(eng))
        assert eng.value == (1 if N % 4 == 0 else 0)
    finally:
        lib.mt_gemm_set_engine(-1)
    dz = dY.astype(np.float64) * (1 - Yh.astype(np.float64) ** 2)
    assert gemm_err(dW.get(), dz.T @ X + 0.5) < 2e-5
    assert gemm_err(dX.get(), (dz @ W) * (1 - X.astype(np.float64) ** 2)) < 2e-5


@pytest.mark.parametrize("M,N,K", [(1024, 512, 768), (640, 1100, 516), (2049, 260, 1028)])
def test_tensor_core_split_modes_agree(lib, M, N, K):
    """The default split (hi.hi on TF32, the two first-order corrections as BF16 products) against the all-TF32 split
    (3xTF32) and float64: forward and both backward GEMMs, every operand-major combination.  Both stay at fp32 level
    (1e-5 here; the contract is 1e-4), so they also agree with each other to that level."""
    lib.mt_gemm_tc_set_mix.argtypes = [C.c_int]
    lib.mt_gemm_tc_set_fwd.argtypes = [C.c_int, C.c_int]
    rng = np.random.default_rng(M + N + K)
    X = np.tanh(rng.standard_normal((M, K))).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    dZ = rng.standard_normal((M, N)).astype(np.float32)
    Xd, Wd, bd, dZd = Dev(X), Dev(W), Dev(b), Dev(dZ)
    Yr = X.astype(np.float64) @ W.astype(np.float64).T + b
    dWr = dZ.astype(np.float64).T @ X.astype(np.float64)
    dXr = dZ.astype(np.float64) @ W.astype(np.float64)
    outs = {}
    lib.mt_gemm_set_engine(1)
    lib.mt_gemm_tc_set_fwd(-1, -1)            # the forward GEMM follows `mix` too (by default it is always 3xTF32)
    try:
        for mix in (1, 0):
            lib.mt_gemm_tc_set_mix(mix)
            Y, dX, dW = Dev(shape=(M, N)), Dev(shape=(M, K)), Dev(shape=(N, K))
            _capi.check(lib.mt_linear_fwd_f32(Y.ptr, Xd.ptr, Wd.ptr, bd.ptr, M, N, K, 0))
            _capi.check(lib.mt_linear_bwd_f32(dX.ptr, dW.ptr, dZd.ptr, None, Xd.ptr, Wd.ptr, M, N, K, 0))
            outs[mix] = (Y.get(), dW.get(), dX.get())
            for got, ref in zip(outs[mix], (Yr, dWr, dXr)):
                assert gemm_err(got, ref) < 1e-5
    finally:
        lib.mt_gemm_set_precision(1)          # back to the library default
        lib.mt_gemm_set_engine(-1)
    for a, bb in zip(outs[1], outs[0]):
        assert gemm_err(a, bb.astype(np.float64)) < 1e-5


def test_matmul_generic_strided_and_batched(lib):
    rng = np.random.default_rng(11)
    A = rng.standard_normal((4, 33, 20)).astype(np.float32)
    B = rng.standard_normal((20, 17)).astype(np.float32)
    Ad, Bd, Cd = Dev(A), Dev(B), Dev(shape=(4, 33, 17))
    _capi.check(lib.mt_matmul_f32(Cd.ptr, Ad.ptr, Bd.ptr, 4, 33, 17, 20, 33 * 20, 20, 1, 0, 17, 1))
    assert gemm_err(Cd.get(), A.astype(np.float64) @ B) < 1e-5
    # transposed operand views: C = A[0]^T-view @ B^T-view
    At = rng.standard_normal((20, 33)).astype(np.float32)     # logical A = At.T  (33 x 20)
    Bt = rng.standard_normal((17, 20)).astype(np.float32)     # logical B = Bt.T  (20 x 17)
    Atd, Btd, C2 = Dev(At), Dev(Bt), Dev(shape=(33, 17))
    _capi.check(lib.mt_matmul_f32(C2.ptr, Atd.ptr, Btd.ptr, 1, 33, 17, 20, 0, 1, 33, 0, 1, 20))
    assert gemm_err(C2.get(), At.T.astype(np.float64) @ Bt.T) < 1e-5


def test_pool_reuses_blocks(lib):
    r0, u0, p0, n0, d0 = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_uint64(), C.c_uint64()
    a = Dev(shape=(1 << 20,))
    a.free()
    lib.mt_pool_stats(C.byref(r0), C.byref(u0), C.byref(p0), C.byref(n0), C.byref(d0))
    b = Dev(shape=(1 << 20,))
    r1, d1 = C.c_size_t(), C.c_uint64()
    lib.mt_pool_stats(C.byref(r1), None, None, None, C.byref(d1))
    assert d1.value == d0.value and r1.value == r0.value      # served from the cache, no new cudaMalloc
    b.free()


@pytest.mark.parametrize("M,N,K", [(4096, 10, 4096), (1000, 3, 256), (2048, 16, 1024), (777, 7, 512), (1001, 10, 1100),
                                   (258, 12, 132), (5000, 10, 2048)])
def test_skinny_head_kernels_match_generic_tile_and_float64(lib, M, N, K):
    """The HBM-bound tiny-N kernels (4096->10 head) against float64 and against the generic SIMT tile,
    including the dX epilogue that emits the producer's pre-activation gradient.  Both prefetch schemes
    (cp.async ring / register double buffer) are run; ragged row blocks and column chunks included."""
    lib.mt_gemm_set_skinny.argtypes = [C.c_int]
    lib.mt_gemm_set_skinny_ring.argtypes = [C.c_int]
    rng = np.random.default_rng(12)
    X = np.tanh(rng.standard_normal((M, K))).astype(np.float32)       # a tanh output, as in a hidden layer
    W = (rng.standard_normal((N, K)) * 0.1).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    dY = rng.standard_normal((M, N)).astype(np.float32)
    Xd, Wd, bd, dYd = Dev(X), Dev(W), Dev(b), Dev(dY)
    outs = {}
    try:
        for mode, (skinny, ring) in {"ring": (1, 1), "regs": (1, 0), "generic": (0, 1)}.items():
            lib.mt_gemm_set_skinny(skinny)
            lib.mt_gemm_set_skinny_ring(ring)
            Y = Dev(shape=(M, N))
            _capi.check(lib.mt_linear_fwd_f32(Y.ptr, Xd.ptr, Wd.ptr, bd.ptr, M, N, K, 1))
            dX, dW = Dev(shape=(M, K)), Dev(np.ones((N, K), np.float32))
            _capi.check(lib.mt_linear_bwd_f32(dX.ptr, dW.ptr, dYd.ptr, Y.ptr, Xd.ptr, Wd.ptr, M, N, K, 1 | 2))
            outs[mode] = (Y.get(), dX.get(), dW.get())
    finally:
        lib.mt_gemm_set_skinny(1)
        lib.mt_gemm_set_skinny_ring(1)
    Yr = np.tanh(X.astype(np.float64) @ W.astype(np.float64).T + b)
    dz = dY.astype(np.float64) * (1 - outs["ring"][0].astype(np.float64) ** 2)
    dXr = (dz @ W) * (1 - X.astype(np.float64) ** 2)
    dWr = dz.T @ X + 1.0
    for got in outs.values():
        assert gemm_err(got[0], Yr) < 1e-5
        assert gemm_err(got[1], dXr) < 1e-5
        assert gemm_err(got[2], dWr) < 1e-5
    np.testing.assert_array_equal(outs["ring"][0], outs["regs"][0])      # same summation order in both schemes
    np.testing.assert_array_equal(outs["ring"][2], outs["regs"][2])
