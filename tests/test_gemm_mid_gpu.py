"""The mid-size tcgen05 kernel (csrc/mt_gemm_tc_mid.cu) through the raw C ABI, against float64 - the Linear forward
(bias + tanh epilogue), dW (accumulating, both operands MN-major) and dX (with the producer's tanh' epilogue) products
that replace `a.data @ b.data`, `a.T @ grad`, `grad @ b.T` of the reference (src/tensor/ops/overload.py:132,142,153 via
src/nn/layer/linear.py:48), for both tile flavours, every slice count and every precision level; the dispatcher's window;
run-to-run determinism; the short / empty K slice case that once dead-locked the register hand-back."""
import ctypes as C

import numpy as np
import pytest

from microtorch_b200 import _capi
from capi_util import Dev

pytestmark = pytest.mark.gpu

GEMM_TOL = 2e-5      # max |err| / max |ref|: fp32-level, far inside the north_star's 1e-4 for GEMM paths


@pytest.fixture(scope="module")
def lib():
    _capi.require_device(0)
    return _capi.lib()


def gemm_err(got, ref):
    return float(np.max(np.abs(got.astype(np.float64) - ref)) / max(np.max(np.abs(ref)), 1e-30))


def mid_launches(lib):
    n = C.c_uint64()
    lib.mt_gemm_tc_mid_launches(C.byref(n))
    return n.value


class forced:
    """engine 2 = the mid kernel only (an unsupported shape is an error, never a silent fallback)"""

    def __init__(self, lib, tile=0, splits=0, precision=1):
        self.lib, self.tile, self.splits, self.precision = lib, tile, splits, precision

    def __enter__(self):
        self.lib.mt_gemm_set_precision(self.precision)
        self.lib.mt_gemm_tc_mid_set(1, 0, 0, self.splits)
        self.lib.mt_gemm_tc_mid_set_tile(self.tile, 0)
        self.lib.mt_gemm_set_engine(2)

    def __exit__(self, *exc):
        self.lib.mt_gemm_set_engine(-1)
        self.lib.mt_gemm_tc_mid_set(1, 0, 0, 0)
        self.lib.mt_gemm_tc_mid_set_tile(0, 0)
        self.lib.mt_gemm_set_precision(1)


def layer(M, N, K, seed):
    rng = np.random.default_rng(seed)
    X = np.tanh(rng.standard_normal((M, K))).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    dY = rng.standard_normal((M, N)).astype(np.float32)
    return X, W, b, dY


def run_layer(lib, X, W, b, dY):
    """forward (bias + tanh), dW accumulated onto 0.5, dX times tanh'(X) - three launches of the GEMM under test"""
    (M, K), N = X.shape, W.shape[0]
    Xd, Wd, bd, dYd = Dev(X), Dev(W), Dev(b), Dev(dY)
    Y, dX, dW = Dev(shape=(M, N)), Dev(shape=(M, K)), Dev(np.full((N, K), 0.5, np.float32))
    _capi.check(lib.mt_linear_fwd_f32(Y.ptr, Xd.ptr, Wd.ptr, bd.ptr, M, N, K, 1))
    _capi.check(lib.mt_linear_bwd_f32(dX.ptr, dW.ptr, dYd.ptr, None, Xd.ptr, Wd.ptr, M, N, K, 1 | 2))
    _capi.check(lib.mt_sync())
    return Y.get(), dW.get(), dX.get()


def check_layer(lib, M, N, K, tile, splits, precision):
    X, W, b, dY = layer(M, N, K, M + 3 * N + 7 * K)
    X64, W64, dY64 = X.astype(np.float64), W.astype(np.float64), dY.astype(np.float64)
    n0 = mid_launches(lib)
    with forced(lib, tile, splits, precision):
        Y, dW, dX = run_layer(lib, X, W, b, dY)
    assert mid_launches(lib) - n0 == 3, "the forced mid kernel did not run all three products"
    assert gemm_err(Y, np.tanh(X64 @ W64.T + b)) < GEMM_TOL
    assert gemm_err(dW, dY64.T @ X64 + 0.5) < GEMM_TOL
    assert gemm_err(dX, (dY64 @ W64) * (1 - X64 ** 2)) < GEMM_TOL


SHAPES = [(256, 128, 64), (1024, 512, 768), (1000, 260, 100), (129, 68, 4100), (640, 1100, 516), (2049, 260, 264),
          (1024, 1024, 1024)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("tile,splits", [(0, 0), (1, 1), (1, 3), (1, 8), (2, 1), (2, 2), (2, 4)])
def test_mid_kernel_matches_float64(lib, M, N, K, tile, splits):
    """ragged M / N / K (TMA zero-fill, predicated stores), K slices that do not divide K (empty and short slices),
    single-CTA and CTA-pair tiles, at the default precision (forward 3xTF32, backward TF32 + 2 BF16)"""
    check_layer(lib, M, N, K, tile, splits, 1)


@pytest.mark.parametrize("precision", [0, 2])
@pytest.mark.parametrize("tile,splits", [(1, 2), (2, 3)])
def test_mid_kernel_precision_levels(lib, precision, tile, splits):
    check_layer(lib, 640, 1100, 516, tile, splits, precision)


@pytest.mark.parametrize("tile", [1, 2])
def test_mid_kernel_bare_matmul_transposed_views(lib, tile):
    """A MN-major x B K-major: the one major pair the Linear products never produce (overload.py:132 on `.T` views)"""
    M, N, K = 512, 384, 640
    rng = np.random.default_rng(5)
    At = rng.standard_normal((K, M)).astype(np.float32)      # A(m, k) = At[k, m]
    Bt = rng.standard_normal((N, K)).astype(np.float32)      # B(k, n) = Bt[n, k]
    Ad, Bd, Cd = Dev(At), Dev(Bt), Dev(shape=(M, N))
    with forced(lib, tile):
        _capi.check(lib.mt_matmul_f32(Cd.ptr, Ad.ptr, Bd.ptr, 1, M, N, K, 0, 1, M, 0, 1, K))
        _capi.check(lib.mt_sync())
    assert gemm_err(Cd.get(), At.astype(np.float64).T @ Bt.astype(np.float64).T) < GEMM_TOL


def test_dispatcher_window(lib):
    """auto mode: the mid kernel takes 2^25 <= M N K < 2^32 multiply-adds (with M >= 128, N >= 64, K >= 32); smaller
    problems stay on the SIMT tile, larger ones on the persistent CTA-pair kernel; a backward with an activation to undo
    makes dZ explicit first (the mid kernel has no prologue flavour) and still runs both products on it"""
    def fwd(M, N, K):
        X, W, b, _ = layer(M, N, K, 1)
        Xd, Wd, bd, Y = Dev(X), Dev(W), Dev(b), Dev(shape=(M, N))
        n0 = mid_launches(lib)
        _capi.check(lib.mt_linear_fwd_f32(Y.ptr, Xd.ptr, Wd.ptr, bd.ptr, M, N, K, 1))
        eng = C.c_int(-5)
        lib.mt_gemm_last_engine(C.byref(eng))
        return mid_launches(lib) - n0, eng.value

    assert fwd(256, 128, 128) == (0, 0)              # 2^22: SIMT
    assert fwd(1024, 512, 768) == (1, 1)             # 2^28.6: mid
    assert fwd(4096, 1024, 512) == (1, 1)            # 2^31: mid
    assert fwd(8192, 1024, 1024) == (0, 1)           # 2^33: persistent pair kernel
    M, N, K = 2048, 256, 512
    X, W, b, dY = layer(M, N, K, 2)
    Xd, Wd, dYd = Dev(X), Dev(W), Dev(dY)
    Yact = np.tanh(X.astype(np.float64) @ W.astype(np.float64).T).astype(np.float32)
    Yd, dX, dW = Dev(Yact), Dev(shape=(M, K)), Dev(shape=(N, K))
    n0 = mid_launches(lib)
    _capi.check(lib.mt_linear_bwd_f32(dX.ptr, dW.ptr, dYd.ptr, Yd.ptr, Xd.ptr, Wd.ptr, M, N, K, 0))
    assert mid_launches(lib) - n0 == 2
    dz = dY.astype(np.float64) * (1 - Yact.astype(np.float64) ** 2)
    assert gemm_err(dW.get(), dz.T @ X.astype(np.float64)) < GEMM_TOL
    assert gemm_err(dX.get(), dz @ W.astype(np.float64)) < GEMM_TOL


def test_mid_kernel_is_deterministic(lib):
    """slices are folded in slice order through distributed shared memory: bit-identical from run to run"""
    X, W, b, dY = layer(1024, 512, 768, 3)
    with forced(lib, 1, 4):
        a = run_layer(lib, X, W, b, dY)
        c = run_layer(lib, X, W, b, dY)
    for u, v in zip(a, c):
        assert np.array_equal(u, v)


@pytest.mark.parametrize("tile,splits", [(1, 4), (1, 8), (2, 4)])
def test_short_and_empty_slices_do_not_deadlock(lib, tile, splits):
    """K = 32 is ONE k-block: every slice but the first is empty, and a CTA with nothing to do reaches the register
    hand-back before its high-register warps have grown (the hang fixed by the named barrier in front of setmaxnreg.inc)"""
    M, N, K = 1024, 1024, 32
    X, W, b, _ = layer(M, N, K, 4)
    Xd, Wd, bd, Y = Dev(X), Dev(W), Dev(b), Dev(shape=(M, N))
    with forced(lib, tile, splits):
        for _ in range(50):
            _capi.check(lib.mt_linear_fwd_f32(Y.ptr, Xd.ptr, Wd.ptr, bd.ptr, M, N, K, 1))
        _capi.check(lib.mt_sync())
    assert gemm_err(Y.get(), np.tanh(X.astype(np.float64) @ W.astype(np.float64).T + b)) < GEMM_TOL


# ----------------------------------------------------------------------------- through the public API
def test_mid_size_training_steps_vs_oracle_and_graph_replay():
    """A training step whose GEMMs all fall into the mid-size window (MLP 256 -> 384 -> 512 -> 128 with Tanh, 2048 rows:
    2^26 - 2^28.6 multiply-adds each) through the unchanged MicroTorch API (demo.ipynb:5178-5189): loss, every weight
    gradient and post-SGD weight of 3 steps against the oracle (the NumPy restatement of the reference's step), 1e-4
    norm-wise; then the same step captured as ONE CUDA graph (cluster launches inside stream capture) replays the eager
    loss stream bit for bit."""
    import microtorch_b200.nn as nn
    from microtorch_b200 import workloads as W
    from microtorch_b200.graph import GraphedStep
    from microtorch_b200.nn.activation import Tanh
    from microtorch_b200.nn.loss import MSELoss
    from microtorch_b200.nn.optimizer import SGD
    from microtorch_b200.tensor import Tensor
    from microtorch_b200.trainer import train_step
    from oracle import workloads as OW

    lib = _capi.lib()
    wl = W.Workload("mid_mlp", (256, 384, 512, 128), "mse", 0.05, 2048, seed=11)
    ref = OW.run_steps(wl, 3)

    def fresh():
        model, x, t = W.setup(wl, nn.Linear, Tanh, nn.Sequential)
        return model, Tensor(x, device="cuda"), Tensor(t, device="cuda"), MSELoss(), SGD(model.parameters, lr=wl.lr)

    model, X, T, loss_fn, opt = fresh()
    n0 = mid_launches(lib)
    losses = [float(train_step(model, loss_fn, opt, X, T, wl.pred_map, wl.squeeze_dim).item()) for _ in range(3)]
    assert mid_launches(lib) - n0 == 3 * 8, "3 forward + 3 dW + 2 dX products per step belong to the mid-size kernel"
    np.testing.assert_allclose(np.array(losses, np.float32), ref["losses"], rtol=1e-4)
    params = list(model.parameters())
    for i, p in enumerate(params):
        got, want = p.data.get().astype(np.float64), ref["params"][i].astype(np.float64)
        assert np.abs(got - want).max() <= 1e-4 * np.abs(want).max()
        gg, gw = p.grad.get().astype(np.float64), ref["grads"][i].astype(np.float64)
        assert np.abs(gg - gw).max() <= 1e-4 * max(np.abs(gw).max(), 1e-30)

    # eager stream of 6 steps vs 2 eager warm-up steps + 4 graph replays
    model, X, T, loss_fn, opt = fresh()
    eager = [train_step(model, loss_fn, opt, X, T, wl.pred_map, wl.squeeze_dim).item() for _ in range(6)]
    model, X, T, loss_fn, opt = fresh()
    step = GraphedStep(lambda: train_step(model, loss_fn, opt, X, T, wl.pred_map, wl.squeeze_dim), warmup=2)
    replayed = []
    for _ in range(4):
        step.replay()
        replayed.append(step.result.item())
    step.close()
    assert replayed == eager[2:6]


@pytest.mark.parametrize("tile,splits", [(0, 0), (1, 8), (1, 1), (2, 4)])
def test_mid_kernel_long_k(lib, tile, splits):
    """K = 16 384 on a 256 x 256 output (the dW of a narrow layer over a long batch: overload.py:153 with K = rows): up to 512
    k-blocks per CTA, i.e. dozens of TMEM chunks through both accumulator buffers and every ring wrap-around"""
    check_layer(lib, 256, 256, 16384, tile, splits, 1)
