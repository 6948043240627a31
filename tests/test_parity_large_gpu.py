"""Parity of the CUDA path against the ORACLE at BASELINE.json sizes, on RANDOM data.

The small-shape goldens in test_parity_gpu.py pin the arithmetic; these tests pin the code paths that only large
inputs take - grid-stride loops, two-pass reductions, split-K dW with K = 16 384 ... 131 072 rows, the CTA-pair
GEMMs at full width - against the NumPy restatement of the reference on the same seeded inputs.

Tolerances (north_star): exact arithmetic (+ - * and x**2) bit-exact; transcendental elementwise and reductions
1e-5 relative; GEMM paths 1e-4 norm-wise.  Needs a B200: `-m gpu` under gpurun.
"""
import numpy as np
import pytest

from microtorch_b200 import workloads as W

pytestmark = pytest.mark.gpu

RED = dict(rtol=1e-5, atol=0.0)


@pytest.fixture(scope="module")
def mt():
    from microtorch_b200 import _capi
    _capi.require_device()
    import microtorch_b200.nn as nn
    from microtorch_b200.nn.activation import Tanh
    from microtorch_b200.nn.loss import BCELoss, CrossEntropyLoss, MSELoss
    from microtorch_b200.nn.optimizer import SGD
    from microtorch_b200.tensor import Tensor

    class NS:
        pass
    ns = NS()
    ns.Tensor, ns.Linear, ns.Sequential, ns.Tanh, ns.SGD = Tensor, nn.Linear, nn.Sequential, Tanh, SGD
    ns.losses = {"mse": MSELoss, "ce": CrossEntropyLoss, "bce": BCELoss}
    ns.lib = _capi.lib()
    return ns


def host(a):
    return a.get() if hasattr(a, "get") else np.asarray(a)


def gemm_close(got, ref, tol=1e-4):
    got, ref = host(got).astype(np.float64), np.asarray(ref, dtype=np.float64)
    scale = np.abs(ref).max() + 1e-30
    err = np.abs(got - ref).max() / scale
    assert err <= tol, err
    return err


def c3_inputs(log2n, layout):
    """rng(0), x, y ~ U(0.5, 1.5) fp32; x is (N/4096, 4096) (1-D for 'flat')."""
    n = 1 << log2n
    rng = np.random.default_rng(0)
    rows = n // 4096
    x = rng.uniform(0.5, 1.5, (rows, 4096)).astype(np.float32)
    if layout == "flat":
        return x.reshape(-1), rng.uniform(0.5, 1.5, (n,)).astype(np.float32)
    yshape = {"same": (rows, 4096), "row": (1, 4096), "col": (rows, 1), "scalar": ()}[layout]
    return x, rng.uniform(0.5, 1.5, yshape).astype(np.float32)


# ----------------------------------------------------------------------------- config #3 (base.py:8-58, overload.py:68-128)
@pytest.mark.parametrize("log2n", [20, 24, 26])
@pytest.mark.parametrize("layout", ["same", "flat", "row", "col", "scalar"])
def test_c3_broadcast_poly_random_vs_oracle(mt, log2n, layout):
    """z = x*y + x**2 - y; z.sum().backward() for the four broadcast layouts of y (plus 1-D).  Forward and every
    elementwise gradient are exact fp32 arithmetic -> bit-exact; the un-broadcast reductions of y.grad -> 1e-5."""
    from oracle.tape import Var
    xv, yv = c3_inputs(log2n, layout)
    ox, oy = Var(xv, requires_grad=True), Var(yv, requires_grad=True)
    oz = ox * oy + ox ** 2 - oy
    osum = oz.sum()
    osum.backward()

    x = mt.Tensor(xv, requires_grad=True, device="cuda")
    y = mt.Tensor(yv, requires_grad=True, device="cuda")
    z = x * y + x ** 2 - y
    s = z.sum()
    s.backward()
    np.testing.assert_array_equal(host(z.data), oz.data)
    np.testing.assert_allclose(host(s.data), osum.data, **RED)
    np.testing.assert_array_equal(host(x.grad), ox.grad)                 # y + 2x: two exact roundings on both sides
    assert y.grad.shape == yv.shape
    if layout in ("same", "flat"):
        np.testing.assert_array_equal(host(y.grad), oy.grad)
    else:
        # y.grad = sum(x) - count over the broadcast axes: two reductions of magnitude ~count that nearly cancel
        # (x ~ U(0.5, 1.5)), so the 1e-5 is relative to the size of the reduced terms, not to their small difference
        count = xv.size // max(yv.size, 1)
        np.testing.assert_allclose(host(y.grad), oy.grad, rtol=1e-5, atol=1e-5 * count)


@pytest.mark.parametrize("log2n", [20, 24, 26])
def test_c3_log_tanh_sum_mean_random_vs_oracle(mt, log2n):
    """x.log().sum(), x.tanh().sum(), x.sum(), x.mean() with backward (math.py:6-23,68-85; reduce.py:8-43)."""
    from oracle.tape import Var
    xv, _ = c3_inputs(log2n, "same")
    for name in ("log", "tanh", "sum", "mean"):
        ox = Var(xv, requires_grad=True)
        x = mt.Tensor(xv, requires_grad=True, device="cuda")
        if name in ("log", "tanh"):
            omid, mid = getattr(ox, name)(), getattr(x, name)()
            np.testing.assert_allclose(host(mid.data), omid.data, rtol=1e-5, atol=1e-7, err_msg=name)
            oout, out = omid.sum(), mid.sum()
        elif name == "sum":
            oout, out = ox.sum(), x.sum()
        else:
            oout, out = ox.mean(), x.mean()
        oout.backward()
        out.backward()
        np.testing.assert_allclose(host(out.data), oout.data, err_msg=name, **RED)
        np.testing.assert_allclose(host(x.grad), ox.grad, rtol=1e-5, atol=1e-12, err_msg=name)


@pytest.mark.parametrize("log2n", [20, 24, 26])
@pytest.mark.parametrize("axis,keepdims", [(0, False), (0, True), (1, False), (1, True)])
def test_c3_axis_sums_random_vs_oracle(mt, log2n, axis, keepdims):
    """x.sum(axis, keepdims) forward and backward with a non-trivial upstream gradient (reduce.py:8-38):
    tall (N/4096 rows) column reductions and 4096-wide row reductions."""
    from oracle.tape import Var
    xv, _ = c3_inputs(log2n, "same")
    ox = Var(xv, requires_grad=True)
    x = mt.Tensor(xv, requires_grad=True, device="cuda")
    oout, out = ox.sum(axis=axis, keepdims=keepdims), x.sum(axis=axis, keepdims=keepdims)
    assert out.shape == oout.shape
    np.testing.assert_allclose(host(out.data), oout.data, **RED)
    wv = np.random.default_rng(1).uniform(0.5, 1.5, oout.shape).astype(np.float32)
    (oout * Var(wv)).sum().backward()
    (out * mt.Tensor(wv, device="cuda")).sum().backward()
    np.testing.assert_array_equal(host(x.grad), ox.grad)      # ones * expand_dims(g): a broadcast copy of wv


def test_c3_tall_thin_column_reduction_vs_numpy(mt):
    """The un-broadcast of a (2^22, 8) gradient onto a (1, 8) operand: the tall-thin shape VERDICT r1 measured at 17 %
    of HBM with millions of shared-memory bank conflicts - pinned for correctness here, timed in tools/c3_sweep.py."""
    rng = np.random.default_rng(3)
    for rows, cols in ((1 << 22, 8), (1 << 20, 32), (1 << 16, 512), (131072, 2048)):
        gv = rng.uniform(0.5, 1.5, (rows, cols)).astype(np.float32)
        bv = rng.uniform(0.5, 1.5, (1, cols)).astype(np.float32)
        g = mt.Tensor(gv, device="cuda")
        b = mt.Tensor(bv, requires_grad=True, device="cuda")
        ((g + b) * g).sum().backward()
        np.testing.assert_allclose(host(b.grad), gv.sum(axis=0, keepdims=True, dtype=np.float64), rtol=1e-5)


# ----------------------------------------------------------------------------- config #5 at full width
def _c5_float64_truth(wl):
    """The same training step in float64 (host): the yardstick that tells device error from ORACLE error."""
    from oracle import workloads as OW
    omodel, ox, ot = OW.setup(wl)
    params = [p.data.astype(np.float64) for p in omodel.parameters()]
    Ws, bs = params[0::2], params[1::2]
    h = [ox.data.astype(np.float64)]
    for Wl, b in zip(Ws, bs):
        h.append(np.tanh(h[-1] @ Wl.T + b))
    g = 2 * (h[-1] - ot.data.astype(np.float64)) / h[-1].size
    grads = [None] * len(Ws)
    for i in reversed(range(len(Ws))):
        dz = g * (1 - h[i + 1] ** 2)
        grads[i] = dz.T @ h[i]
        g = dz @ Ws[i]
    return grads


def test_c5_full_width_one_step_vs_oracle(mt):
    """BASELINE config #5's model at FULL width - 8 x [Linear(4096, 4096), Tanh], MSELoss, SGD - on 4096 rows (the CPU
    oracle needs ~15 s): loss, all 8 weight gradients and post-SGD weights against the oracle (overload.py:130-164).

    This stack (gain ~3.7 per layer with the reference's uniform(-1,1)*0.1 init) amplifies FORWARD rounding about 100x
    into every gradient: the NumPy oracle itself sits 4e-5 (norm-wise) from a float64 evaluation of the same step
    (profiles/r02_gemm_accuracy.jsonl).  Per GEMM every precision level is >= 70x inside the 1e-4 tolerance; end to end:
      "higher"  (the DEFAULT: 3xTF32 forward GEMMs)  must meet the north_star 1e-4 against the oracle  (measured 7.5e-5)
      "highest" (3xTF32 everywhere, SGEMM-class)     must meet it too                                  (measured 5.3e-5)
      "high"    (the opt-in fast mode)               measured 1.27e-4 against the oracle, 1.26e-4 against float64:
                asserted <= 2e-4 here - which is why it is not the default."""
    import ctypes as C
    from oracle import workloads as OW
    from microtorch_b200 import _capi
    from microtorch_b200.trainer import train_step
    wl = W.scaled(W.C5, 4096)
    ref = OW.run_steps(wl, 1)
    truth = _c5_float64_truth(wl)
    oracle_vs_truth = max(gemm_close(ref["grads"][2 * i], truth[i], tol=1.0) for i in range(8))
    measured = {}
    try:
        for level, tol in ((None, 1e-4), ("highest", 1e-4), ("high", 2e-4)):
            if level is not None:                      # None = whatever the library starts with: must be "higher"
                _capi.set_float32_matmul_precision(level)
            model, x, t = W.setup(wl, mt.Linear, mt.Tanh, mt.Sequential)
            X, T = mt.Tensor(x, device="cuda"), mt.Tensor(t, device="cuda")
            loss = train_step(model, mt.losses["mse"](), mt.SGD(model.parameters, lr=wl.lr), X, T)
            np.testing.assert_allclose(loss.item(), ref["losses"][0], rtol=1e-4)
            params = list(model.parameters())
            assert len(params) == 16
            errs = []
            for i, p in enumerate(params):
                if i % 2 == 0:
                    errs.append(gemm_close(p.grad, ref["grads"][i], tol=tol))
                    gemm_close(p.data, ref["params"][i], tol=tol)
                else:
                    np.testing.assert_array_equal(host(p.data), ref["params"][i])      # biases never move (Q1)
            measured[level] = (max(errs), max(gemm_close(p.grad, truth[i // 2], tol=1.0) for i, p in enumerate(params) if i % 2 == 0))
            engine = C.c_int(-1)
            mt.lib.mt_gemm_last_engine(C.byref(engine))
            assert engine.value == 1, "the 4096-wide layers must run on the tcgen05 engine"
    finally:
        _capi.set_float32_matmul_precision("higher")
    print("c5 full width: max dW error vs oracle / vs float64:", measured, "oracle vs float64:", oracle_vs_truth)
    # the default level is no further from the TRUTH than 2.5x the oracle's own distance
    assert measured[None][1] <= 2.5 * oracle_vs_truth + 1e-5, (measured, oracle_vs_truth)


def test_c5_full_width_layer_by_layer_default_precision(mt):
    """Default precision, teacher-forced: every layer of config #5 at full width gets the ORACLE's input activation and
    the ORACLE's output gradient, so no error is carried from layer to layer - forward output, dW and dX of each of the
    8 layers within the north_star's 1e-4 of the oracle's (measured ~2e-6: the per-GEMM error)."""
    from oracle import workloads as OW
    from oracle.tape import Var
    wl = W.scaled(W.C5, 4096)
    omodel, ox, ot = OW.setup(wl)
    model, x, t = W.setup(wl, mt.Linear, mt.Tanh, mt.Sequential)
    oh = ox
    rng = np.random.default_rng(4)
    for li in range(8):
        olin = omodel.modules[2 * li]
        oin = Var(oh.data.copy(), requires_grad=True)
        oout = olin(oin).tanh()
        gv = (rng.standard_normal(oout.shape) * 1e-7).astype(np.float32)          # the size of real MSE gradients here
        oout.backward(gv)
        lin = model.modules[2 * li]
        inp = mt.Tensor(oin.data, requires_grad=True, device="cuda")
        sub = mt.Sequential(lin, mt.Tanh())
        out = sub(inp)
        out.backward(mt.Tensor(gv, device="cuda").data)
        lin_now = sub.modules[0]
        gemm_close(out.data, oout.data)
        gemm_close(lin_now.weight.grad, olin.weight.grad)
        gemm_close(inp.grad, oin.grad)
        oh = oout


# ----------------------------------------------------------------------------- config #4 with backward
def test_c4_step_with_backward_32x512x2048_vs_oracle(mt):
    """BASELINE config #4 - 3-D input (32, 512, 2048) through 2 x [Linear(2048, 2048), Tanh], pred*0.49+0.5, BCELoss,
    backward, SGD - against the oracle (repairs S1-S3): the dW GEMMs reduce over K = 16 384 flattened rows and the
    fused BCE runs on 33.5 M elements (overload.py:130-164 + S3, loss.py:87-100)."""
    from oracle import workloads as OW
    from microtorch_b200.trainer import train_step
    wl = W.Workload("c4_32x512", (2048, 2048, 2048), "bce", 0.01, 32, pred_map=(0.49, 0.5), seed=0,
                    input_shape=(32, 512, 2048))
    ref = OW.run_steps(wl, 1)
    model, x, t = W.setup(wl, mt.Linear, mt.Tanh, mt.Sequential)
    X, T = mt.Tensor(x, device="cuda"), mt.Tensor(t, device="cuda")
    loss = train_step(model, mt.losses["bce"](), mt.SGD(model.parameters, lr=wl.lr), X, T, wl.pred_map)
    np.testing.assert_allclose(loss.item(), ref["losses"][0], rtol=1e-5)
    for i, p in enumerate(model.parameters()):
        if i % 2 == 0:
            gemm_close(p.grad, ref["grads"][i])
            gemm_close(p.data, ref["params"][i])
        else:
            np.testing.assert_array_equal(host(p.data), ref["params"][i])


def test_c4_full_size_backward_properties(mt):
    """Config #4 at its FULL size (256 x 512 x 2048, K = 131 072 rows in every dW GEMM): the oracle cannot run it in
    the suite's time, so the step is pinned through (i) the first 8 rows of dW against a float64 NumPy product over all
    131 072 rows, (ii) linearity: the gradients of the two halves of the batch add up to the full-batch gradients."""
    wl = W.Workload("c4_full", (2048, 2048, 2048), "bce", 0.01, 256, pred_map=(0.49, 0.5), seed=0,
                    input_shape=(256, 512, 2048))
    model, x, t = W.setup(wl, mt.Linear, mt.Tanh, mt.Sequential)
    loss_fn = mt.losses["bce"]()

    def grads_of(xs, ts, world):
        loss_fn.world = world
        model.zero_grad()
        pred = model(mt.Tensor(xs, device="cuda")) * wl.pred_map[0] + wl.pred_map[1]
        keep = {}
        l = loss_fn(pred, mt.Tensor(ts, device="cuda"))
        l.backward()
        keep["loss"] = l.item()
        return [host(p.grad) for p in model.parameters()][0::2], keep

    full, k_full = grads_of(x, t, 1)
    a, k_a = grads_of(x[:128], t[:128], 2)
    b, k_b = grads_of(x[128:], t[128:], 2)
    loss_fn.world = 1
    for gf, ga, gb in zip(full, a, b):
        gemm_close(ga.astype(np.float64) + gb, gf)
    assert abs((k_a["loss"] + k_b["loss"]) - k_full["loss"]) <= 1e-5 * abs(k_full["loss"])
    # (i) rows of the FIRST layer's dW straight from the definition dW1 = dZ1^T X, float64 on the host:
    #     dZ1 is recovered from the device (it is the gradient that reaches layer 1's pre-activation), X is the input
    model.zero_grad()
    X = mt.Tensor(x, device="cuda")
    lin1, lin2 = model.modules[0], model.modules[2]
    h1 = lin1(X).tanh()
    h1_leaf = mt.Tensor(h1.data, requires_grad=True, device="cuda")
    pred = lin2(h1_leaf).tanh() * wl.pred_map[0] + wl.pred_map[1]
    loss_fn(pred, mt.Tensor(t, device="cuda")).backward()
    dh1 = host(h1_leaf.grad).reshape(-1, 2048)
    y1 = host(h1.data).reshape(-1, 2048)
    dz1_rows = (dh1[:, :8].astype(np.float64) * (1.0 - y1[:, :8].astype(np.float64) ** 2))       # (131072, 8)
    want = dz1_rows.T @ x.reshape(-1, 2048).astype(np.float64)                                   # (8, 2048)
    gemm_close(full[0][:8], want)


def test_fused_bce_268m_elements_vs_oracle_formula(mt):
    """The fused loss kernel on config #4's 268 M elements (256 x 512 x 2048) against the oracle's BCELoss evaluated
    chunk by chunk on the host (loss.py:87-100; 16 chunks of 2^24 - elementwise work is chunk-independent, the
    mean is re-assembled in float64)."""
    from oracle.nnref import LOSSES
    from oracle.tape import Var
    n, chunk = 256 * 512 * 2048, 1 << 24
    rng = np.random.default_rng(11)
    pv = rng.uniform(0.02, 0.98, n).astype(np.float32)
    tv = (rng.random(n) > 0.5).astype(np.float32)
    p = mt.Tensor(pv.reshape(256, 512, 2048), requires_grad=True, device="cuda")
    loss = mt.losses["bce"]()(p, mt.Tensor(tv.reshape(256, 512, 2048), device="cuda"))
    loss.backward()
    got_grad = host(p.grad).reshape(-1)
    total = 0.0
    for c0 in range(0, n, chunk):
        op = Var(pv[c0:c0 + chunk], requires_grad=True)
        ol = LOSSES["bce"](op, Var(tv[c0:c0 + chunk]))
        ol.backward()
        total += float(ol.data) * (chunk / n)
        # d(mean over the chunk) * chunk/n == d(mean over everything); chunk/n is a power of two, so this is exact
        np.testing.assert_allclose(got_grad[c0:c0 + chunk], op.grad * np.float32(chunk / n), rtol=1e-5, atol=0)
    assert abs(loss.item() - total) <= 1e-5 * abs(total)


# ----------------------------------------------------------------------------- parity edges
def test_loss_sgd_zero_on_odd_offset_slices(mt):
    """mini-batch code slices X[i:i+b] / T[i:i+b]; a basic slice is a zero-copy view whose pointer is only
    4-byte aligned.  The fused loss, the multi-tensor SGD and fill must take them (scalar path), like the reference."""
    from oracle.nnref import LOSSES
    from oracle.tape import Var
    rng = np.random.default_rng(5)
    pv = rng.uniform(0.1, 0.9, (403, 1)).astype(np.float32)
    tv = (rng.random((403, 1)) > 0.5).astype(np.float32)
    P, T = mt.Tensor(pv, device="cuda"), mt.Tensor(tv, device="cuda")
    for kind in ("mse", "ce", "bce"):
        for i, b in ((50, 101), (1, 7), (3, 400), (0, 403)):
            ps = mt.Tensor(P.data[i:i + b], requires_grad=True, device="cuda")
            assert ps.data.ptr == P.data.ptr + 4 * i                         # really a view
            l = mt.losses[kind]()(ps, T[i:i + b])
            l.backward()
            op = Var(pv[i:i + b], requires_grad=True)
            ol = LOSSES[kind](op, Var(tv[i:i + b]))
            ol.backward()
            np.testing.assert_allclose(l.item(), ol.data, rtol=1e-5, err_msg=f"{kind} {i}:{i + b}")
            np.testing.assert_allclose(host(ps.grad), op.grad, rtol=1e-5, atol=1e-9, err_msg=f"{kind} {i}:{i + b}")
    # SGD on mis-aligned parameter / gradient views, bit-exact
    from microtorch_b200.cuda import kernels as K
    from microtorch_b200.cuda.array import DeviceArray
    wv = rng.standard_normal(1001).astype(np.float32)
    gv = rng.standard_normal(1001).astype(np.float32)
    Wd, Gd = DeviceArray.from_host(wv), DeviceArray.from_host(gv)
    K.sgd_multi([Wd[1:998], Wd[998:1001]], [Gd[3:1000], Gd[0:3]], 0.37)
    want = wv.copy()
    want[1:998] = wv[1:998] + (-(np.float32(0.37) * gv[3:1000]))
    want[998:1001] = wv[998:1001] + (-(np.float32(0.37) * gv[0:3]))
    np.testing.assert_array_equal(Wd.get(), want)
    K.zero_multi([Wd[1:6], Wd[7:1001]])
    want[1:6] = 0
    want[7:1001] = 0
    np.testing.assert_array_equal(Wd.get(), want)
    Wd[5:].fill(2.5)
    want[5:] = 2.5
    np.testing.assert_array_equal(Wd.get(), want)


def test_matmul_1d_operands_forward_backward_on_cuda(mt):
    """The 1-D branches of matmul (overload.py:143,154: np.outer(...).squeeze()) on the device vs the oracle:
    matrix @ vector, vector @ matrix, vector @ vector."""
    from oracle.tape import Var, matmul
    rng = np.random.default_rng(9)
    cases = [((7, 5), (5,)), ((5,), (5, 6)), ((5,), (5,)), ((300, 129), (129,)), ((129,), (129, 300))]
    for sa, sb in cases:
        av, bv = rng.standard_normal(sa).astype(np.float32), rng.standard_normal(sb).astype(np.float32)
        oa, ob = Var(av, requires_grad=True), Var(bv, requires_grad=True)
        oc = matmul(oa, ob)
        a, b = mt.Tensor(av, requires_grad=True, device="cuda"), mt.Tensor(bv, requires_grad=True, device="cuda")
        c = a @ b
        assert c.shape == oc.shape, (sa, sb)
        np.testing.assert_allclose(host(c.data), oc.data, rtol=1e-5, atol=1e-5)
        gv = rng.standard_normal(oc.shape).astype(np.float32)
        oc.backward(gv)
        c.backward(mt.Tensor(gv, device="cuda").data)
        assert a.grad.shape == sa and b.grad.shape == sb
        np.testing.assert_allclose(host(a.grad), oa.grad, rtol=1e-5, atol=1e-5, err_msg=f"dA {sa}@{sb}")
        np.testing.assert_allclose(host(b.grad), ob.grad, rtol=1e-5, atol=1e-5, err_msg=f"dB {sa}@{sb}")


def test_clip_grad_norm_matches_numpy_for_matrices(mt):
    """Tensor.clip_grad_norm (reference tensor.py:297-310) calls np.linalg.norm(grad, norm_type): the vector p-norm
    for 1-D gradients and the MATRIX norm for 2-D ones (norm_type 2 = largest singular value).  The CUDA path now
    gives the same scale factor as the reference's NumPy path for both."""
    rng = np.random.default_rng(12)
    for shape, ord_ in (((64, 48), 2.0), ((48, 64), 2.0), ((64, 48), 1), ((64, 48), np.inf), ((777,), 2.0), ((777,), 1),
                        ((777,), 3.0)):
        gv = rng.standard_normal(shape).astype(np.float32)
        res = {}
        for dev in ("cpu", "cuda"):
            t = mt.Tensor(np.zeros(shape, np.float32), requires_grad=True, device=dev)
            t.grad = mt.Tensor(gv, device=dev).data
            t.clip_grad_norm(max_norm=0.5, norm_type=ord_)
            res[dev] = host(t.grad)
        want = gv * np.float32(0.5 / (np.linalg.norm(gv, ord_) + 1e-6))
        np.testing.assert_allclose(res["cpu"], want, rtol=1e-6)
        np.testing.assert_allclose(res["cuda"], res["cpu"], rtol=1e-5, err_msg=f"{shape} ord={ord_}")
    t = mt.Tensor(np.zeros((2, 3, 4), np.float32), requires_grad=True, device="cuda")
    t.grad = mt.Tensor(np.ones((2, 3, 4), np.float32), device="cuda").data
    with pytest.raises(ValueError):                      # NumPy: "Improper number of dimensions to norm."
        t.clip_grad_norm(1.0)


def test_infinite_gemm_operand_behaviour_is_pinned(mt):
    """an INFINITE operand of the tcgen05 split-fp32 GEMM yields NaN in every output that touches
    it (lo = x - trunc(x) = inf - inf) where NumPy's SGEMM yields +-inf; everything else stays finite and correct, and
    NaN operands give NaN on both.  Reachable only by a run that has already diverged.  Pinned so that it cannot widen
    silently: outputs that do not touch the bad element must match NumPy to 1e-4."""
    rng = np.random.default_rng(2)
    M, N, K = 512, 512, 1024                              # 2^28 multiply-adds: the tensor-core engine
    av, bv = rng.standard_normal((M, K)).astype(np.float32), rng.standard_normal((K, N)).astype(np.float32)
    av[3, 17] = np.inf
    av[100, 5] = np.nan
    import ctypes as C
    a, b = mt.Tensor(av, device="cuda"), mt.Tensor(bv, device="cuda")
    c = host((a @ b).data)
    engine = C.c_int(-1)
    mt.lib.mt_gemm_last_engine(C.byref(engine))
    assert engine.value == 1
    with np.errstate(invalid="ignore"):
        ref = av @ bv
    clean = np.ones(M, bool)
    clean[[3, 100]] = False
    gemm_close(c[clean], ref[clean])
    assert np.isnan(c[100]).all() and np.isnan(ref[100]).all()
    assert np.isinf(ref[3]).all()
    assert not np.isfinite(c[3]).any()                    # NaN (today) or +-inf (NumPy): never a finite wrong number
