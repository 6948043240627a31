"""CPU-only tests of the host-side logic above the C ABI (no GPU, no compute calls):
the reduction planner, product/oracle separation, loud failure without the extension."""
import ast
import os

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def simulate(shape, axes):
    """Run the [outer, red, inner] passes of plan_reduction with NumPy and return the result."""
    from microtorch_b200.cuda.kernels import plan_reduction
    x = np.arange(int(np.prod(shape)), dtype=np.float64).reshape(shape) if shape else np.float64(3.0).reshape(())
    cur = x.reshape(-1)
    for outer, red, inner in plan_reduction(shape, axes):
        cur = cur.reshape(outer, red, inner).sum(axis=1).reshape(-1)
    kept = [d for i, d in enumerate(shape) if i not in {a % max(len(shape), 1) for a in axes}]
    return cur.reshape(kept), x


@pytest.mark.parametrize("shape,axes", [((6, 8), (0,)), ((6, 8), (1,)), ((6, 8), (0, 1)), ((3, 6, 8), (0, 2)),
                                        ((3, 6, 8), (1,)), ((2, 3, 4, 5), (0, 2)), ((2, 3, 4, 5), (1, 3)),
                                        ((1, 7, 1), (0, 2)), ((5, 1, 7), (1,)), ((4,), (0,)), ((2, 1, 3, 1, 4), (0, 2))])
def test_plan_reduction_matches_numpy(shape, axes):
    got, x = simulate(shape, axes)
    np.testing.assert_array_equal(got, x.sum(axis=tuple(axes)))


def test_plan_reduction_merges_adjacent_axes_into_one_pass():
    from microtorch_b200.cuda.kernels import plan_reduction
    assert plan_reduction((256, 512, 2048), (0, 1)) == [(1, 256 * 512, 2048)]      # C4: b3.grad = one column reduction
    assert plan_reduction((65536, 10), (0, 1)) == [(1, 655360, 1)]
    assert plan_reduction((64, 4096), (1,)) == [(64, 4096, 1)]
    assert plan_reduction((4, 1, 1), (1, 2)) == []


def _imports(path):
    tree = ast.parse(open(path).read(), path)
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            for a in node.names:
                yield a.name
        elif isinstance(node, ast.ImportFrom) and node.module:
            yield ("." * node.level) + node.module


def test_product_never_imports_the_oracle_or_torch():
    """The oracle is test infrastructure: nothing under microtorch_b200/ may import it (or fall back
    to torch for compute).  torch appears only in trainer.GlooCommunicator (CPU tests) and bench.py."""
    bad = []
    for root, _, files in os.walk(os.path.join(REPO, "microtorch_b200")):
        for f in files:
            if not f.endswith(".py"):
                continue
            path = os.path.join(root, f)
            for mod in _imports(path):
                top = mod.lstrip(".").split(".")[0]
                if top == "oracle":
                    bad.append((path, mod))
                if top in ("torch", "cupy", "triton") and not path.endswith("trainer.py"):
                    bad.append((path, mod))
    assert not bad, bad


def test_only_tests_smoke_and_bench_touch_the_oracle():
    """tools/ and the package must not import the oracle (it is the checker, never the thing measured or shipped)."""
    bad = []
    for sub in ("tools", "microtorch_b200"):
        for root, _, files in os.walk(os.path.join(REPO, sub)):
            for f in files:
                if f.endswith(".py"):
                    path = os.path.join(root, f)
                    if any(m.lstrip(".").split(".")[0] == "oracle" for m in _imports(path)):
                        bad.append(path)
    assert not bad, bad


def test_cuda_requested_without_backend_raises():
    from microtorch_b200 import _capi
    from microtorch_b200.tensor import Tensor
    from microtorch_b200.tensor import device as devmod
    if _capi.device_count() > 0:
        pytest.skip("a GPU is visible")
    assert devmod.CUDA_AVAILABLE is False
    with pytest.raises(ImportError):
        Tensor([1.0], device="cuda")
    with pytest.raises(ImportError):
        devmod._device("tpu")
    with pytest.raises(_capi.MicrotorchB200Error):          # enum bypasses the availability check, the backend does not
        Tensor([1.0], device=devmod.Device.CUDA)


def test_cpu_path_training_step_matches_oracle():
    """Device.CPU of this package (NumPy backend, topological backward) against the oracle."""
    from microtorch_b200 import workloads as W
    import microtorch_b200.nn as nn
    from microtorch_b200.nn.activation import Tanh
    from microtorch_b200.nn.loss import BCELoss, CrossEntropyLoss, MSELoss
    from microtorch_b200.nn.optimizer import SGD
    from microtorch_b200.tensor import Tensor
    from microtorch_b200.trainer import train_step
    from oracle import workloads as OW
    for wl, cls in [(W.C1A, MSELoss), (W.scaled(W.C2, 96, (48, 80, 64, 10)), CrossEntropyLoss),
                    (W.scaled(W.C4, 4, (32, 48, 32), input_shape=(4, 6, 32)), BCELoss)]:
        ref = OW.run_steps(wl, 3)
        model, x, t = W.setup(wl, nn.Linear, Tanh, nn.Sequential)
        X, T, opt, loss_fn = Tensor(x), Tensor(t), SGD(model.parameters, lr=wl.lr), cls()
        losses = [train_step(model, loss_fn, opt, X, T, wl.pred_map, wl.squeeze_dim).item() for _ in range(3)]
        np.testing.assert_allclose(losses, ref["losses"], rtol=1e-5)
        for p, want in zip(model.parameters(), ref["params"]):
            np.testing.assert_allclose(p.data, want, rtol=1e-5, atol=1e-7)


def test_checkpoint_round_trip_and_resume_cpu(tmp_path):
    """a model checkpoint resumes the trajectory bit-exactly (Device.CPU), loads in
    place (the optimizer keeps working on the same Parameter objects) and rejects mismatching files."""
    from microtorch_b200 import checkpoint as ck, workloads as W
    import microtorch_b200.nn as nn
    from microtorch_b200.nn.activation import Tanh
    from microtorch_b200.nn.loss import MSELoss
    from microtorch_b200.nn.optimizer import SGD
    from microtorch_b200.tensor import Tensor
    from microtorch_b200.trainer import train_step
    wl = W.C1A
    model, x, t = W.setup(wl, nn.Linear, Tanh, nn.Sequential)
    X, T, opt, loss_fn = Tensor(x), Tensor(t), SGD(model.parameters, lr=wl.lr), MSELoss()
    for _ in range(3):
        train_step(model, loss_fn, opt, X, T, wl.pred_map, wl.squeeze_dim)
    names = [n for n, _ in ck.named_parameters(model)]
    assert names[:2] == ["modules.0.weight", "modules.0.bias"] and len(names) == len(list(model.parameters()))
    path = tmp_path / "c1a.npz"
    ck.save_checkpoint(path, model, step=3, lr=wl.lr, extra={"note": "x"})
    want = [train_step(model, loss_fn, opt, X, T, wl.pred_map, wl.squeeze_dim).item() for _ in range(3)]
    final = [np.array(p.data, copy=True) for p in model.parameters()]

    fresh, _, _ = W.setup(wl, nn.Linear, Tanh, nn.Sequential)
    held = list(fresh.parameters())
    header = ck.load_checkpoint(path, fresh)
    assert header["step"] == 3 and header["lr"] == wl.lr and header["extra"] == {"note": "x"}
    assert [id(p) for p in fresh.parameters()] == [id(p) for p in held]          # loaded in place
    opt2 = SGD(fresh.parameters, lr=header["lr"])
    got = [train_step(fresh, loss_fn, opt2, X, T, wl.pred_map, wl.squeeze_dim).item() for _ in range(3)]
    assert got == want
    for p, w in zip(fresh.parameters(), final):
        np.testing.assert_array_equal(p.data, w)

    other = nn.Sequential(nn.Linear(2, 8), Tanh())
    with pytest.raises(KeyError):
        ck.load_checkpoint(path, other)
    bad = nn.Sequential(nn.Linear(2, 64), Tanh(), nn.Linear(64, 31), Tanh(), nn.Linear(32, 16), Tanh(), nn.Linear(16, 1), Tanh())
    with pytest.raises(ValueError):
        ck.load_checkpoint(path, bad)
    with pytest.raises(FileNotFoundError):
        ck.load_checkpoint(tmp_path / "absent.npz", fresh)
    junk = tmp_path / "junk.npz"
    junk.write_bytes(b"not a checkpoint")
    with pytest.raises(ValueError):
        ck.load_checkpoint(junk, fresh)


def test_fit_on_cpu_is_the_per_step_loop():
    """trainer.fit away from CUDA is the reference's loop verbatim (the single-launch kernel is CUDA-only and
    `fused=True` refuses instead of falling back silently)."""
    from microtorch_b200 import workloads as W
    import microtorch_b200.nn as nn
    from microtorch_b200.nn.activation import Tanh
    from microtorch_b200.nn.loss import MSELoss
    from microtorch_b200.nn.optimizer import SGD
    from microtorch_b200.tensor import Tensor
    from microtorch_b200.trainer import fit, train_step
    wl = W.C1A
    model, x, t = W.setup(wl, nn.Linear, Tanh, nn.Sequential)
    X, T = Tensor(x), Tensor(t)
    got = fit(model, MSELoss(), SGD(model.parameters, lr=wl.lr), X, T, 5, wl.pred_map, wl.squeeze_dim)
    model2, _, _ = W.setup(wl, nn.Linear, Tanh, nn.Sequential)
    opt2, loss2 = SGD(model2.parameters, lr=wl.lr), MSELoss()
    want = [train_step(model2, loss2, opt2, X, T, wl.pred_map, wl.squeeze_dim).item() for _ in range(5)]
    np.testing.assert_array_equal(got, np.array(want, np.float32))
    with pytest.raises(ValueError):
        fit(model, MSELoss(), SGD(model.parameters, lr=wl.lr), X, T, 1, wl.pred_map, wl.squeeze_dim, fused=True)


def test_relu_sigmoid_activations_cpu():
    """ReLU / Sigmoid (not in the reference) composed from Tensor ops: values and gradients vs closed forms."""
    from microtorch_b200.nn.activation import ReLU, Sigmoid
    from microtorch_b200.tensor import Tensor
    xv = np.random.default_rng(3).standard_normal((7, 5)).astype(np.float32)
    x = Tensor(xv, requires_grad=True)
    r = ReLU()(x)
    np.testing.assert_array_equal(r.data, np.maximum(xv, 0))
    r.sum().backward()
    np.testing.assert_array_equal(x.grad, (xv > 0).astype(np.float32))
    x = Tensor(xv, requires_grad=True)
    s = Sigmoid()(x)
    want = 1.0 / (1.0 + np.exp(-xv.astype(np.float64)))
    np.testing.assert_allclose(s.data, want, rtol=1e-6)
    s.sum().backward()
    np.testing.assert_allclose(x.grad, want * (1 - want), rtol=1e-5)


def test_parameter_hooks_fire_as_soon_as_their_gradient_is_final():
    """ADVICE r1 (tensor.py backward order): a Parameter is finalised the moment its last incoming edge has been
    processed - the hook of layer L must fire BEFORE the grad_fn of layer L-1 runs, otherwise every gradient
    all-reduce of the data-parallel trainer is queued behind the whole backward pass."""
    from microtorch_b200.tensor import Tensor
    from microtorch_b200.tensor.types import Leaf
    rng = np.random.default_rng(0)
    events = []
    x = Tensor(rng.standard_normal((4, 3)).astype(np.float32))
    weights = [Tensor(rng.standard_normal((3, 3)).astype(np.float32), requires_grad=True) for _ in range(3)]
    h = x
    for li, w in enumerate(weights):
        w._grad_hook = (lambda t, li=li: events.append(("hook", li)))
        inp = h

        def gw(g, li=li, inp=inp):
            events.append(("grad_fn_w", li))
            return inp.data.T @ g

        def gx(g, li=li, w=w):
            events.append(("grad_fn_x", li))
            return g @ w.data.T
        deps = [Leaf(value=w, grad_fn=gw)]
        if inp.requires_grad:
            deps.append(Leaf(value=inp, grad_fn=gx))
        h = Tensor(inp.data @ w.data, requires_grad=True, dependencies=deps)
    h.sum().backward()
    for li in (2, 1):
        assert events.index(("hook", li)) < events.index(("grad_fn_w", li - 1)), events
    assert [e for e in events if e[0] == "hook"] == [("hook", 2), ("hook", 1), ("hook", 0)]
    # and the values are what the plain sweep gives
    ref_h = [x.data]
    for w in weights:
        ref_h.append(ref_h[-1] @ w.data)
    g = np.ones_like(ref_h[-1])
    for li in (2, 1, 0):
        np.testing.assert_allclose(weights[li].grad, ref_h[li].T @ g, rtol=1e-6)
        g = g @ weights[li].data.T


def test_shared_parameter_is_finalised_after_its_last_use_only():
    """A tensor used twice receives the SUM of both edges, once, and its hook fires once."""
    from microtorch_b200.tensor import Tensor
    w = Tensor(np.array([2.0, 3.0], dtype=np.float32), requires_grad=True)
    fired = []
    w._grad_hook = lambda t: fired.append(np.array(t.grad))
    y = (w * w + w).sum()
    y.backward()
    assert len(fired) == 1
    np.testing.assert_allclose(fired[0], 2 * w.data + 1)
    np.testing.assert_allclose(w.grad, 2 * w.data + 1)


def test_single_launch_plan_rejects_a_target_of_another_shape():
    """fit() must not train on a target whose SHAPE differs from pred (same element count)."""
    from microtorch_b200.trainer import _small_mlp_plan
    import microtorch_b200.nn as nn
    from microtorch_b200.nn.activation import Tanh
    from microtorch_b200.nn.loss import MSELoss
    from microtorch_b200.nn.optimizer import SGD
    from microtorch_b200.tensor.device import Device

    class Fake:                                  # stands in for a CUDA tensor: only shape / device / ndim are read
        def __init__(self, shape):
            self.shape, self.device, self.ndim = shape, Device.CUDA, len(shape)
            self.size = int(np.prod(shape))
    model = nn.Sequential(nn.Linear(4, 6), Tanh(), nn.Linear(6, 3), Tanh())
    opt = SGD(model.parameters, lr=0.1)
    assert _small_mlp_plan(model, MSELoss(), opt, Fake((8, 4)), Fake((8, 3))) is not None
    assert _small_mlp_plan(model, MSELoss(), opt, Fake((8, 4)), Fake((3, 8))) is None      # transposed
    assert _small_mlp_plan(model, MSELoss(), opt, Fake((8, 4)), Fake((24,))) is None       # flattened
