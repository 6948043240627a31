"""Random MLP training steps on the CUDA path against the oracle (the CPU restatement pinned bit-for-bit to the
reference): layer counts, widths (odd ones next to powers of two), batch sizes, 2-D and 3-D inputs and losses drawn from
a seed, so that the dispatcher's boundaries - SIMT tile / skinny head / mid-size tcgen05 kernel / persistent tcgen05
kernel, fused Linear+Tanh epilogues and the tanh' hand-over between layers, ragged M / N / K, the fused losses on
non-multiple-of-4 sizes - are crossed in combinations nobody listed by hand.  Two full steps (zero_grad, forward, loss,
backward, SGD); losses 1e-4 relative, last-step gradients and post-SGD parameters 1e-4 norm-wise (north_star, GEMM paths).
"""
import numpy as np
import pytest

from microtorch_b200 import workloads as W

pytestmark = pytest.mark.gpu

WIDTHS = [1, 2, 3, 10, 17, 33, 64, 100, 128, 200, 255, 256, 384, 512, 768, 1024]
BATCHES = [1, 5, 64, 100, 333, 512, 1000, 2048, 4096]
MAX_FLOPS_PER_STEP = 2.0e10          # keeps the oracle at a fraction of a second per step on the box's host cores


def random_workload(seed: int) -> W.Workload:
    rng = np.random.default_rng(1000 + seed)
    while True:
        n_layers = int(rng.integers(1, 5))
        dims = tuple(int(rng.choice(WIDTHS)) for _ in range(n_layers + 1))
        loss = ["mse", "ce", "bce"][int(rng.integers(0, 3))]
        pred_map = None if loss == "mse" else (0.49, 0.5)
        lr = [0.01, 0.1][int(rng.integers(0, 2))]
        if rng.random() < 0.25:
            b, s = int(rng.choice([2, 4, 8])), int(rng.choice([3, 16, 50, 128]))
            batch, input_shape = b, (b, s, dims[0])
        else:
            batch, input_shape = int(rng.choice(BATCHES)), None
        wl = W.Workload(f"fuzz_mlp_{seed}", dims, loss, lr, batch, pred_map, None, None, seed, input_shape)
        rows = batch if input_shape is None else input_shape[0] * input_shape[1]
        biggest = max(rows * a * b for a, b in zip(dims[:-1], dims[1:]))          # multiply-adds of the largest GEMM
        if seed % 2 == 0 and biggest < (1 << 25):
            continue                          # every other case must reach the tensor-core engines (2^25 multiply-adds up)
        if wl.gemm_flops_per_step() <= MAX_FLOPS_PER_STEP and rows * max(dims) <= (1 << 21):
            return wl


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
    return ns


def host(a):
    return a.get() if hasattr(a, "get") else np.asarray(a)


def close(got, ref, what, tol=1e-4):
    got, ref = host(got).astype(np.float64), np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} != {ref.shape}"
    scale = np.abs(ref).max() + 1e-30
    err = np.abs(got - ref).max()
    assert err <= tol * scale, f"{what}: off by {err / scale:.3e}"


@pytest.mark.parametrize("seed", range(40))
def test_random_mlp_training_steps_match_the_oracle(mt, seed):
    from microtorch_b200.trainer import train_step
    from oracle import workloads as OW
    wl = random_workload(seed)
    tag = f"{wl.dims} batch {wl.batch} input {wl.input_shape} {wl.loss} lr {wl.lr}"
    ref = OW.run_steps(wl, 2)
    model, x, t = W.setup(wl, mt.Linear, mt.Tanh, mt.Sequential)
    X, T = mt.Tensor(x, device="cuda"), mt.Tensor(t, device="cuda")
    loss_fn, opt = mt.losses[wl.loss](), mt.SGD(model.parameters, lr=wl.lr)
    losses = [host(train_step(model, loss_fn, opt, X, T, wl.pred_map).data) for _ in range(2)]
    np.testing.assert_allclose(np.array(losses, np.float32), ref["losses"], rtol=1e-4, err_msg=tag)
    params = list(model.parameters())
    assert len(params) == len(ref["params"])
    for i, p in enumerate(params):
        close(p.data, ref["params"][i], f"{tag}: parameter {i}")
        if np.abs(ref["grads"][i]).max() > 0:              # biases: identically zero (Q1) - compared exactly below
            close(p.grad, ref["grads"][i], f"{tag}: gradient {i}")
        else:
            assert not host(p.grad).any(), f"{tag}: gradient {i} must stay zero"
