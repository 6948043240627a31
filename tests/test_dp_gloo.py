"""Data-parallel host logic on CPU: world_size-2 and -4 gloo processes, Device.CPU replicas.

Checks what the N>1 path adds on top of the single-GPU step - row sharding, the global-count
loss scale, the per-parameter gradient hooks and exchange order, and that bias gradients are
never exchanged - by comparing N-rank post-SGD weights with a single-process full-batch step
and with the oracle."""
import os
import socket
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from microtorch_b200 import workloads as W
    from microtorch_b200.nn import Linear, Sequential
    from microtorch_b200.nn.activation import Tanh
    from microtorch_b200.nn.loss import MSELoss
    from microtorch_b200.tensor import Tensor
    from microtorch_b200.trainer import DataParallelTrainer, GlooCommunicator, shard_rows
    wl = W.scaled(W.C5, 64, (32,) * 5)
    model, x, t = W.setup(wl, Linear, Tanh, Sequential)          # same seed on every rank: identical replicas
    rows = shard_rows(x.shape[0], rank, world)
    comm = GlooCommunicator()
    tr = DataParallelTrainer(model, MSELoss(), wl.lr, comm)
    losses = []
    for _ in range(2):
        loss = tr.step(Tensor(x[rows]), Tensor(t[rows]))
        losses.append(tr.global_loss(loss))
    params = [p.data.copy() for p in model.parameters()]
    grads = [p.grad.copy() for p in model.parameters()]
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), losses=np.array(losses), exchanged=tr.exchanged_bytes,
             **{f"p{i}": p for i, p in enumerate(params)}, **{f"g{i}": g for i, g in enumerate(grads)})
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_n_rank_dp_matches_single_process_and_oracle(tmp_path, world):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / f"rank{world - 1}.npz")
    from microtorch_b200 import workloads as W
    from oracle import workloads as OW
    wl = W.scaled(W.C5, 64, (32,) * 5)
    ref = OW.run_steps(wl, 2)
    n = len(ref["params"])
    for i in range(n):
        np.testing.assert_array_equal(r0[f"p{i}"], r1[f"p{i}"])               # replicas stay identical
        np.testing.assert_allclose(r0[f"p{i}"], ref["params"][i], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(r0[f"g{i}"], ref["grads"][i], rtol=1e-4, atol=1e-8)
    np.testing.assert_allclose(r0["losses"], ref["losses"], rtol=1e-5)
    # only the 4 weight matrices (32x32 fp32) travel, twice; biases never do (Q1)
    assert int(r0["exchanged"]) == 2 * 4 * 32 * 32 * 4


def test_shard_rows():
    from microtorch_b200.trainer import shard_rows
    assert shard_rows(8, 1, 4) == slice(2, 4)
    with pytest.raises(ValueError):
        shard_rows(10, 0, 4)
