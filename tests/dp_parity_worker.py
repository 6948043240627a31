#!/usr/bin/env python
"""Worker of tests/test_dp_gpu.py (lives under tests/ because it checks against the oracle).
Data-parallel parity on real GPUs (launched by torchrun, one rank per GPU): N-rank post-SGD weights and
global loss against the oracle (CPU) on a small global batch; 1e-4 norm-wise.  Rank 0 prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def main():
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from microtorch_b200 import _capi, workloads as W
    _capi.require_device(local)
    import microtorch_b200.nn as nn
    from microtorch_b200.nn.activation import Tanh
    from microtorch_b200.nn.loss import MSELoss
    from microtorch_b200.tensor import Tensor
    from microtorch_b200.trainer import Communicator, DataParallelTrainer, shard_rows

    wl = W.scaled(W.C5, 1024, (256,) * 4)
    model, x, t = W.setup(wl, nn.Linear, Tanh, nn.Sequential)      # same seed everywhere: identical replicas + data
    rows = shard_rows(x.shape[0], rank, world)

    def exchange(uid):
        box = [uid]
        dist.broadcast_object_list(box, src=0)
        return box[0]
    def allgather(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out
    comm = Communicator(rank, world, exchange, allgather)
    tr = DataParallelTrainer(model, MSELoss(), wl.lr, comm)
    X, T = Tensor(x[rows], device="cuda"), Tensor(t[rows], device="cuda")
    losses = []
    for _ in range(3):
        losses.append(tr.global_loss(tr.step(X, T)))
    params = [p.data.get() for p in model.parameters()]
    # replicas must be bit-identical after the exchange (the peer-memory path sums in a fixed rank order)
    import hashlib
    digests = [None] * world
    dist.all_gather_object(digests, hashlib.sha256(b"".join(p.tobytes() for p in params)).hexdigest())
    if rank == 0:
        from oracle import workloads as OW
        ref = OW.run_steps(wl, 3)
        err = max(float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30)) for a, b in zip(params, ref["params"]))
        lerr = float(np.abs(np.array(losses) - ref["losses"]).max() / np.abs(ref["losses"]).max())
        print("DP_PARITY " + json.dumps({"world": world, "param_err": err, "loss_err": lerr, "losses": losses,
                                         "exchanged_bytes": tr.exchanged_bytes, "peer_memory": bool(comm.p2p),
                                         "p2p_fallbacks": comm.p2p_fallbacks,
                                         "replicas_bit_identical": len(set(digests)) == 1}), flush=True)
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
