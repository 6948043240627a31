#!/usr/bin/env python
"""Where does data-parallel time go?  Same process, same GPUs, three schedules interleaved (run under torchrun):

  replicas   no communicator at all: every rank trains on its own shard - each GPU's own pace (chip-to-chip clock
             spread under the power cap shows up here, it is not communication)
  overlap    the shipped schedule: per-parameter all-reduce + update on the communication stream, started as soon as
             that gradient is final, hidden behind the remaining backward GEMMs
  serial     round 1's schedule: all-reduce after the whole backward, one fused SGD launch

Each round times `--steps` steps of every schedule back to back (CUDA events per rank, barrier before and after);
rank 0 prints one JSON line with, per schedule, every rank's median ms/step and the max over ranks - the data-parallel
overhead is  max_over_ranks(overlap) / max_over_ranks(replicas) - 1, free of box-to-box and chip-to-chip variance.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           tools/dp_ab.py [--workload c2|c5] [--rounds 4] [--steps 6]
"""
import argparse
import ctypes as C
import json
import os
import statistics
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"])
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--rounds", type=int, default=4)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--max-units", type=int, default=0, help="cap the persistent GEMM grid at this many CTA pairs")
    ap.add_argument("--sync-each-step", action="store_true", help="read the loss back every step (ranks re-align)")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from microtorch_b200 import _capi, workloads as W
    _capi.require_device(local)
    lib = _capi.lib()
    lib.mt_gemm_tc_set_max_units.argtypes = [C.c_int]
    lib.mt_gemm_tc_set_max_units(args.max_units)
    import microtorch_b200.nn as nn
    from microtorch_b200.nn.activation import Tanh
    from microtorch_b200.nn.loss import CrossEntropyLoss, MSELoss
    from microtorch_b200.tensor import Tensor
    from microtorch_b200.trainer import Communicator, DataParallelTrainer

    wl = W.scaled(W.C2 if args.workload == "c2" else W.C5, args.batch)
    rng = np.random.default_rng(1000 + rank)
    x = rng.standard_normal((args.batch, wl.dims[0]), dtype=np.float32)
    if wl.loss == "ce":
        t = np.eye(wl.dims[-1], dtype=np.float32)[rng.integers(0, wl.dims[-1], args.batch)]
    else:
        t = rng.uniform(-1, 1, (args.batch, wl.dims[-1])).astype(np.float32)
    X, T = Tensor(x, device="cuda"), Tensor(t, device="cuda")

    def exchange(uid):
        box = [uid]
        dist.broadcast_object_list(box, src=0)
        return box[0]
    def allgather(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out
    comm = Communicator(rank, world, exchange, allgather)

    def make(schedule):
        np.random.seed(wl.seed)
        model = W.build_model(wl, nn.Linear, Tanh, nn.Sequential)
        loss = CrossEntropyLoss() if wl.loss == "ce" else MSELoss()
        return DataParallelTrainer(model, loss, wl.lr, None if schedule == "replicas" else comm, wl.pred_map, wl.squeeze_dim,
                                   overlap=(schedule != "serial"))
    trainers = {s: make(s) for s in ("replicas", "overlap", "serial")}

    def barrier():
        _capi.check(lib.mt_sync())
        dist.barrier()

    e0, e1 = C.c_void_p(), C.c_void_p()
    lib.mt_event_create(C.byref(e0))
    lib.mt_event_create(C.byref(e1))
    for tr in trainers.values():
        for _ in range(3):
            tr.step(X, T)
    times = {s: [] for s in trainers}
    for _ in range(args.rounds):
        for s, tr in trainers.items():
            barrier()
            lib.mt_event_record(e0)
            for _ in range(args.steps):
                loss = tr.step(X, T)
                if args.sync_each_step:
                    loss.item()
            lib.mt_event_record(e1)
            lib.mt_event_sync(e1)
            ms = C.c_float()
            lib.mt_event_elapsed_ms(e0, e1, C.byref(ms))
            times[s].append(ms.value / args.steps)
    barrier()
    mine = torch.tensor([statistics.median(times[s]) for s in trainers], dtype=torch.float64)
    allr = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allr, mine)
    if rank == 0:
        out = {"workload": args.workload, "world": world, "per_gpu_batch": args.batch, "rounds": args.rounds,
               "steps_per_round": args.steps, "max_units": args.max_units, "sync_each_step": args.sync_each_step,
               "nccl_max_ctas": os.environ.get("NCCL_MAX_CTAS")}
        for i, s in enumerate(trainers):
            per_rank = [round(float(a[i]), 3) for a in allr]
            out[s] = {"ms_per_step_by_rank": per_rank, "max_over_ranks": max(per_rank), "min_over_ranks": min(per_rank)}
        base = out["replicas"]["max_over_ranks"]
        out["overhead_vs_replicas"] = {s: round(out[s]["max_over_ranks"] / base - 1.0, 4) for s in ("overlap", "serial")}
        out["chip_spread"] = round(out["replicas"]["max_over_ranks"] / out["replicas"]["min_over_ranks"] - 1.0, 4)
        print(json.dumps(out), flush=True)
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
