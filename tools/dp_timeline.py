#!/usr/bin/env python
"""Device-side timeline of one data-parallel training step (nsys is not installed on this pool).

Launch like bench.py:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1
                              --master-port P tools/dp_timeline.py [--workload c2|c5] [--batch ROWS] [--no-overlap]

While tracing (mt_trace_*), every kernel launch of the library is followed by a timed CUDA event on its stream
(0 = compute, 1 = communication).  Every rank writes gpurun_out/dp_timeline_<tag>_rank<r>.json (all events of two
steps, ms since the first); rank 0 prints one JSON summary line: step time, when the last backward GEMM ended, when
each all-reduce began / ended, what was left exposed after the last compute kernel.  `--no-overlap` reproduces the
round-1 schedule (exchange after the whole backward, one fused SGD) for the A/B.
"""
import argparse
import ctypes as C
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"])
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--steps", type=int, default=10, help="timed steps (no tracing) for the step time")
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo", rank=rank, world_size=world)
    from microtorch_b200 import _capi, workloads as W
    _capi.require_device(local)
    lib = _capi.lib()
    import microtorch_b200.nn as nn
    from microtorch_b200.nn.activation import Tanh
    from microtorch_b200.nn.loss import CrossEntropyLoss, MSELoss
    from microtorch_b200.tensor import Tensor
    from microtorch_b200.trainer import Communicator, DataParallelTrainer

    wl = W.scaled(W.C2 if args.workload == "c2" else W.C5, args.batch)
    np.random.seed(wl.seed)
    model = W.build_model(wl, nn.Linear, Tanh, nn.Sequential)
    rng = np.random.default_rng(1000 + rank)
    x = rng.standard_normal((args.batch, wl.dims[0]), dtype=np.float32)
    if wl.loss == "ce":
        t = np.eye(wl.dims[-1], dtype=np.float32)[rng.integers(0, wl.dims[-1], args.batch)]
    else:
        t = rng.uniform(-1, 1, (args.batch, wl.dims[-1])).astype(np.float32)
    comm = None
    if world > 1:
        def exchange(uid):
            box = [uid]
            dist.broadcast_object_list(box, src=0)
            return box[0]
        def allgather(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out
        comm = Communicator(rank, world, exchange, allgather)
    trainer = DataParallelTrainer(model, CrossEntropyLoss() if wl.loss == "ce" else MSELoss(), wl.lr, comm, wl.pred_map,
                                  wl.squeeze_dim, overlap=not args.no_overlap)
    X, T = Tensor(x, device="cuda"), Tensor(t, device="cuda")

    def barrier():
        _capi.check(lib.mt_sync())
        if dist is not None:
            dist.barrier()

    for _ in range(args.warmup):
        trainer.step(X, T)
    barrier()
    e0, e1 = C.c_void_p(), C.c_void_p()
    lib.mt_event_create(C.byref(e0))
    lib.mt_event_create(C.byref(e1))
    lib.mt_event_record(e0)
    for _ in range(args.steps):
        trainer.step(X, T)
    lib.mt_event_record(e1)
    lib.mt_event_sync(e1)
    ms = C.c_float()
    lib.mt_event_elapsed_ms(e0, e1, C.byref(ms))
    barrier()
    step_ms = ms.value / args.steps
    if dist is not None:
        import torch
        tt = torch.tensor([step_ms], dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        step_ms_max = float(tt[0])
    else:
        step_ms_max = step_ms

    # ---- two traced steps
    lib.mt_trace_reset()
    lib.mt_trace_enable(1)
    for _ in range(2):
        trainer.step(X, T)
    _capi.check(lib.mt_comm_wait() if comm is not None else 0)
    _capi.check(lib.mt_sync())
    lib.mt_trace_enable(0)
    n = C.c_int()
    lib.mt_trace_count(C.byref(n))
    ev = []
    name = C.create_string_buffer(64)
    for i in range(n.value):
        sid, t_ms = C.c_int(), C.c_float()
        _capi.check(lib.mt_trace_get(i, name, 64, C.byref(sid), C.byref(t_ms)))
        ev.append({"label": name.value.decode(), "stream": sid.value, "ms": round(t_ms.value, 4)})
    tag = args.tag or f"{args.workload}_n{world}{'_noverlap' if args.no_overlap else ''}"
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    with open(os.path.join(REPO, "gpurun_out", f"dp_timeline_{tag}_rank{rank}.json"), "w") as f:
        json.dump({"rank": rank, "world": world, "step_ms": step_ms, "events": ev}, f)

    # ---- summary of the SECOND traced step (rank 0)
    if rank == 0:
        sgd_marks = [i for i, e in enumerate(ev) if e["label"].startswith("mt_sgd_multi") or e["label"] == "sgd_one"]
        comp = [e for e in ev if e["stream"] == 0]
        # split the compute events into steps at the fused loss kernel
        loss_idx = [i for i, e in enumerate(comp) if "loss" in e["label"]]
        summary = {"tag": tag, "world": world, "workload": args.workload, "per_gpu_batch": args.batch,
                   "overlap": not args.no_overlap, "step_ms_rank0": step_ms, "step_ms_max_over_ranks": step_ms_max,
                   "events_traced": len(ev)}
        if len(loss_idx) >= 2:
            start = comp[loss_idx[1]]["ms"]
            second = [e for e in ev if e["ms"] >= start]
            comp2 = [e for e in second if e["stream"] == 0]
            comm2 = [e for e in second if e["stream"] == 1]
            gemms = [e for e in comp2 if e["label"].startswith("tc_gemm") or "fold" in e["label"] or "skinny" in e["label"]]
            last_gemm_end = max([e["ms"] for e in gemms], default=None)
            ar = [(a["ms"], b["ms"]) for a, b in zip([e for e in comm2 if e["label"] == "allreduce_begin"],
                                                      [e for e in comm2 if e["label"] == "allreduce_end"])]
            end_all = max(e["ms"] for e in second)
            summary.update({
                "backward_from_loss_ms": None if last_gemm_end is None else round(last_gemm_end - start, 3),
                "allreduce_spans_ms_rel_loss": [[round(a - start, 3), round(b - start, 3)] for a, b in ar],
                "allreduce_total_ms": round(sum(b - a for a, b in ar), 3),
                "exposed_after_last_gemm_ms": None if last_gemm_end is None else round(end_all - last_gemm_end, 3),
                "comm_stream_events": [[e["label"], round(e["ms"] - start, 3)] for e in comm2],
            })
        print(json.dumps(summary), flush=True)
    if dist is not None:
        dist.barrier()
        if comm is not None:
            comm.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
