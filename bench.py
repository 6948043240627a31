#!/usr/bin/env python
"""bench.py - MLP training-step throughput of the B200 backend (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one full training step (zero_grad, forward, loss, backward, SGD) of BASELINE.json
config #2 - synthetic MLP 1024->4096->4096->10, Tanh + the reference's CrossEntropyLoss on
pred*0.49+0.5, SGD, fp32 - on a per-GPU batch of 65536 rows.  N>1 is the data
parallel trainer: every rank keeps 65536 rows (weak scaling, global batch 65536*N), weight
gradients are all-reduced with NCCL.  Rank 0 prints ONE JSON line.

  value      samples/s over all GPUs, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        same metric through the public API with HOST inputs: every step uploads X,T from
             pinned host memory and reads the loss back
  roofline   the dominant kernel (tcgen05 split-fp32 GEMM): logical 2MNK flops / device time measured
             live with CUDA events around every GEMM launch of the timed region
  cpu_baseline   the oracle port of the reference's NumPy path on the host cores, bounded sample (batch 4096, 3 steps)
  --impl reference   the same oracle port at the FULL per-GPU batch of the workload (65536 rows, ~11 s per step on 16
             cores, ~25 GiB of host memory), so both arms run the same config; smaller only if the host lacks the memory
  dp_parity  (N > 1) before timing: 2 data-parallel steps on a small global batch against a single-process
             recompute of the same steps on rank 0 - the scaling record carries correctness, not only speed
  c5         BASELINE config #5 (8 x 4096-wide layers, MSE) at the same rows per GPU, a few steps, as a sub-object
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

CPU_SAMPLE_BATCH = 4096


def load_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16": d["bf16_tflops"], "bf16_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, period_ms=200):
        self.rows, self.proc, self.gpu, self.period_ms = [], None, gpu_index, period_ms

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", str(self.period_ms)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        return self.window(t0, t1)

    def window(self, t0, t1):
        """Clocks / throttle reasons of the samples taken in [t0, t1] (the sampler keeps running)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, reasons = [], [], set()
        for ts, line in list(self.rows):
            if ts < t0 or ts > t1 + 0.3:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def all_host_threads():
    """Let BLAS use every host core (torchrun exports OMP_NUM_THREADS=1 to its children)."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass


def ncu_traffic_per_launch():
    """dram__bytes_read+write per GEMM launch from the committed `ncu --set full` capture, or None."""
    import csv
    import glob
    paths = sorted(glob.glob(os.path.join(REPO, "profiles", "r*_gemm_ncu_full_summary*.csv")))
    if not paths:
        return None
    try:
        rows = list(csv.reader(open(paths[-1])))
        hdr = rows[0]
        r, w = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        vals = [(float(x[r]) + float(x[w])) * 1e9 for x in rows[2:] if x and x[r]]
        return sum(vals) / len(vals) if vals else None
    except Exception:
        return None


def cpu_baseline(steps=3, warmup=1):
    """The oracle port of the reference's NumPy training step, timed on this box's host cores."""
    all_host_threads()
    from microtorch_b200 import workloads as W
    from oracle import workloads as OW
    from oracle.nnref import train_step
    wl = W.scaled(W.C2, CPU_SAMPLE_BATCH)
    model, x, t = OW.setup(wl)
    best = None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        train_step(model, x, t, wl.loss, wl.lr, wl.pred_map, wl.squeeze_dim)
        dt = time.perf_counter() - t0
        if i >= warmup:
            best = dt if best is None else min(best, dt)
    try:
        from threadpoolctl import threadpool_info
        blas_threads = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        blas_threads = os.cpu_count()
    return {"value": CPU_SAMPLE_BATCH / best, "unit": "samples/s", "cores": int(blas_threads), "kind": "port",
            "host_cpus": os.cpu_count(),
            "sample": f"config #2 model at batch {CPU_SAMPLE_BATCH} (of 65536), best of {steps} steps after {warmup} warm-up; "
                      "NumPy elementwise is single-threaded, only matmul uses all cores"}


def reference_batch(want):
    """The per-step batch of the reference arm: the workload's own (65536 rows) when the host has the memory for the
    oracle's eager zero-filled gradients and temporaries (~25 GiB), else the largest power of two that fits."""
    try:
        import psutil
        avail = psutil.virtual_memory().available / 2 ** 30
    except Exception:
        avail = 0.0
    batch = want
    while batch > CPU_SAMPLE_BATCH and avail < 40.0 * batch / 65536 + 4.0:
        batch //= 2
    return batch, avail


def run_reference(args, emit):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; the reference
    is pure Python and cannot travel to the GPU box) at the SAME config as the GPU arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    all_host_threads()
    from microtorch_b200 import workloads as W
    from oracle import workloads as OW
    from oracle.nnref import train_step
    global CPU_SAMPLE_BATCH
    full = args.batch
    CPU_SAMPLE_BATCH, avail = reference_batch(full)
    wl = W.scaled(W.C2, CPU_SAMPLE_BATCH)
    model, x, t = OW.setup(wl)
    for _ in range(args.warmup):
        train_step(model, x, t, wl.loss, wl.lr, wl.pred_map, wl.squeeze_dim)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        train_step(model, x, t, wl.loss, wl.lr, wl.pred_map, wl.squeeze_dim)
    dt = time.perf_counter() - t0
    value = CPU_SAMPLE_BATCH * args.steps / dt
    try:
        from threadpoolctl import threadpool_info
        cores = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        cores = os.cpu_count()
    sample = (f"config #2 model, batch {CPU_SAMPLE_BATCH} per step " +
              ("(the full per-GPU batch: same config as the GPU arm)" if CPU_SAMPLE_BATCH == full else
               f"(bounded sample of {full}: host memory available {avail:.0f} GiB)"))
    emit({
        "impl": "reference", "metric": "mlp_train_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "c2_mlp_1024_4096_4096_10_tanh_ce_sgd", "per_gpu_batch": full, "cpu_sample_batch": CPU_SAMPLE_BATCH,
                   "same_config": CPU_SAMPLE_BATCH == full},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": int(cores), "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0})


def dp_parity_check(comm, dist, rank, world, ns):
    """Two data-parallel steps of the config #2 model (full widths) on a small global batch (256 rows per rank)
    against a single-process recompute of the same two steps on rank 0 (same backend, no communicator): norm-wise
    error of every updated weight and of the global loss.  SURVEY.md 8(e) 'Parity test', tolerance 1e-4."""
    W, nn, Tanh, Tensor = ns["W"], ns["nn"], ns["Tanh"], ns["Tensor"]
    from microtorch_b200.nn.loss import CrossEntropyLoss
    from microtorch_b200.nn.optimizer import SGD
    from microtorch_b200.trainer import DataParallelTrainer, shard_rows, train_step
    wl = W.scaled(W.C2, 256 * world)
    model, x, t = W.setup(wl, nn.Linear, Tanh, nn.Sequential)          # same seed everywhere: same replica, same data
    rows = shard_rows(x.shape[0], rank, world)
    tr = DataParallelTrainer(model, CrossEntropyLoss(), wl.lr, comm, wl.pred_map, wl.squeeze_dim)
    Xs, Ts = Tensor(x[rows], device="cuda"), Tensor(t[rows], device="cuda")
    losses = [tr.global_loss(tr.step(Xs, Ts)) for _ in range(2)]
    out = None
    if rank == 0:
        ref, x2, t2 = W.setup(wl, nn.Linear, Tanh, nn.Sequential)
        Xf, Tf = Tensor(x2, device="cuda"), Tensor(t2, device="cuda")
        loss_fn, opt = CrossEntropyLoss(), SGD(ref.parameters, lr=wl.lr)
        ref_losses = [float(train_step(ref, loss_fn, opt, Xf, Tf, wl.pred_map, wl.squeeze_dim).item()) for _ in range(2)]
        perr = 0.0
        for a, b in zip(model.parameters(), ref.parameters()):
            a, b = a.data.get().astype(np.float64), b.data.get().astype(np.float64)
            perr = max(perr, float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30)))
        lerr = max(abs(a - b) / abs(b) for a, b in zip(losses, ref_losses))
        out = {"param_err": perr, "loss_err": lerr, "steps": 2, "global_batch": 256 * world, "tolerance": 1e-4,
               "pass": bool(perr <= 1e-4 and lerr <= 1e-4),
               # gradients must become final (and their exchange start) last layer first, during the backward pass
               "exchange_order_last_layer_first":
               tr.hook_order == [id(p) for p in reversed(list(model.parameters())[0::2])],
               "against": "single-process recompute of the same steps on rank 0 (device path, no communicator)"}
    dist.barrier()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"])
    ap.add_argument("--batch", type=int, default=65536, help="rows per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c5", action="store_true", help="skip the config #5 sub-object")
    ap.add_argument("--precision", default="higher", choices=["high", "higher", "highest"],
                    help="fp32 emulation level of the tensor-core GEMMs (library default: higher)")
    ap.add_argument("--no-fast-mode", action="store_true", help="skip the precision=high sub-object")
    ap.add_argument("--no-mid", action="store_true", help="skip the mid-size-regime sub-object")
    args = ap.parse_args()
    # stdout carries exactly ONE line - the JSON result.  Libraries write banners to fd 1 (NCCL prints its version
    # there when NCCL_DEBUG asks for it), so everything else is sent to stderr for the duration of the run.
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(result_fd, (json.dumps(obj) + "\n").encode())

    if args.impl == "reference":
        return run_reference(args, emit)
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run --nproc-per-node N")
    dist = None
    if world > 1:
        import torch.distributed as dist      # control plane only (rendezvous, barrier, max over ranks)
        dist.init_process_group("gloo", rank=rank, world_size=world)

    from microtorch_b200 import _capi, workloads as W
    _capi.require_device(local_rank)
    lib = _capi.lib()
    _capi.set_float32_matmul_precision(args.precision)
    import microtorch_b200.nn as nn
    from microtorch_b200.cuda import kernels as K
    from microtorch_b200.cuda.array import DeviceArray
    from microtorch_b200.nn.activation import Tanh
    from microtorch_b200.nn.loss import CrossEntropyLoss, MSELoss
    from microtorch_b200.tensor import Tensor
    from microtorch_b200.trainer import Communicator, DataParallelTrainer

    base = W.C2 if args.workload == "c2" else W.C5
    wl = W.scaled(base, args.batch)
    # identical replicas (same seed for the model); each rank draws its own shard of synthetic rows
    np.random.seed(wl.seed)
    model = W.build_model(wl, nn.Linear, Tanh, nn.Sequential)
    rng = np.random.default_rng(1000 + rank)
    x_host = rng.standard_normal((args.batch, wl.dims[0]), dtype=np.float32)
    if wl.loss == "ce":
        t_host = np.eye(wl.dims[-1], dtype=np.float32)[rng.integers(0, wl.dims[-1], args.batch)]
    else:
        t_host = rng.uniform(-1, 1, (args.batch, wl.dims[-1])).astype(np.float32)

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
    dp_parity = None
    if world > 1:
        dp_parity = dp_parity_check(comm, dist, rank, world, {"W": W, "nn": nn, "Tanh": Tanh, "Tensor": Tensor})
    loss_fn = CrossEntropyLoss() if wl.loss == "ce" else MSELoss()
    trainer = DataParallelTrainer(model, loss_fn, wl.lr, comm, wl.pred_map, wl.squeeze_dim)

    def barrier():
        _capi.check(lib.mt_sync())
        if dist is not None:
            dist.barrier()

    def max_over_ranks(v):
        if dist is None:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    ev0, ev1 = C.c_void_p(), C.c_void_p()
    lib.mt_event_create(C.byref(ev0))
    lib.mt_event_create(C.byref(ev1))

    def timed(step_fn, steps):
        barrier()
        lib.mt_event_record(ev0)
        for _ in range(steps):
            step_fn()
        lib.mt_event_record(ev1)
        lib.mt_event_sync(ev1)
        barrier()
        ms = C.c_float()
        lib.mt_event_elapsed_ms(ev0, ev1, C.byref(ms))
        return max_over_ranks(ms.value)

    # ------------------------------------------------------------- resident-input arm ("value")
    X, T = Tensor(x_host, device="cuda"), Tensor(t_host, device="cuda")
    last = {}

    def step_resident():
        last["loss"] = trainer.step(X, T)

    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    lib.mt_prof_reset()
    lib.mt_launch_count_reset()
    # N = 1: every tcgen05 launch of the timed region is bracketed by CUDA events (mt_prof_*): the roofline numbers are
    # measured live over the timed region itself.  N > 1: the timed region runs WITHOUT per-kernel events and a second
    # pass of the same steps right after it is profiled instead - timing events at every kernel boundary of the compute
    # stream open gaps in which the concurrent NCCL kernel takes SMs from the next persistent GEMM, which then waits
    # for the peer rank: measured 33.2 vs 30.0 ms per step at N = 2 (profiles/r02_bench_n2_prof_events_ab.txt)
    prof_live = world == 1
    lib.mt_prof_enable(1 if prof_live else 0)
    t_wall0 = time.time()
    ms_total = timed(step_resident, args.steps)
    t_wall1 = time.time()
    launches = C.c_uint64()
    lib.mt_launch_count(C.byref(launches))
    prof_steps = args.steps
    if not prof_live:
        prof_steps = max(3, args.steps // 2)
        lib.mt_prof_reset()
        lib.mt_prof_enable(1)
        timed(step_resident, prof_steps)
    n_gemm, gemm_ms, gemm_flops, gemm_unit_flops = C.c_uint64(), C.c_double(), C.c_double(), C.c_double()
    lib.mt_prof_gemm_units(C.byref(n_gemm), C.byref(gemm_ms), C.byref(gemm_flops), C.byref(gemm_unit_flops))
    lib.mt_prof_enable(0)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    per_rank = None
    if dist is not None:
        # every rank's own tcgen05 time per step: under the power cap the chips of one box do not run at the same
        # clock, and the step of a data-parallel job is paced by the slowest one - that spread is not communication
        import torch
        mine = torch.tensor([gemm_ms.value / prof_steps], dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"gemm_ms_per_step_by_rank": [round(float(a[0]), 3) for a in allr]}
    loss_value = trainer.global_loss(last["loss"])
    samples = args.batch * world * args.steps
    value = samples / (ms_total / 1e3)

    # ------------------------------------------------------------- end-to-end arm (host inputs every step)
    from microtorch_b200 import cuda as mtcuda
    px, pt = mtcuda.pinned_empty(x_host.shape), mtcuda.pinned_empty(t_host.shape)     # pinned host staging
    px[...] = x_host
    pt[...] = t_host
    e2e_loss = []
    pending = {"x": mtcuda.prefetch(px), "t": mtcuda.prefetch(pt)}     # upload of the first batch

    def step_e2e():
        # public API, double-buffered: this step's inputs were uploaded from pinned host memory while the previous
        # step computed; the next step's upload is issued now and overlaps this step.  The loss is read back (D2H)
        # every step.  All of it is inside the timed region.
        mtcuda.wait_prefetch()
        xs, ts = Tensor(pending["x"], device="cuda"), Tensor(pending["t"], device="cuda")
        pending["x"], pending["t"] = mtcuda.prefetch(px), mtcuda.prefetch(pt)
        loss = trainer.step(xs, ts)
        e2e_loss.append(loss.item())

    step_e2e()
    e2e_steps = args.steps          # the same number of steps as `value` (under the power cap 10 vs 20 steps differ)
    ms_e2e = timed(step_e2e, e2e_steps)
    e2e_value = args.batch * world * e2e_steps / (ms_e2e / 1e3)

    # ------------------------------------------------------------- the opt-in fast mode, for information (N = 1)
    fast = None
    if world == 1 and args.precision != "high" and not args.no_fast_mode:
        _capi.set_float32_matmul_precision("high")
        for _ in range(2):
            step_resident()
        fsteps = max(3, args.steps // 2)
        ms_fast = timed(step_resident, fsteps)
        fast = {"matmul_precision": "high", "value": args.batch * fsteps / (ms_fast / 1e3), "unit": "samples/s",
                "ms_per_step": ms_fast / fsteps, "steps": fsteps,
                "note": "every GEMM as TF32 hi.hi + 2 BF16 corrections; config #5 at full width then sits 1.3e-4 from the "
                        "oracle end to end (tests/test_parity_large_gpu.py), which is why it is not the default"}
        _capi.set_float32_matmul_precision(args.precision)

    # ------------------------------------------------------------- the mid-size regime, for information (N = 1)
    # A training step whose GEMMs have 2^29 - 2^31 multiply-adds each (MLP 512 -> 1024 -> 1024 -> 512, 2048 rows) through
    # the same public API: with the dispatcher as shipped (mt_gemm_tc_mid.cu takes them) and with that kernel switched
    # off (round 1's choice between the SIMT tile and the persistent CTA-pair kernel).
    mid = None
    if world == 1 and not args.no_mid:
        from microtorch_b200.nn.optimizer import SGD
        np.random.seed(7)
        mdims, mrows = (512, 1024, 1024, 512), 2048
        mods = []
        for a, b in zip(mdims[:-1], mdims[1:]):
            mods += [nn.Linear(a, b), Tanh()]
        mmodel = nn.Sequential(*mods[:-1])
        mx = Tensor(rng.standard_normal((mrows, mdims[0]), dtype=np.float32), device="cuda")
        mt_ = Tensor(rng.uniform(-1, 1, (mrows, mdims[-1])).astype(np.float32), device="cuda")
        mloss, mopt = MSELoss(), SGD(mmodel.parameters, lr=0.01)
        from microtorch_b200.trainer import train_step

        def step_mid():
            train_step(mmodel, mloss, mopt, mx, mt_, None, None)

        def mid_launches():
            n = C.c_uint64()
            lib.mt_gemm_tc_mid_launches(C.byref(n))
            return n.value

        msteps = 50
        res = {}
        for name, on in (("without_mid_kernel", 0), ("with_mid_kernel", 1)):
            lib.mt_gemm_tc_mid_set(on, 0, 0, 0)
            for _ in range(5):
                step_mid()
            n0 = mid_launches()
            ms = timed(step_mid, msteps)
            res[name] = {"ms_per_step": ms / msteps, "samples_per_s": mrows * msteps / (ms / 1e3),
                         "mid_kernel_launches_per_step": (mid_launches() - n0) / msteps}
        lib.mt_gemm_tc_mid_set(1, 0, 0, 0)
        # the same step as ONE CUDA graph (graph.GraphedStep): what remains when the platform's per-launch cost is gone
        from microtorch_b200.graph import GraphedStep
        gstep = GraphedStep(step_mid, warmup=2)
        for _ in range(5):
            gstep.replay()
        ms = timed(gstep.replay, msteps)
        gstep.close()
        res["with_mid_kernel_graphed"] = {"ms_per_step": ms / msteps, "samples_per_s": mrows * msteps / (ms / 1e3)}
        flops = sum(2.0 * mrows * a * b * (3 if i else 2) for i, (a, b) in enumerate(zip(mdims[:-1], mdims[1:])))
        mid = {"workload": "mlp_512_1024_1024_512_tanh_mse_sgd, 2048 rows (8 GEMMs of 2^30 - 2^31 multiply-adds per step)",
               "steps": msteps, **res,
               "speedup": res["without_mid_kernel"]["ms_per_step"] / res["with_mid_kernel"]["ms_per_step"],
               "step_tflops_logical": flops / (res["with_mid_kernel"]["ms_per_step"] * 1e-3) / 1e12}

    # ------------------------------------------------------------- what data parallelism costs, same process (N > 1)
    dp_overhead = None
    if world > 1:
        # the same step WITHOUT the communicator (independent replicas: every GPU at its own pace) against the
        # data-parallel step, interleaved on the same GPUs: the ratio of the two max-over-ranks times is the cost of the
        # exchange itself, free of the box-to-box and chip-to-chip clock spread that a comparison with a separate
        # N = 1 run carries (tools/dp_ab.py is the long form of this measurement)
        np.random.seed(wl.seed)
        solo = DataParallelTrainer(W.build_model(wl, nn.Linear, Tanh, nn.Sequential),
                                   CrossEntropyLoss() if wl.loss == "ce" else MSELoss(), wl.lr, None, wl.pred_map, wl.squeeze_dim)
        for _ in range(3):
            solo.step(X, T)
        import torch
        t_solo, t_dp, mine_solo = [], [], []
        k = max(3, min(args.steps, 8))
        for _ in range(3):
            barrier()
            lib.mt_event_record(ev0)
            for _ in range(k):
                solo.step(X, T)
            lib.mt_event_record(ev1)
            lib.mt_event_sync(ev1)
            barrier()
            ms = C.c_float()
            lib.mt_event_elapsed_ms(ev0, ev1, C.byref(ms))
            mine_solo.append(ms.value / k)
            t_solo.append(max_over_ranks(ms.value) / k)
            t_dp.append(timed(step_resident, k) / k)
        mine = torch.tensor([statistics.median(mine_solo)], dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        solo_by_rank = [round(float(a[0]), 3) for a in allr]
        dp_overhead = {"replicas_ms_per_step_max_over_ranks": statistics.median(t_solo),
                       "data_parallel_ms_per_step_max_over_ranks": statistics.median(t_dp),
                       "overhead": statistics.median(t_dp) / statistics.median(t_solo) - 1.0,
                       "replicas_ms_per_step_by_rank": solo_by_rank,
                       "chip_spread": max(solo_by_rank) / min(solo_by_rank) - 1.0,
                       "rounds": 3, "steps_per_round": k,
                       "how": "same process, same GPUs, interleaved: independent replicas (no communicator) vs the data-parallel step"}
        solo = None

    # ------------------------------------------------------------- BASELINE config #5 (the named data-parallel config)
    c5 = None
    if args.workload == "c2" and not args.no_c5:
        X = T = px = pt = trainer = model = None      # release config #2 (closures above keep only the names)
        pending.clear()
        last.clear()
        lib.mt_pool_trim()
        wl5 = W.scaled(W.C5, args.batch)
        np.random.seed(wl5.seed)
        model5 = W.build_model(wl5, nn.Linear, Tanh, nn.Sequential)
        rng5 = np.random.default_rng(2000 + rank)
        X5 = Tensor(rng5.standard_normal((args.batch, wl5.dims[0]), dtype=np.float32), device="cuda")
        T5 = Tensor(rng5.uniform(-1, 1, (args.batch, wl5.dims[-1])).astype(np.float32), device="cuda")
        tr5 = DataParallelTrainer(model5, MSELoss(), wl5.lr, comm, wl5.pred_map, wl5.squeeze_dim)
        for _ in range(3):
            last["loss"] = tr5.step(X5, T5)
        c5_steps = max(3, min(args.steps, 5))
        lib.mt_prof_reset()
        lib.mt_prof_enable(1 if prof_live else 0)
        ms5 = timed(lambda: last.__setitem__("loss", tr5.step(X5, T5)), c5_steps)
        if not prof_live:                  # N > 1: per-kernel events in a separate pass (see above)
            lib.mt_prof_enable(1)
            timed(lambda: last.__setitem__("loss", tr5.step(X5, T5)), 2)
        n5, gms5, gfl5 = C.c_uint64(), C.c_double(), C.c_double()
        lib.mt_prof_gemm(C.byref(n5), C.byref(gms5), C.byref(gfl5))
        lib.mt_prof_enable(0)
        c5 = {"workload": "c5_mlp_4096x8_tanh_mse_sgd", "per_gpu_batch": args.batch, "global_batch": args.batch * world,
              "value": args.batch * world * c5_steps / (ms5 / 1e3), "unit": "samples/s", "ms_per_step": ms5 / c5_steps,
              "steps": c5_steps, "warmup": 3, "loss": tr5.global_loss(last["loss"]),
              "gemm_tflops_logical_per_gpu": (gfl5.value / (gms5.value * 1e-3) / 1e12) if gms5.value else None,
              "step_tflops_logical_per_gpu": wl5.gemm_flops_per_step(args.batch) / (ms5 / c5_steps * 1e-3) / 1e12}

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    # fp32-emulated peak: a logical multiply-add costs 4 bf16-rate units on the tensor pipe as TF32 hi.hi (2) + two BF16
    # correction products (backward GEMMs, and every GEMM at precision "high"), 6 units as 3xTF32 (forward GEMMs at
    # the default precision "higher").  EMU is the flop-weighted mean over the GEMM launches of the timed region (the
    # library reports it per launch), so the logical peak is (bf16 dense peak) / EMU
    EMU = gemm_unit_flops.value / gemm_flops.value if gemm_flops.value else 4.0
    peak_tf = peaks["bf16_sustained"] / EMU
    achieved_tf = (gemm_flops.value / n_gemm.value) / ((gemm_ms.value / n_gemm.value) * 1e-3) / 1e12 if n_gemm.value else 0.0
    out = {
        "metric": "mlp_train_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": ("c2_mlp_1024_4096_4096_10_tanh_ce_sgd" if args.workload == "c2" else "c5_mlp_4096x8_tanh_mse_sgd"),
                   "per_gpu_batch": args.batch, "global_batch": args.batch * world, "parallelism": f"dp{world}",
                   "l2": "inputs larger than L2 (activations 1 GiB per layer); no flush needed",
                   "matmul_precision": args.precision,
                   "exchange": (None if comm is None else
                                "NVLink peer memory: reduce-scatter + all-gather + SGD in the library's own kernels (mt_p2p.cu)"
                                if getattr(comm, "p2p", False) else "ncclAllReduce + SGD on the communication stream (mt_comm.cu)"),
                   "gemm": "tcgen05 split-fp32, chunked TMEM accumulation: " +
                           {"high": "TF32 hi.hi + 2 BF16 corrections per k-slice in every GEMM",
                            "higher": "forward GEMMs 3xTF32, backward GEMMs TF32 hi.hi + 2 BF16 corrections",
                            "highest": "3xTF32 in every GEMM"}[args.precision]},
        "loss": loss_value,
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(x_host.nbytes + t_host.nbytes) * world,
                "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps},
        "gpu_launches": int(launches.value),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "gemm_splitfp32_kernel (tcgen05)", "achieved": achieved_tf, "peak": peak_tf,
                     "unit": "TFLOP/s", "frac": achieved_tf / peak_tf if peak_tf else None,
                     "frac_of_burst_peak": achieved_tf / (peaks["bf16"] / EMU) if peaks["bf16"] else None,
                     "traffic": ncu_traffic_per_launch(),
                     "traffic_unit": "bytes per launch (dram read+write, ncu --set full capture under profiles/)",
                     "launches": int(n_gemm.value), "kernel_ms_per_step": gemm_ms.value / prof_steps,
                     "share_of_step": (gemm_ms.value / prof_steps) / (ms_total / args.steps),
                     "measured": ("CUDA events around every tcgen05 launch of the timed region" if prof_live else
                                  f"CUDA events around every tcgen05 launch of a second pass of {prof_steps} steps right after the "
                                  "timed region (N > 1: per-kernel events perturb the compute / NCCL overlap)"),
                     "bf16_units_per_logical_mac": EMU,
                     "peak_source": f"{peaks['source']}: bf16 sustained {peaks['bf16_sustained']} TF/s / {EMU:.3f} - the flop-weighted "
                                    "bf16-rate units per logical multiply-add of this step's GEMMs (TF32 product = 2 units, BF16 "
                                    "product = 1: 3xTF32 forward = 6, TF32 + 2 BF16 backward = 4) - fp32-emulated peak; burst "
                                    "would be %.1f" % (peaks["bf16"] / EMU)},
    }
    if dp_parity is not None:
        out["dp_parity"] = dp_parity
    if dp_overhead is not None:
        out["dp_overhead"] = dp_overhead
    if per_rank is not None:
        out["ranks"] = per_rank
    if c5 is not None:
        out["c5"] = c5
    if fast is not None:
        out["fast_mode"] = fast
    if mid is not None:
        out["mid_regime"] = mid
    if not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline()
    emit(out)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
