#!/usr/bin/env python
"""BASELINE.json config #3: broadcasting elementwise / reduction sweep through the public API,
achieved HBM GB/s of each workload against the measured copy bandwidth (run under gpurun).

Workloads, x,y ~ U(0.5,1.5) fp32, N in 2^20..2^28 (2^30 with --big):
  poly   z = x*y + x**2 - y ; z.sum().backward()      y same-shape / (1,C) / (R,1) / scalar
  log    x.log().sum().backward()      tanh   x.tanh().sum().backward()
  sum    x.sum()      mean   x.mean() + backward      sum0/sum1  x.sum(axis=0/1)
Bytes are ALGORITHMIC: the sum of the per-kernel minimum traffic of the launches
the workload needs.  One JSON line per (workload, N)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min", type=int, default=20)
    ap.add_argument("--max", type=int, default=28)
    ap.add_argument("--step", type=int, default=2)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", default="")
    ap.add_argument("--graph", action="store_true", help="also time CUDA-graph replays of the tape workloads")
    ap.add_argument("--graph-max", type=int, default=24, help="largest log2(N) to capture as a graph")
    args = ap.parse_args()
    from microtorch_b200 import _capi
    _capi.require_device(0)
    lib = _capi.lib()
    from microtorch_b200.tensor import Tensor
    peak = 6543.7
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    e0, e1 = C.c_void_p(), C.c_void_p()
    lib.mt_event_create(C.byref(e0))
    lib.mt_event_create(C.byref(e1))
    from bench import ClockSampler            # nvidia-smi clocks / throttle reasons sampled while each row runs
    sampler = ClockSampler(0, period_ms=50)
    sampler.start()
    time.sleep(0.3)

    def timed(fn, iters):
        for _ in range(3):
            fn()
        tot, reps, t0 = 0.0, 0, time.time()
        while reps < iters or time.time() - t0 < 0.3:          # long enough for the clock sampler to see the row
            lib.mt_l2_flush()
            lib.mt_event_record(e0)
            fn()
            lib.mt_event_record(e1)
            lib.mt_event_sync(e1)
            ms = C.c_float()
            lib.mt_event_elapsed_ms(e0, e1, C.byref(ms))
            tot += ms.value
            reps += 1
        return tot / reps, sampler.window(t0, time.time())

    rng = np.random.default_rng(0)
    for p in range(args.min, args.max + 1, args.step):
        n = 1 << p
        R, Cc = n // 4096, 4096
        xh = rng.uniform(0.5, 1.5, (R, Cc)).astype(np.float32)
        x = Tensor(xh, requires_grad=True, device="cuda")
        ys = {"same": rng.uniform(0.5, 1.5, (R, Cc)).astype(np.float32), "row": rng.uniform(0.5, 1.5, (1, Cc)).astype(np.float32),
              "col": rng.uniform(0.5, 1.5, (R, 1)).astype(np.float32), "scalar": np.float32(1.25).reshape(())}
        B = 4.0 * n
        work = {}
        graphs = []
        for tag, yh in ys.items():
            y = Tensor(yh, requires_grad=True, device="cuda")
            m = 4.0 * yh.size

            def poly(y=y):
                x.zero_grad(); y.zero_grad()
                z = x * y + x ** 2 - y
                z.sum().backward()
            # fwd: mul (2B+m), pow (2B), add (3B), neg (2m), add (2B+m), sum (B)
            # bwd (g = dL/dz is a stride-0 view of one scalar, never materialised): pow' g*2x (2B); x.grad += g*y
            #      (read y m, write B; accumulate 3B); y.grad: same-shape g*x (2B) + accumulate (2B), broadcast y:
            #      one read of x (B) folded to m elements
            bytes_alg = ((2 * B + m) + 2 * B + 3 * B + 2 * m + (2 * B + m) + B) + 2 * B + (m + B) + 3 * B \
                + (4 * B if tag == "same" else B + m)
            work[f"poly_{tag}"] = (poly, bytes_alg)

        def f_log():
            x.zero_grad()
            x.log().sum().backward()

        def f_tanh():
            x.zero_grad()
            x.tanh().sum().backward()
        work["log_fwd_bwd"] = (f_log, 2 * B + B + 2 * B)          # log (8N) + sum (4N) + g/x (8N: g is a broadcast scalar)
        work["tanh_fwd_bwd"] = (f_tanh, 2 * B + B + 2 * B)
        work["sum_fwd"] = (lambda: x.sum(), B)

        def f_mean():
            x.zero_grad()
            x.mean().backward()
        work["mean_fwd_bwd"] = (f_mean, B)                           # backward is a stride-0 view: no bytes until read
        work["sum_axis0"] = (lambda: x.sum(axis=0), B)
        work["sum_axis1"] = (lambda: x.sum(axis=1), B)
        xd = x.data
        yd = Tensor(ys["same"], device="cuda").data
        from microtorch_b200.cuda import kernels as K
        work["k_add_same"] = (lambda: K.binary(_capi.BIN_ADD, xd, yd), 3 * B)          # raw kernels, distinct operands
        work["k_tanh"] = (lambda: K.unary(_capi.UN_TANH, xd), 2 * B)
        work["k_tanh_bwd"] = (lambda: K.binary(_capi.BIN_TANH_BWD, xd, yd), 3 * B)
        work["k_mul_reduce_cols"] = (lambda: K.reduce_sum(xd, (0,), yd), 2 * B)         # fused grad*other + unbroadcast
        from microtorch_b200.nn.loss import MSELoss
        lossf = MSELoss()
        xp = Tensor(xd, requires_grad=True, device="cuda")
        yt = Tensor(yd, device="cuda")

        def f_loss():
            xp.zero_grad()
            lossf(xp, yt).backward()
        work["mse_loss_fwd_bwd"] = (f_loss, 3 * B)
        # launch-bound sizes: the same workload captured once and replayed as ONE CUDA graph (graph.GraphedStep)
        if args.graph and p <= args.graph_max:
            from microtorch_b200.graph import GraphedStep
            for name in [k for k in list(work) if k.startswith(("poly_", "log_", "tanh_", "mean_", "mse_"))]:
                fn, nbytes = work[name]
                gs = GraphedStep(fn, warmup=2)
                graphs.append(gs)
                work[name + "_cuda_graph"] = (gs.replay, nbytes)
        for name, (fn, nbytes) in work.items():
            if args.only and not any(k in name for k in args.only.split(',')):
                continue
            ms, clk = timed(fn, args.iters)
            rec = {"workload": name, "log2_n": p, "ms": ms, "alg_gb": nbytes / 1e9, "gbs": nbytes / ms / 1e6,
                   "frac_of_measured_hbm": nbytes / ms / 1e6 / peak, "clocks": clk}
            print(json.dumps(rec), flush=True)
        for gs in graphs:
            gs.close()
        del x, work
    sampler.stop(0, time.time())


if __name__ == "__main__":
    main()
