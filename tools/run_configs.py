#!/usr/bin/env python
"""Measure the BASELINE.json configs other than the bench headline through the public API (run under gpurun):

  c1a / c1b   spiral demo on device="cuda": steps/s (latency-bound), loss parity with the notebook stream
  c4          3-D Linear stack (256 x 512 x 2048, k layers) + BCELoss: samples/s, GEMM TFLOP/s
  c5          4096-wide x 8 layers, 65536 rows per GPU (the per-GPU share of global batch 524288 at 8 GPUs)

One JSON line per config.  (The CPU-port timings of the same configs come from tests/cpu_config_baselines.py:
only tests/, smoke() and bench.py's CPU legs may touch the oracle.)"""
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
    ap.add_argument("--only", default="")
    ap.add_argument("--c4-layers", type=int, default=2)
    args = ap.parse_args()
    from microtorch_b200 import _capi, workloads as W
    _capi.require_device(0)
    lib = _capi.lib()
    import microtorch_b200.nn as nn
    from microtorch_b200.nn.activation import Tanh
    from microtorch_b200.nn.loss import BCELoss, CrossEntropyLoss, MSELoss
    from microtorch_b200.nn.optimizer import SGD
    from microtorch_b200.tensor import Tensor
    from microtorch_b200.trainer import train_step
    losses = {"mse": MSELoss, "ce": CrossEntropyLoss, "bce": BCELoss}
    e0, e1 = C.c_void_p(), C.c_void_p()
    lib.mt_event_create(C.byref(e0))
    lib.mt_event_create(C.byref(e1))

    def run(wl, steps, warmup, tag, extra=None):
        model, x, t = W.setup(wl, nn.Linear, Tanh, nn.Sequential)
        X, T = Tensor(x, device="cuda"), Tensor(t, device="cuda")
        loss_fn, opt = losses[wl.loss](), SGD(model.parameters, lr=wl.lr)
        first = None
        for _ in range(warmup):
            l = train_step(model, loss_fn, opt, X, T, wl.pred_map, wl.squeeze_dim)
            if first is None:
                first = l.item()
        lib.mt_sync()
        lib.mt_prof_reset()
        lib.mt_prof_enable(1)
        lib.mt_launch_count_reset()
        lib.mt_event_record(e0)
        for _ in range(steps):
            l = train_step(model, loss_fn, opt, X, T, wl.pred_map, wl.squeeze_dim)
        lib.mt_event_record(e1)
        lib.mt_event_sync(e1)
        ms = C.c_float()
        lib.mt_event_elapsed_ms(e0, e1, C.byref(ms))
        n, gms, gfl, nl = C.c_uint64(), C.c_double(), C.c_double(), C.c_uint64()
        lib.mt_prof_gemm(C.byref(n), C.byref(gms), C.byref(gfl))
        lib.mt_prof_enable(0)
        lib.mt_launch_count(C.byref(nl))
        rows = int(np.prod(x.shape[:-1]))
        rec = {"config": tag, "steps": steps, "ms_per_step": ms.value / steps, "steps_per_s": steps / (ms.value / 1e3),
               "samples_per_s": wl.batch * steps / (ms.value / 1e3), "rows_per_step": rows, "first_loss": first,
               "last_loss": l.item(), "launches_per_step": nl.value / steps,
               "tcgen05_gemm_ms_per_step": gms.value / steps,
               "tcgen05_gemm_tflops": (gfl.value / gms.value / 1e9) if gms.value else None}
        if extra:
            rec.update(extra)
        print(json.dumps(rec), flush=True)
        del model, X, T
        lib.mt_pool_trim()

    def run_graphed(wl, steps, tag):
        from microtorch_b200.graph import GraphedStep
        model, x, t = W.setup(wl, nn.Linear, Tanh, nn.Sequential)
        X, T = Tensor(x, device="cuda"), Tensor(t, device="cuda")
        loss_fn, opt = losses[wl.loss](), SGD(model.parameters, lr=wl.lr)
        g = GraphedStep(lambda: train_step(model, loss_fn, opt, X, T, wl.pred_map, wl.squeeze_dim), warmup=3)
        for _ in range(20):
            g.replay()
        lib.mt_sync()
        lib.mt_event_record(e0)
        for _ in range(steps):
            g.replay()
        lib.mt_event_record(e1)
        lib.mt_event_sync(e1)
        ms = C.c_float()
        lib.mt_event_elapsed_ms(e0, e1, C.byref(ms))
        print(json.dumps({"config": tag + "_cuda_graph", "steps": steps, "ms_per_step": ms.value / steps,
                          "steps_per_s": steps / (ms.value / 1e3), "samples_per_s": wl.batch * steps / (ms.value / 1e3),
                          "last_loss": g.result.item()}), flush=True)
        g.close()

    def run_single_launch(wl, steps, tag):
        from microtorch_b200.trainer import fit
        model, x, t = W.setup(wl, nn.Linear, Tanh, nn.Sequential)
        X, T = Tensor(x, device="cuda"), Tensor(t, device="cuda")
        loss_fn, opt = losses[wl.loss](), SGD(model.parameters, lr=wl.lr)
        if os.environ.get("MT_MLP_THREADS"):                 # tuning knobs of the on-chip trainer
            _capi.check(lib.mt_mlp_set_threads(int(os.environ["MT_MLP_THREADS"])))
        if os.environ.get("MT_MLP_CLUSTER"):
            _capi.check(lib.mt_mlp_set_cluster(int(os.environ["MT_MLP_CLUSTER"])))
        fit(model, loss_fn, opt, X, T, 50, wl.pred_map, wl.squeeze_dim, fused=True)       # warm-up (module load, migration)
        lib.mt_sync()
        lib.mt_event_record(e0)
        ls = fit(model, loss_fn, opt, X, T, steps, wl.pred_map, wl.squeeze_dim, fused=True)
        lib.mt_event_record(e1)
        lib.mt_event_sync(e1)
        ms = C.c_float()
        lib.mt_event_elapsed_ms(e0, e1, C.byref(ms))
        clk = (C.c_ulonglong * 8)()
        lib.mt_mlp_debug_clocks(1, clk)                     # phase clocks of rank 0 for a second, instrumented run
        fit(model, loss_fn, opt, X, T, steps, wl.pred_map, wl.squeeze_dim, fused=True)
        lib.mt_mlp_debug_clocks(0, clk)
        phases = dict(zip(["fwd", "loss", "dx", "dw", "cluster_barrier", "exchange_update"], [c / steps for c in clk[:6]]))
        print(json.dumps({"config": tag + "_single_launch_loop", "steps": steps, "sm_cycles_per_step_by_phase": phases, "ms_per_step": ms.value / steps,
                          "us_per_step": 1e3 * ms.value / steps, "steps_per_s": steps / (ms.value / 1e3),
                          "samples_per_s": wl.batch * steps / (ms.value / 1e3), "launches": 1,
                          "last_loss": float(ls[-1])}), flush=True)

    want = lambda k: (not args.only and k not in ("c1fit", "mid")) or k in args.only.split(",")
    if want("c1fit"):
        run_single_launch(W.C1A, 5000, "c1a_spiral_notebook_mlp_2_64_32_16_1")
        run_single_launch(W.C1B, 5000, "c1b_spiral_2_100_3")
    if want("c1a"):
        run(W.C1A, 2000, 50, "c1a_spiral_notebook_mlp_2_64_32_16_1")
        run_graphed(W.C1A, 5000, "c1a_spiral_notebook_mlp_2_64_32_16_1")
        run_single_launch(W.C1A, 5000, "c1a_spiral_notebook_mlp_2_64_32_16_1")
    if want("c1b"):
        run(W.C1B, 2000, 50, "c1b_spiral_2_100_3")
        run_graphed(W.C1B, 5000, "c1b_spiral_2_100_3")
        run_single_launch(W.C1B, 5000, "c1b_spiral_2_100_3")
    if want("mid"):      # between the on-chip trainer and the big persistent GEMM: a few thousand rows, a few hundred features
        for batch, dims in ((4096, (256, 512, 512, 10)), (8192, (784, 1024, 1024, 10)), (1024, (128, 256, 256, 10))):
            wl = W.scaled(W.C2, batch, dims)
            tag = "mid_mlp_" + "_".join(map(str, dims)) + f"_b{batch}"
            run(wl, 200, 20, tag, {"gemm_flops_per_step": wl.gemm_flops_per_step()})
            run_graphed(wl, 500, tag)
    if want("c4"):
        k = args.c4_layers
        wl = W.Workload("c4", (2048,) * (k + 1), "bce", 0.01, 256, pred_map=(0.49, 0.5), seed=0, input_shape=(256, 512, 2048))
        run(wl, 5, 2, f"c4_3d_linear_256x512x2048_k{k}_bce", {"gemm_flops_per_step": wl.gemm_flops_per_step()})
    if want("c5"):
        wl = W.scaled(W.C5, 65536)
        run(wl, 4, 2, "c5_mlp_4096x8_rows65536_per_gpu", {"gemm_flops_per_step": wl.gemm_flops_per_step()})


if __name__ == "__main__":
    main()
