#!/usr/bin/env python
"""The HBM-bound kernels of the training step, one by one, at 2^26 and 2^28 elements (run under gpurun).

For VERDICT r1 item 3: achieved GB/s (ALGORITHMIC bytes of SURVEY.md 8(d) / CUDA-event time, L2 flushed between
launches - mt_l2_flush leaves it cold and clean) of bin_flat / bin_2d / unary / reduce_rows / reduce_cols / loss_kernel / sgd_multi_kernel / zero_multi and
the skinny head GEMMs, each row with the nvidia-smi clocks and throttle reasons sampled while it ran.

    python tools/hbm_kernels.py                # JSON lines (timings + clocks)
    python tools/hbm_kernels.py --once         # one launch per kernel: the run to put under
                                               #   ncu --set full -k regex:'bin_|unary|reduce_|loss_|sgd_|zero_|skinny_'
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="26,28")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--only", default="")
    ap.add_argument("--two-launch", action="store_true", help="split reductions as main pass + fold launch (round 1's form)")
    args = ap.parse_args()
    from bench import ClockSampler, load_peaks
    from microtorch_b200 import _capi
    _capi.require_device(0)
    lib = _capi.lib()
    if args.two_launch:
        lib.mt_reduce_set_single_launch(0)
    from microtorch_b200.cuda import kernels as K
    from microtorch_b200.cuda.array import DeviceArray
    peak = load_peaks()["hbm_gbs"]
    e0, e1 = C.c_void_p(), C.c_void_p()
    lib.mt_event_create(C.byref(e0))
    lib.mt_event_create(C.byref(e1))
    sampler = None
    if not args.once:
        sampler = ClockSampler(0, period_ms=50)
        sampler.start()
        time.sleep(0.3)

    def timed(fn, iters):
        for _ in range(3):
            fn()
        tot = 0.0
        t0 = time.time()
        reps = 0
        while reps < iters or time.time() - t0 < 0.35:        # long enough for the clock sampler to see it
            lib.mt_l2_flush()
            lib.mt_event_record(e0)
            fn()
            lib.mt_event_record(e1)
            lib.mt_event_sync(e1)
            ms = C.c_float()
            lib.mt_event_elapsed_ms(e0, e1, C.byref(ms))
            tot += ms.value
            reps += 1
        return tot / reps, t0, time.time()

    rng = np.random.default_rng(0)
    for p in (int(v) for v in args.sizes.split(",")):
        n = 1 << p
        R, Cc = n // 4096, 4096
        B = 4.0 * n
        x = DeviceArray.from_host(rng.uniform(0.5, 1.5, (R, Cc)).astype(np.float32))
        y = DeviceArray.from_host(rng.uniform(0.5, 1.5, (R, Cc)).astype(np.float32))
        out = DeviceArray.empty((R, Cc))
        row = DeviceArray.from_host(rng.uniform(0.5, 1.5, (1, Cc)).astype(np.float32))
        col = DeviceArray.from_host(rng.uniform(0.5, 1.5, (R, 1)).astype(np.float32))
        prob = DeviceArray.from_host(rng.uniform(0.02, 0.98, (R, Cc)).astype(np.float32))     # BCE needs p in (0, 1)
        tgt = DeviceArray.from_host(rng.uniform(0.0, 1.0, (R, Cc)).astype(np.float32))
        work = {
            "bin_flat add (x+y)": (lambda: K.binary(_capi.BIN_ADD, x, y, out=out), 3 * B),
            "bin_flat tanh' (g*(1-y^2))": (lambda: K.binary(_capi.BIN_TANH_BWD, x, y, out=out), 3 * B),
            "bin_2d mul row-broadcast": (lambda: K.binary(_capi.BIN_MUL, x, row, out=out), 2 * B),
            "bin_2d mul col-broadcast": (lambda: K.binary(_capi.BIN_MUL, x, col, out=out), 2 * B),
            "unary tanh": (lambda: K.unary(_capi.UN_TANH, x), 2 * B),
            "unary log": (lambda: K.unary(_capi.UN_LOG, x), 2 * B),
            "reduce_rows sum(axis=1)": (lambda: K.sum(x, 1), B),
            "reduce_rows sum(all)": (lambda: K.sum(x), B),
            "reduce_rows fused g*x -> col": (lambda: K.reduce_sum(x, (1,), y), 2 * B),
            "reduce_cols sum(axis=0)": (lambda: K.sum(x, 0), B),
            "reduce_cols fused g*x -> row": (lambda: K.reduce_sum(x, (0,), y), 2 * B),
            "loss_kernel mse fwd+bwd": (lambda: K.loss_forward_backward(_capi.LOSS_MSE, x, y, n, True), 3 * B),
            "loss_kernel bce fwd+bwd": (lambda: K.loss_forward_backward(_capi.LOSS_BCE, prob, tgt, n, True), 3 * B),
            "sgd_multi_kernel": (lambda: K.sgd_multi([x], [y], 1e-9), 3 * B),
            "zero_multi_kernel": (lambda: K.zero_multi([out]), B),
        }
        # tall-thin un-broadcast shapes (reduce_cols at 17-54 % with bank conflicts) and the short-fat one
        for rr, cc in ((n // 8, 8), (n // 32, 32), (n // 512, 512), (2, n // 2)):
            tt = x.reshape((rr, cc))
            work[f"reduce_cols tall-thin ({rr},{cc}) sum(axis=0)"] = (lambda tt=tt: K.sum(tt, 0), B + 4.0 * cc)
        if n >= 65536 * 4096:
            # the 4096 -> 10 head of config #2 at 65536 rows: skinny forward / dX / dW (1 GiB operand streamed once)
            M, Kd, N = 65536, 4096, 10
            hx = x.reshape((-1,))[:M * Kd].reshape((M, Kd))
            hw = DeviceArray.from_host((rng.standard_normal((N, Kd)) * 0.05).astype(np.float32))
            hb = DeviceArray.from_host(np.zeros(N, np.float32))
            hg = DeviceArray.from_host(rng.standard_normal((M, N)).astype(np.float32))
            head = 4.0 * M * Kd
            work["skinny_fwd 65536x10x4096"] = (lambda: K.linear_forward(hx, hw, hb, _capi.ACT_TANH), head + 4.0 * M * N)
            work["skinny_dw+dx 65536x10x4096"] = (lambda: K.linear_backward(hg, None, hx, hw, True), 2 * head + 4.0 * M * N * 2)
        for name, (fn, nbytes) in work.items():
            if args.only and not any(k in name for k in args.only.split(",")):
                continue
            if args.once:
                fn()
                _capi.check(lib.mt_sync())
                print(json.dumps({"kernel": name, "log2_n": p, "alg_gb": nbytes / 1e9}), flush=True)
                continue
            ms, t0, t1 = timed(fn, args.iters)
            clk = sampler.window(t0, t1)
            print(json.dumps({"kernel": name, "log2_n": p, "ms": ms, "alg_gb": nbytes / 1e9, "gbs": nbytes / ms / 1e6,
                              "frac_of_measured_hbm": nbytes / ms / 1e6 / peak, "clocks": clk}), flush=True)
        del work, x, y, out, prob, tgt
    if sampler is not None:
        sampler.stop(0, time.time())


if __name__ == "__main__":
    main()
