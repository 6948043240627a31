#!/usr/bin/env python
"""Accuracy of the split-fp32 tcgen05 GEMM against float64, next to NumPy's own fp32 SGEMM (run under gpurun).

Why: at config #5's full width (8 x 4096-wide layers) the network amplifies rounding differences - the ORACLE
(NumPy fp32) is itself 4e-5 (norm-wise) away from a float64 evaluation of the same step - so how far the device path
may sit from the oracle depends on its per-GEMM error.  Knobs: mix (1 = TF32 + 2 BF16 corrections, 0 = all TF32),
chunk_kb (k-blocks of 32 accumulated in TMEM - with truncation - between folds into IEEE fp32 sums).
One JSON line per setting: per-GEMM errors (max norm-wise, rms, mean signed = bias), time of the big forward GEMM,
and the C5 full-width step: every dW vs float64 and vs the oracle."""
import ctypes as C
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402


def main():
    from microtorch_b200 import _capi, workloads as W
    _capi.require_device(0)
    lib = _capi.lib()
    import microtorch_b200.nn as nn
    from microtorch_b200.cuda.array import DeviceArray
    from microtorch_b200.nn.activation import Tanh
    from microtorch_b200.nn.loss import MSELoss
    from microtorch_b200.nn.optimizer import SGD
    from microtorch_b200.tensor import Tensor
    from microtorch_b200.trainer import train_step
    for f in ("mt_gemm_tc_set_mix", "mt_gemm_tc_set_chunk_kb"):
        getattr(lib, f).argtypes = [C.c_int]
    e0, e1 = C.c_void_p(), C.c_void_p()
    lib.mt_event_create(C.byref(e0))
    lib.mt_event_create(C.byref(e1))
    rng = np.random.default_rng(0)
    M = N = K = 4096
    X = np.tanh(rng.standard_normal((M, K))).astype(np.float32)
    Wm = (rng.uniform(-1, 1, (N, K)) * 0.1).astype(np.float32)
    ref = X.astype(np.float64) @ Wm.astype(np.float64).T
    scale = np.abs(ref).max()
    np32 = X @ Wm.T

    def stats(got):
        d = got.astype(np.float64) - ref
        return {"max_normwise": float(np.abs(d).max() / scale), "rms_rel": float(np.sqrt((d ** 2).mean()) / np.sqrt((ref ** 2).mean())),
                "mean_signed_rel": float((d * np.sign(ref)).mean() / np.abs(ref).mean())}
    print(json.dumps({"impl": "numpy fp32 sgemm (OpenBLAS)", **stats(np32)}), flush=True)

    # C5 full width, 4096 rows: float64 truth and the oracle (host), once
    from oracle import workloads as OW
    wl = W.scaled(W.C5, 4096)
    t0 = time.time()
    oref = OW.run_steps(wl, 1)
    omodel, ox, ot = OW.setup(wl)
    params = [p.data.astype(np.float64) for p in omodel.parameters()]
    Ws, bs = params[0::2], params[1::2]
    h = [ox.data.astype(np.float64)]
    for Wl, b in zip(Ws, bs):
        h.append(np.tanh(h[-1] @ Wl.T + b))
    g = 2 * (h[-1] - ot.data.astype(np.float64)) / h[-1].size
    g64 = [None] * len(Ws)
    for i in reversed(range(len(Ws))):
        dz = g * (1 - h[i + 1] ** 2)
        g64[i] = dz.T @ h[i]
        g = dz @ Ws[i]
    nw = lambda a, b: float(np.abs(a.astype(np.float64) - b).max() / np.abs(b).max())
    print(json.dumps({"c5_full_width": "oracle (NumPy fp32) vs float64", "host_seconds": round(time.time() - t0, 1),
                      "dW_err": [nw(oref["grads"][2 * i], g64[i]) for i in range(len(Ws))]}), flush=True)

    Xd, Wd, Y = DeviceArray.from_host(X), DeviceArray.from_host(Wm), DeviceArray.empty((M, N))
    lib.mt_gemm_tc_set_fwd.argtypes = [C.c_int, C.c_int]
    # config #2 at full size for the cost of each setting (whole training step, interleaved over two rounds)
    from microtorch_b200.nn.loss import CrossEntropyLoss
    wl2 = W.C2
    m2, x2, t2 = W.setup(wl2, nn.Linear, Tanh, nn.Sequential)
    X2, T2 = Tensor(x2, device="cuda"), Tensor(t2, device="cuda")
    l2, o2 = CrossEntropyLoss(), SGD(m2.parameters, lr=wl2.lr)
    del x2, t2
    variants = [(1, 8, -1, -1), (1, 4, -1, -1), (1, 2, -1, -1), (1, 8, 0, 2), (1, 8, 0, 4), (1, 8, 1, 2), (1, 4, 0, 4),
                (0, 8, -1, -1), (0, 4, -1, -1), (0, 2, -1, -1)]
    recs = []
    for mix, chunk, fmix, fchunk in variants:
        lib.mt_gemm_tc_set_mix(mix)
        lib.mt_gemm_tc_set_chunk_kb(chunk)
        lib.mt_gemm_tc_set_fwd(fmix, fchunk)
        lib.mt_gemm_set_engine(1)
        _capi.check(lib.mt_linear_fwd_f32(Y.ptr, Xd.ptr, Wd.ptr, None, M, N, K, 0))
        rec = {"impl": "tcgen05 split-fp32", "mix": mix, "chunk_kb": chunk, "fwd_mix": fmix, "fwd_chunk_kb": fchunk,
               "fwd_gemm": stats(Y.get())}
        lib.mt_gemm_set_engine(-1)
        model, x, t = W.setup(wl, nn.Linear, Tanh, nn.Sequential)
        train_step(model, MSELoss(), SGD(model.parameters, lr=wl.lr), Tensor(x, device="cuda"), Tensor(t, device="cuda"))
        grads = [p.grad.get() for p in list(model.parameters())[0::2]]
        rec["c5_dW_vs_f64_max"] = max(nw(gd, g64[i]) for i, gd in enumerate(grads))
        rec["c5_dW_vs_oracle"] = [nw(gd, oref["grads"][2 * i].astype(np.float64)) for i, gd in enumerate(grads)]
        rec["c5_dW_vs_oracle_max"] = max(rec["c5_dW_vs_oracle"])
        rec["c2_step_ms"] = []
        recs.append(rec)
    for rnd in range(3):
        for rec in recs:
            lib.mt_gemm_tc_set_mix(rec["mix"])
            lib.mt_gemm_tc_set_chunk_kb(rec["chunk_kb"])
            lib.mt_gemm_tc_set_fwd(rec["fwd_mix"], rec["fwd_chunk_kb"])
            train_step(m2, l2, o2, X2, T2, wl2.pred_map)
            lib.mt_event_record(e0)
            for _ in range(3):
                train_step(m2, l2, o2, X2, T2, wl2.pred_map)
            lib.mt_event_record(e1)
            lib.mt_event_sync(e1)
            ms = C.c_float()
            lib.mt_event_elapsed_ms(e0, e1, C.byref(ms))
            if rnd:
                rec["c2_step_ms"].append(round(ms.value / 3, 3))
    for rec in recs:
        print(json.dumps(rec), flush=True)
    lib.mt_gemm_tc_set_fwd(-1, -1)
    lib.mt_gemm_tc_set_mix(1)
    lib.mt_gemm_tc_set_chunk_kb(8)
    lib.mt_gemm_set_engine(-1)


if __name__ == "__main__":
    main()
