#!/usr/bin/env python
"""One FULL-SIZE training step of BASELINE.json config #2 (B = 65536, 1024 -> 4096 -> 4096 -> 10, CE, SGD) on the
device against the oracle on the host ("full B=65536 ... run once").  Too slow and too large for
the pytest suite (~20-40 s of NumPy, ~25 GiB of host memory), so it is a script: run under gpurun, one JSON line with
the norm-wise errors of the loss, every weight gradient and every updated weight.  Falls back to a smaller batch when
the host does not have the memory, and says so.

    python tests/full_size_parity.py > gpurun_out/full_size_parity.json
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def run():
    import psutil
    from microtorch_b200 import _capi, workloads as W
    from oracle import workloads as OW
    _capi.require_device(0)
    import microtorch_b200.nn as nn
    from microtorch_b200.nn.activation import Tanh
    from microtorch_b200.nn.loss import CrossEntropyLoss
    from microtorch_b200.nn.optimizer import SGD
    from microtorch_b200.tensor import Tensor
    from microtorch_b200.trainer import train_step

    avail = psutil.virtual_memory().available / 2 ** 30
    batch = 65536 if avail >= 48 else (16384 if avail >= 16 else 4096)
    wl = W.scaled(W.C2, batch)

    t0 = time.perf_counter()
    ref = OW.run_steps(wl, 1)
    cpu_s = time.perf_counter() - t0

    model, x, t = W.setup(wl, nn.Linear, Tanh, nn.Sequential)
    X, T = Tensor(x, device="cuda"), Tensor(t, device="cuda")
    loss_fn, opt = CrossEntropyLoss(), SGD(model.parameters, lr=wl.lr)
    loss = train_step(model, loss_fn, opt, X, T, wl.pred_map, wl.squeeze_dim)
    got_loss = float(loss.item())
    params = list(model.parameters())

    def err(a, b):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))

    rec = {"config": "c2_full_size_one_step", "batch": batch, "host_mem_available_gib": round(avail, 1),
           "oracle_cpu_seconds": round(cpu_s, 2), "host_cpus": os.cpu_count(),
           "loss_device": got_loss, "loss_oracle": float(ref["losses"][0]),
           "loss_rel_err": abs(got_loss - float(ref["losses"][0])) / abs(float(ref["losses"][0])),
           "grad_normwise_err": [err(p.grad.get(), g) for p, g in zip(params, ref["grads"])][0::2],
           "weight_after_sgd_normwise_err": [err(p.data.get(), w) for p, w in zip(params, ref["params"])][0::2],
           "bias_unchanged": all(np.array_equal(p.data.get(), w) for p, w in list(zip(params, ref["params"]))[1::2]),
           "tolerance": 1e-4}
    rec["pass"] = bool(rec["loss_rel_err"] <= 1e-4 and max(rec["grad_normwise_err"]) <= 1e-4
                       and max(rec["weight_after_sgd_normwise_err"]) <= 1e-4 and rec["bias_unchanged"])
    return rec


if __name__ == "__main__":
    result = run()
    print(json.dumps(result), flush=True)
    sys.exit(0 if result["pass"] else 1)
