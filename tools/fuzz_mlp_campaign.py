#!/usr/bin/env python
"""One-off: more seeds of the random MLP training-step comparison of tests/test_fuzz_mlp_gpu.py (CUDA path vs the oracle;
run under gpurun from the repo root).  Prints one JSON line: how many agree, and the findings."""
import sys, os, json
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
from test_fuzz_mlp_gpu import random_workload, host, close
from microtorch_b200 import _capi, workloads as W
_capi.require_device()
import microtorch_b200.nn as nn
from microtorch_b200.nn.activation import Tanh
from microtorch_b200.nn.loss import BCELoss, CrossEntropyLoss, MSELoss
from microtorch_b200.nn.optimizer import SGD
from microtorch_b200.tensor import Tensor
from microtorch_b200.trainer import train_step
from oracle import workloads as OW
losses = {"mse": MSELoss, "ce": CrossEntropyLoss, "bce": BCELoss}
ok, bad = 0, []
for seed in range(40, 160):
    wl = random_workload(seed)
    try:
        ref = OW.run_steps(wl, 2)
        model, x, t = W.setup(wl, nn.Linear, Tanh, nn.Sequential)
        X, T = Tensor(x, device="cuda"), Tensor(t, device="cuda")
        lf, opt = losses[wl.loss](), SGD(model.parameters, lr=wl.lr)
        ls = [host(train_step(model, lf, opt, X, T, wl.pred_map).data) for _ in range(2)]
        np.testing.assert_allclose(np.array(ls, np.float32), ref["losses"], rtol=1e-4)
        for i, p in enumerate(model.parameters()):
            close(p.data, ref["params"][i], f"param {i}")
            if np.abs(ref["grads"][i]).max() > 0: close(p.grad, ref["grads"][i], f"grad {i}")
        ok += 1
    except Exception as e:
        bad.append({"seed": seed, "workload": str(wl), "error": f"{type(e).__name__}: {str(e)[:200]}"})
print(json.dumps({"seeds": "40..159", "agree": ok, "findings": bad}))
