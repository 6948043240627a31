#!/usr/bin/env python
"""One launch of the single-launch trainer on the notebook model (for ncu captures): python tools/fit_once.py [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    from microtorch_b200 import _capi, workloads as W
    _capi.require_device(0)
    import microtorch_b200.nn as nn
    from microtorch_b200.nn.activation import Tanh
    from microtorch_b200.nn.loss import MSELoss
    from microtorch_b200.nn.optimizer import SGD
    from microtorch_b200.tensor import Tensor
    from microtorch_b200.trainer import fit
    wl = W.C1A
    model, x, t = W.setup(wl, nn.Linear, Tanh, nn.Sequential)
    X, T = Tensor(x, device="cuda"), Tensor(t, device="cuda")
    losses = fit(model, MSELoss(), SGD(model.parameters, lr=wl.lr), X, T, steps, wl.pred_map, wl.squeeze_dim, fused=True)
    print(f"{steps} steps in one launch: loss {losses[0]:.6f} -> {losses[-1]:.6f}")


if __name__ == "__main__":
    main()
