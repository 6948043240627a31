#!/usr/bin/env python
"""CPU timings of the BASELINE.json configs on the oracle (the NumPy port of the reference), bounded samples.
Lives under tests/: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may execute the oracle.
    python tests/cpu_config_baselines.py            # one JSON line per config"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def cpu(wl, steps, tag):
    from oracle import workloads as OW
    from oracle.nnref import train_step
    model, x, t = OW.setup(wl)
    train_step(model, x, t, wl.loss, wl.lr, wl.pred_map, wl.squeeze_dim, x.ndim == 3)
    t0 = time.perf_counter()
    for _ in range(steps):
        train_step(model, x, t, wl.loss, wl.lr, wl.pred_map, wl.squeeze_dim, x.ndim == 3)
    dt = (time.perf_counter() - t0) / steps
    print(json.dumps({"config": tag + "_cpu_oracle", "ms_per_step": dt * 1e3, "steps_per_s": 1 / dt,
                      "samples_per_s": wl.batch / dt, "host_cpus": os.cpu_count()}), flush=True)


if __name__ == "__main__":
    from microtorch_b200 import workloads as W
    cpu(W.C1A, 500, "c1a")
    cpu(W.C1B, 500, "c1b")
    cpu(W.scaled(W.C2, 4096), 2, "c2_at_4096_rows")
    cpu(W.Workload("c4s", (2048,) * 3, "bce", 0.01, 8, pred_map=(0.49, 0.5), seed=0, input_shape=(8, 512, 2048)), 1,
        "c4_at_8x512x2048")
    cpu(W.scaled(W.C5, 2048), 1, "c5_at_2048_rows")
