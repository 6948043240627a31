"""Oracle runs of the BASELINE.json workloads.  TEST INFRASTRUCTURE ONLY.

Builds the oracle model for a `microtorch_b200.workloads.Workload` (data generation is
shared with the product so both see the same bytes; no product arithmetic is imported)
and steps it the way the reference's notebook loop does (demo.ipynb:5178-5189).
"""

from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np

from microtorch_b200 import workloads as W

from .nnref import Linear, Sequential, Tanh, train_step
from .tape import Var


def setup(w: W.Workload, rank: int = 0, world: int = 1):
    model, x, t = W.setup(w, Linear, Tanh, Sequential, rank, world)
    return model, Var(x), Var(t)


def run_steps(w: W.Workload, steps: int, model=None, x=None, t=None) -> Dict[str, object]:
    """Run `steps` training steps; return losses, last-step grads and final parameters."""
    if model is None:
        model, x, t = setup(w)
    repair = (x.ndim == 3)
    losses: List[np.ndarray] = []
    for _ in range(steps):
        out = train_step(model, x, t, w.loss, w.lr, w.pred_map, w.squeeze_dim, repair)
        losses.append(out.data.copy())
    params = list(model.parameters())
    return {
        "losses": np.array(losses, dtype=np.float32),
        "grads": [p.grad.copy() for p in params],
        "params": [p.data.copy() for p in params],
    }


def forward_only(w: W.Workload, model, x: Var) -> np.ndarray:
    return np.ascontiguousarray(model(x, x.ndim == 3).data)
