"""Synthetic workloads of BASELINE.json (C1..C5): seeds, shapes and host-side data.

Pure NumPy data generation, shared by bench.py, the tests and the oracle so that every
backend sees the same bytes.  No arithmetic of the training path lives here.
Definitions follow SURVEY.md section 8(d); the spiral dataset restates demo.ipynb:55-89
of the reference (same np.random call order, so seed 1337 reproduces the notebook).
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

F32 = np.float32


@dataclass(frozen=True)
class Workload:
    name: str
    dims: Tuple[int, ...]                 # layer widths, input first
    loss: str                             # "mse" | "ce" | "bce"
    lr: float
    batch: int
    pred_map: Optional[Tuple[float, float]] = None   # pred*a + b before the loss
    squeeze_dim: Optional[int] = None
    #: construction order of the Linear layers (RNG consumption order); None = in order
    init_order: Optional[Tuple[int, ...]] = None
    seed: int = 0
    input_shape: Optional[Tuple[int, ...]] = None     # default (batch, dims[0])

    @property
    def n_layers(self) -> int:
        return len(self.dims) - 1

    def gemm_flops_per_step(self, batch: Optional[int] = None) -> float:
        """Logical GEMM flops of one step: fwd + dW for every layer, dX for all but the first."""
        b = float(batch if batch is not None else self.batch)
        rows = b if self.input_shape is None else float(np.prod(self.input_shape[:-1])) * b / self.batch
        total = 0.0
        for i in range(self.n_layers):
            per = 2.0 * rows * self.dims[i] * self.dims[i + 1]
            total += per * (3 if i > 0 else 2)
        return total


def spiral_dataset(n_samples: int = 200, noise: float = 0.9) -> Tuple[np.ndarray, np.ndarray]:
    """Two noisy interleaved spirals scaled into [-1,1]^2, labels -1/+1 (float64, as the
    notebook hands them to Tensor()).  RNG order: rand(n), rand(n), randn(n_samples, 2)."""
    half = n_samples // 2
    arms = []
    for sign in (+1.0, -1.0):
        theta = np.sqrt(np.random.rand(half)) * 4 * np.pi
        radius = sign * (2 * theta + np.pi)
        arms.append(np.stack([radius * np.cos(theta), radius * np.sin(theta)], axis=1))
    pts = np.vstack(arms)
    pts += np.random.randn(n_samples, 2) * noise
    for col in (0, 1):
        lo, hi = pts[:, col].min(), pts[:, col].max()
        pts[:, col] = np.interp(pts[:, col], (lo, hi), (-1, 1))
    labels = np.zeros(n_samples)
    labels[half:] = 1
    return pts, labels * 2 - 1


C1A = Workload("c1a_spiral_notebook", (2, 64, 32, 16, 1), "mse", 0.5, 200,
               squeeze_dim=1, init_order=(1, 2, 3, 0), seed=1337)
C1B = Workload("c1b_spiral_2_100_3", (2, 100, 3), "mse", 0.5, 200, seed=1337)
C2 = Workload("c2_mlp_1024_4096_4096_10", (1024, 4096, 4096, 10), "ce", 0.01, 65536,
              pred_map=(0.49, 0.5), seed=0)
C2_RAW = Workload("c2_raw_nan", (1024, 4096, 4096, 10), "ce", 0.01, 65536, seed=0)
C4 = Workload("c4_3d_linear_bce", (2048, 2048, 2048), "bce", 0.01, 256,
              pred_map=(0.49, 0.5), seed=0, input_shape=(256, 512, 2048))
C5 = Workload("c5_dp_mlp_4096x8", (4096,) * 9, "mse", 0.01, 524288, seed=0)


def scaled(w: Workload, batch: int, dims: Optional[Sequence[int]] = None,
           input_shape: Optional[Tuple[int, ...]] = None) -> Workload:
    """Same workload at a size the CPU oracle finishes in seconds."""
    return Workload(w.name + f"_b{batch}", tuple(dims) if dims else w.dims, w.loss, w.lr, batch,
                    w.pred_map, w.squeeze_dim, w.init_order, w.seed, input_shape)


def build_model(w: Workload, linear: Callable, tanh: Callable, sequential: Callable):
    """Construct Linear/Tanh/... with the given classes, consuming np.random in the
    reference's order (weight then bias per Linear, layers in `init_order`)."""
    n = w.n_layers
    order = w.init_order or tuple(range(n))
    layers: List = [None] * n
    for i in order:
        layers[i] = linear(w.dims[i], w.dims[i + 1])
    mods = []
    for lyr in layers:
        mods += [lyr, tanh()]
    return sequential(*mods)


def make_data(w: Workload, rank: int = 0, world: int = 1) -> Tuple[np.ndarray, np.ndarray]:
    """Host inputs/targets for workload `w` (call AFTER build_model for C2/C4/C5, BEFORE it
    for C1 - see `setup`).  With world > 1 returns this rank's contiguous row shard."""
    shape = w.input_shape or (w.batch, w.dims[0])
    out_shape = shape[:-1] + (w.dims[-1],)
    if w.name.startswith("c1"):
        x, y = spiral_dataset(w.batch, 0.9)
        if w.dims[-1] == 1:
            t = y
        else:                                   # C1b: one-hot(3)*2-1 of label -1/+1 -> classes 0/2
            cls = ((y + 1)).astype(np.int64)
            t = np.eye(w.dims[-1])[cls] * 2 - 1
    elif w.loss == "ce":
        x = np.random.randn(*shape).astype(F32)
        t = np.eye(w.dims[-1], dtype=F32)[np.random.randint(0, w.dims[-1], shape[:-1])]
    elif w.loss == "bce":
        x = np.random.randn(*shape).astype(F32)
        t = (np.random.rand(*out_shape) > 0.5).astype(F32)
    else:
        x = np.random.randn(*shape).astype(F32)
        t = np.random.uniform(-1, 1, out_shape).astype(F32)
    if world > 1:
        n = shape[0] // world
        x, t = x[rank * n:(rank + 1) * n], t[rank * n:(rank + 1) * n]
    return x, t


def setup(w: Workload, linear: Callable, tanh: Callable, sequential: Callable,
          rank: int = 0, world: int = 1):
    """Seed, then build data and model in the order the workload defines."""
    import random
    np.random.seed(w.seed)
    random.seed(w.seed)
    if w.name.startswith("c1"):                 # notebook: dataset first, then the model
        x, t = make_data(w, rank, world)
        model = build_model(w, linear, tanh, sequential)
    else:                                       # model on host first, then X, T
        model = build_model(w, linear, tanh, sequential)
        x, t = make_data(w, rank, world)
    return model, x, t
