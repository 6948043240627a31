"""Whole-step CUDA graphs for the training loop.

The reference's user loop (demo.ipynb:5178-5189) issues ~30 tiny array-library calls per step; at the spiral-demo
sizes the step is pure launch latency (0.30 ms for 14 fused launches here, 2.3 ms on the notebook's CuPy run).
`GraphedStep` captures one step - forward, loss, backward, SGD on FIXED tensors - into a CUDA graph and replays it
with a single launch.  The tape, the lazily adopted gradient buffers and every intermediate live in pool blocks
that stay reserved for the graph, so after each replay `loss`, `param.grad` and `param.data` hold that step's values.

    step = GraphedStep(lambda: train_step(model, loss_fn, opt, X, T))
    for _ in range(5000):
        step.replay()
    print(step.result.item())

Rules: the callable must not synchronise (no `.item()`, `.get()`, `Tensor(host_data, device="cuda")`); new input
data is fed by copying into the captured input tensors (`X.data[...] = new`, or `cuda.prefetch` + `mt_d2d`)."""

from __future__ import annotations

import ctypes as C
from typing import Callable

from . import _capi


class GraphedStep:
    def __init__(self, fn: Callable[[], object], warmup: int = 2) -> None:
        _capi.require_device()
        self._lib = _capi.lib()
        self.result = None
        for _ in range(max(warmup, 1)):         # eager: fills the constant cache, kernel attributes, scratch buffers
            self.result = fn()
        _capi.check(self._lib.mt_sync(), "mt_sync")
        _capi.check(self._lib.mt_graph_begin(), "mt_graph_begin")
        handle = C.c_void_p()
        try:
            self.result = fn()
        finally:
            rc = self._lib.mt_graph_end(C.byref(handle))
        _capi.check(rc, "mt_graph_end")
        self._handle = handle
        self.replays = 0

    def replay(self) -> None:
        """Run the captured step once (asynchronous; read `result` / parameters to synchronise)."""
        _capi.check(self._lib.mt_graph_launch(self._handle), "mt_graph_launch")
        self.replays += 1

    def close(self) -> None:
        h, self._handle = getattr(self, "_handle", None), None
        if h:
            self._lib.mt_graph_destroy(h)

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:
            pass
