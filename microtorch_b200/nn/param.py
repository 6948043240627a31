"""Parameter: a Tensor that always requires grad, initialised ON THE HOST with the global
NumPy RNG exactly like the reference (src/nn/param.py:11-66), so the same seed
gives the same weights on every backend.  Moving it to the GPU is one H2D copy."""

from __future__ import annotations

from typing import Any, Literal, Optional, Tuple

import numpy as np

from ..tensor import Tensor
from ..tensor.device import Device

InitMethod = Literal["xavier", "he", "normal", "uniform"]


class Parameter(Tensor):
    def __init__(self, *shape: int, data: Optional[np.ndarray] = None, init_method: InitMethod = "normal",
                 gain: float = 1.0, device: Device = Device.CPU) -> None:
        if data is None:
            data = self._initialize(shape, init_method, gain, device)
        super().__init__(data=data, requires_grad=True, device=device)

    def _initialize(self, shape: Tuple[int, ...], method: InitMethod | Any, gain: float, device: Device):
        """Draws come from np.random on the host; the scaling is a float32 Tensor multiply."""
        draw = {
            "xavier": lambda: gain * Tensor(2.0 / sum(shape), device=device).sqrt() * Tensor.randn(shape, device=device),
            "he": lambda: gain * Tensor(2.0 / shape[0], device=device).sqrt() * Tensor.randn(shape, device=device),
            "normal": lambda: gain * Tensor.randn(shape, device=device),
            "uniform": lambda: gain * Tensor.uniform(-1, 1, shape, device=device),
        }.get(method)
        if draw is None:
            raise ValueError(f"Unknown initialization method: {method}")
        return draw()

    def to(self, device: Device) -> "Parameter":
        """Always a NEW Parameter object around the moved (or shared) data, fresh zero grad
        (reference param.py:60-66) - optimizers must re-enumerate parameters."""
        moved = super().to(device)
        return Parameter(*moved.shape, data=moved.data, device=device)
