"""Activations.  The reference ships `Tanh` only (src/nn/activation.py:6-12); `ReLU` and `Sigmoid` are composed from
ops that run on the device here (comparisons / where / exp), which the reference's CuPy path cannot do."""

from __future__ import annotations

from ..tensor import Tensor
from .module import Module


class Tanh(Module):
    """y = tanh(x).  Inside a `Sequential` on CUDA a Linear -> Tanh pair runs as ONE fused node: tanh in the GEMM
    epilogue, tanh' folded into the backward GEMMs (see sequential.py / layer/linear.py)."""

    def forward(self, input: Tensor) -> Tensor:
        return input.tanh()


class ReLU(Module):
    """y = max(x, 0), gradient 1 where x > 0 (a `where` select on the operand's device)."""

    def forward(self, input: Tensor) -> Tensor:
        return Tensor.where(input > 0, input, input * 0.0)


class Sigmoid(Module):
    """y = 1 / (1 + exp(-x)), composed from neg / exp / add / pow(-1)."""

    def forward(self, input: Tensor) -> Tensor:
        return ((-input).exp() + 1.0) ** -1
