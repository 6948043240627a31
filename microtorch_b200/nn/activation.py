from ..tensor import Tensor
from .module import Module


class Tanh(Module):
    """Tanh activation (reference src/nn/activation.py).  Inside a Sequential on CUDA a
    Linear -> Tanh pair is executed as ONE fused GEMM-epilogue node (see sequential.py)."""

    def forward(self, input: Tensor) -> Tensor:
        return input.tanh()
