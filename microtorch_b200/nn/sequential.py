"""Sequential container (interface of the reference's src/nn/sequential.py): children kept in
a tuple, parameters enumerated child by child (weight, bias per Linear).

CUDA peephole: a `Linear` immediately followed by a `Tanh` runs as one fused kernel and one
autograd node (bias + tanh in the GEMM epilogue, tanh' in the backward GEMM prologue)."""

from __future__ import annotations

from typing import Iterator

from ..tensor.device import Device
from ..tensor.tensor import Tensor
from .activation import Tanh
from .layer.linear import Linear
from .module import Module, Parameter, auto_device


class Sequential(Module):
    def __init__(self, *modules: Module) -> None:
        super().__init__()
        self.modules = modules

    def parameters(self) -> Iterator[Parameter]:
        for module in self.modules:
            if isinstance(module, Module):
                yield from module.parameters()

    def train(self) -> None:
        super().train()
        for module in self.modules:
            module.train()

    def eval(self) -> None:
        super().eval()
        for module in self.modules:
            module.eval()

    def forward(self, x: Tensor) -> Tensor:
        mods = self.modules
        i, n = 0, len(mods)
        private_ctx = None      # tape context of x when x is an intermediate only this loop holds
        while i < n:
            m = mods[i]
            fuse = (type(m) is Linear and i + 1 < n and type(mods[i + 1]) is Tanh
                    and isinstance(x, Tensor) and x.device == Device.CUDA)
            if fuse:
                for part in (m, mods[i + 1]):           # keep the first-call device migration
                    if not getattr(part, "_device_initialized", False):
                        part.to(x.device)
                        part._device_initialized = True
                x = m.forward_fused(x, activation="tanh", x_fused_ctx=private_ctx)
                private_ctx = getattr(x, "_fused_ctx", None)
                i += 2
            else:
                x = m(x)
                private_ctx = None
                i += 1
        return x
