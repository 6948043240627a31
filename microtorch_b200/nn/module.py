"""Module base class (interface of the reference's src/nn/module.py): first-call device
migration, train/eval flags, parameter enumeration in attribute order, zero_grad, to() - plus what the
reference lacks for resuming a run: named parameters and a state dict (see checkpoint.py)."""

from __future__ import annotations

from functools import wraps
from typing import Any, Dict, Iterator, Tuple

from ..tensor import Tensor
from ..tensor.device import Device
from .param import Parameter


def auto_device(forward_fn):
    """On the first call, move the module to the device of the first Tensor argument."""
    @wraps(forward_fn)
    def call(self: "Module", *args, **kwargs):
        if not getattr(self, "_device_initialized", False):
            first = next((a for a in (*args, *kwargs.values()) if isinstance(a, Tensor)), None)
            if first is not None:
                self.to(first.device)
                self._device_initialized = True
        return forward_fn(self, *args, **kwargs)
    return call


class Module:
    def __init__(self) -> None:
        self.train_mode = True
        self._device_initialized = False

    @auto_device
    def __call__(self, *args: Any) -> Tensor:
        return self.forward(*args)

    def forward(self, *input: Any) -> Tensor:
        raise NotImplementedError()

    def _children(self) -> Iterator["Module"]:
        return (v for v in self.__dict__.values() if isinstance(v, Module))

    def train(self) -> None:
        self.train_mode = True
        for child in self._children():
            child.train()

    def eval(self) -> None:
        self.train_mode = False
        for child in self._children():
            child.eval()

    def parameters(self) -> Iterator[Parameter]:
        for value in self.__dict__.values():
            if isinstance(value, Parameter):
                yield value
            if isinstance(value, Module):
                yield from value.parameters()

    def zero_grad(self) -> None:
        for p in self.parameters():
            p.zero_grad()

    def params_count(self) -> int:
        return sum(p.data.size for p in self.parameters())

    def to(self, device: Device) -> "Module":
        for name, value in list(self.__dict__.items()):
            if isinstance(value, Tensor):
                setattr(self, name, value.to(device))
            elif isinstance(value, Module):
                value.to(device)
        return self

    # ------------------------------------------------------------------ names / persistence (not in the reference)
    def named_parameters(self) -> Iterator[Tuple[str, Parameter]]:
        """(attribute path, Parameter) in the order of `parameters()`, e.g. "modules.0.weight"."""
        from ..checkpoint import named_parameters
        return named_parameters(self)

    def state_dict(self) -> Dict[str, Any]:
        """Host copies of every parameter by name."""
        from ..checkpoint import state_dict
        return state_dict(self)

    def load_state_dict(self, state: Dict[str, Any], strict: bool = True) -> None:
        """Write `state` into the existing parameter buffers (host or device) in place."""
        from ..checkpoint import load_state_dict
        load_state_dict(self, state, strict=strict)
