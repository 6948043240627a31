"""Optimizers (interface of the reference's src/nn/optimizer.py).  `parameters` is the CALLABLE
`model.parameters`, re-enumerated every step, because Parameter objects are replaced when a
module migrates on its first call."""

from __future__ import annotations

from typing import Callable, Iterator

from ..tensor.backend import backend_for
from ..tensor.device import Device
from ..tensor.tensor import _LAZY
from .param import Parameter


class Optimizer(object):
    def __init__(self, parameters: Callable[[], Iterator[Parameter]]) -> None:
        self.parameters = parameters

    def step(self) -> None:
        raise NotImplementedError()


class SGD(Optimizer):
    def __init__(self, parameters: Callable[[], Iterator[Parameter]], lr: float = .01) -> None:
        super(SGD, self).__init__(parameters=parameters)
        self.lr = lr

    def step(self, skip=None) -> None:
        """p <- p - lr * grad.  CUDA parameters are updated by ONE fused multi-tensor launch
        (reference: three kernels and two temporaries per parameter, optimizer.py:39 ->
        tensor.py:564-571).  A parameter whose gradient is still the unallocated zero - every
        Linear bias, Q1 - is skipped: p - lr*0 == p.  `skip`: ids of parameters somebody already
        updated in this step (the data-parallel trainer updates a weight right behind its all-reduce)."""
        dev_params, dev_grads = [], []
        for parameter in self.parameters():
            if skip and id(parameter) in skip:
                continue
            if parameter.device == Device.CUDA:
                g = parameter._grad
                if g is None or g is _LAZY:
                    continue
                dev_params.append(parameter.data)
                dev_grads.append(parameter.grad)
            else:
                parameter -= self.lr * parameter.grad
        if dev_params:
            backend_for(Device.CUDA).K.sgd_multi(dev_params, dev_grads, self.lr)
