"""Losses (interface of the reference's src/nn/loss.py): same classes, same shape assert, same
formulas - including the reference's "CrossEntropyLoss", which is mean(log(pred) - target)
(loss.py:79-84), not textbook cross-entropy.

On CUDA, MSE / CE / BCE are one fused kernel: loss value and dL/dpred in a single pass over
pred and target (12 B/element) instead of 6 / 7 / 13 forward nodes and as many backward
visits.  L1Loss needs `abs` (comparisons) and stays CPU-only, as in the reference (Q8)."""

from __future__ import annotations

from typing import Literal, Optional

from .. import _capi
from ..tensor import Tensor
from ..tensor.backend import backend_for
from ..tensor.device import Device
from ..tensor.types import Leaf
from .module import Module


def reduction_loss(loss: Tensor, reduction: Literal["mean", "sum", "none"] = "mean") -> Tensor:
    if reduction == "mean":
        return loss.mean()
    if reduction == "sum":
        return loss.sum()
    return loss


class Loss(Module):
    #: fused CUDA kernel id (None = compose from Tensor ops)
    _kind: Optional[int] = None

    def __init__(self, reduction: Literal["mean", "sum", "none"] = "mean"):
        self.reduction = reduction
        #: data-parallel training: number of ranks whose shards make up the global batch.  The mean
        #: is taken over the GLOBAL element count so that summed rank gradients equal the
        #: single-process gradient.
        self.world = 1

    def compute_loss(self, pred: Tensor, target: Tensor) -> Tensor:
        raise NotImplementedError

    def forward(self, pred: Tensor, target: Tensor) -> Tensor:
        assert pred.shape == target.shape, "Input and target must have the same shape"
        if (self._kind is not None and self.reduction == "mean" and pred.device == Device.CUDA
                and target.device == Device.CUDA):
            return self._fused(pred, target)
        loss = reduction_loss(self.compute_loss(pred, target), self.reduction)
        world = getattr(self, "world", 1)
        return loss if world == 1 else loss * (1.0 / world)

    def _fused(self, pred: Tensor, target: Tensor) -> Tensor:
        backend = backend_for(Device.CUDA)
        count = pred.size * getattr(self, "world", 1)
        value, dpred = backend.K.loss_forward_backward(self._kind, pred.data, target.data, count, pred.requires_grad)
        deps = []
        if pred.requires_grad:
            from ..cuda.array import device_constant
            seed = device_constant(1.0)         # what Tensor.backward() seeds a 0-d root with (read-only, shared)

            def to_pred(g):
                # dpred already holds dL/dpred for an upstream gradient of 1 - recognised by IDENTITY with the shared
                # constant, never by a host mirror of the value that a write through a view could leave stale
                if g is seed:
                    return dpred
                return backend.mul(dpred, g)
            deps.append(Leaf(value=pred, grad_fn=to_pred))
        return Tensor(value, requires_grad=pred.requires_grad, dependencies=deps, device=Device.CUDA, dtype=pred.dtype)


class L1Loss(Loss):
    def __init__(self):
        super().__init__(reduction="mean")

    def compute_loss(self, prediction: Tensor, target: Tensor) -> Tensor:
        return (prediction - target).abs()


class MSELoss(Loss):
    _kind = _capi.LOSS_MSE

    def __init__(self):
        super().__init__(reduction="mean")

    def compute_loss(self, prediction: Tensor, target: Tensor) -> Tensor:
        return (prediction - target).pow(2)


class CrossEntropyLoss(Loss):
    _kind = _capi.LOSS_CE_REF

    def __init__(self):
        super().__init__(reduction="mean")

    def compute_loss(self, prediction: Tensor, target: Tensor) -> Tensor:
        return -(target - prediction.log())


class BCELoss(Loss):
    _kind = _capi.LOSS_BCE

    def __init__(self):
        super().__init__(reduction="mean")

    def compute_loss(self, prediction: Tensor, target: Tensor) -> Tensor:
        return -(target * prediction.log() + (1 - target) * (1 - prediction).log())
