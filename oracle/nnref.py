"""Oracle restatement of the reference's nn layer (src/nn/*) on top of oracle.tape.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Citations are /root/reference paths.
"""

from __future__ import annotations

from typing import Callable, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from .tape import F32, Var, matmul

# ----------------------------------------------------------------------------
# Parameter / Linear  (src/nn/param.py:11-66, src/nn/layer/linear.py:11-64)
# ----------------------------------------------------------------------------


def uniform_param(*shape: int, gain: float = 0.1) -> Var:
    """Parameter(*shape, init_method="uniform", gain=gain) - param.py:54-55:
    ``gain * Tensor.uniform(-1, 1, shape)``: float64 draws from the GLOBAL np.random
    state, cast to float32 (tensor.py:429-434), then a float32 multiply by gain."""
    draws = np.random.uniform(-1, 1, shape)          # tensor.py:429
    data = np.array(gain, dtype=F32) * draws.astype(F32)
    return Var(data, requires_grad=True)


class Linear:
    """src/nn/layer/linear.py.  Bias is added IN PLACE and so never gets a gradient (Q1)."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True) -> None:
        self.in_features, self.out_features = in_features, out_features
        self.weight = uniform_param(out_features, in_features)          # linear.py:20
        self.bias = uniform_param(out_features) if bias else None       # linear.py:21

    def parameters(self) -> Iterable[Var]:
        yield self.weight
        if self.bias is not None:
            yield self.bias

    def __call__(self, x: Var, repair_3d: bool = False) -> Var:
        assert x.ndim in (2, 3), f"Input must be 2D or 3D Tensor! x.ndim={x.ndim}"   # linear.py:35
        if x.shape[-1] != self.in_features:                                            # linear.py:38-41
            raise ValueError(f"Last dimension of input: {x.shape[-1]} does not match "
                             f"in_features: {self.in_features}")
        axes = (0, 2, 1) if x.ndim == 3 else None                # linear.py:44
        out = matmul(self.weight, x.transpose(axes), repair_3d)  # linear.py:48   W @ x^T
        out = out.transpose(axes)                                # linear.py:52   F-contiguous view
        if self.bias is not None:
            b = self.bias.unsqueeze(0)                           # linear.py:57
            if out.ndim == 3:
                b = b.unsqueeze(0)                               # linear.py:59 (repair S2)
            out.iadd_(b)                                         # linear.py:62  no tape
        return out


class Tanh:
    def parameters(self):
        return iter(())

    def __call__(self, x: Var, repair_3d: bool = False) -> Var:   # src/nn/activation.py:11-12
        return x.tanh()


class Sequential:
    """src/nn/sequential.py: children in order; parameters = weight, bias per Linear."""

    def __init__(self, *modules) -> None:
        self.modules = modules

    def parameters(self) -> Iterable[Var]:
        for m in self.modules:
            yield from m.parameters()

    def zero_grad(self) -> None:            # src/nn/module.py:83-89
        for p in self.parameters():
            p.zero_grad()

    def __call__(self, x: Var, repair_3d: bool = False) -> Var:
        for m in self.modules:
            x = m(x, repair_3d)
        return x


# ----------------------------------------------------------------------------
# Losses  (src/nn/loss.py)
# ----------------------------------------------------------------------------

def _check(pred: Var, target: Var) -> None:
    assert pred.shape == target.shape, "Input and target must have the same shape"   # loss.py:40-42


def mse_loss(pred: Var, target: Var) -> Var:            # loss.py:63-76
    _check(pred, target)
    return ((pred - target) ** 2).mean()


def ce_loss(pred: Var, target: Var) -> Var:             # loss.py:79-84  (Q3: log(p) - t)
    _check(pred, target)
    return (-(target - pred.log())).mean()


def bce_loss(pred: Var, target: Var) -> Var:            # loss.py:87-100
    _check(pred, target)
    return (-(target * pred.log() + (1 - target) * (1 - pred).log())).mean()


LOSSES = {"mse": mse_loss, "ce": ce_loss, "bce": bce_loss}


# ----------------------------------------------------------------------------
# SGD  (src/nn/optimizer.py:21-39)
# ----------------------------------------------------------------------------

def sgd_step(parameters: Callable[[], Iterable[Var]], lr: float) -> None:
    for p in parameters():
        p.isub_(lr * p.grad)      # optimizer.py:39 -> tensor.py:564-571


# ----------------------------------------------------------------------------
# One training step, as the notebook loop does it (demo.ipynb:5178-5189)
# ----------------------------------------------------------------------------

def train_step(model: Sequential, x: Var, target: Var, loss: str, lr: float,
               pred_map: Optional[Tuple[float, float]] = None,
               squeeze_dim: Optional[int] = None, repair_3d: bool = False) -> Var:
    """zero_grad -> forward -> loss -> backward -> SGD.  Returns the loss Var.

    pred_map=(a, b) applies ``pred*a + b`` before the loss (keeps
    log() finite for CE/BCE); squeeze_dim mirrors ``y_pred.squeeze(1)`` of the notebook."""
    model.zero_grad()
    pred = model(x, repair_3d)
    if squeeze_dim is not None:
        pred = pred.squeeze(squeeze_dim)
    if pred_map is not None:
        pred = pred * pred_map[0] + pred_map[1]
    out = LOSSES[loss](pred, target)
    out.backward()
    sgd_step(model.parameters, lr)
    return out
