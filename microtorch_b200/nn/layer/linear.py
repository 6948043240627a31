"""Linear layer (interface of the reference's src/nn/layer/linear.py).

Semantics kept from the reference:
  * weight (out, in) and bias (out,) initialised uniform(-1,1)*0.1 on the host (linear.py:20-21);
  * ndim assert and feature-dimension ValueError (linear.py:35-41);
  * the bias is added IN PLACE, outside the autograd tape, so it never receives a gradient and
    SGD never moves it (linear.py:61-62 -> tensor.py:547-550).
Repaired: 3-D inputs (the reference raises in both forward and backward, Q4) are handled as
the flattened (B*S, in) problem, so dW = sum over all rows - the intended semantics.

On CUDA the layer is one fused node: Y = act(X W^T + b) in a single tcgen05 GEMM launch whose
epilogue adds the bias (and applies tanh when fused by Sequential); backward is
`mt_linear_bwd_f32`: dW and dX GEMMs with tanh' applied in their operand prologue.
"""

from __future__ import annotations

from typing import Optional

from ... import _capi
from ...tensor.backend import backend_for
from ...tensor.device import Device
from ...tensor.tensor import Tensor
from ...tensor.types import Leaf
from ..module import Module
from ..param import Parameter


class Linear(Module):
    def __init__(self, in_features: int, out_features: int, bias: bool = True) -> None:
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.weight = Parameter(out_features, in_features, init_method="uniform", gain=.1)
        self.bias = Parameter(out_features, init_method="uniform", gain=.1) if bias else None

    def _check(self, x: Tensor) -> None:
        assert x.ndim in (2, 3), f"Input must be 2D or 3D Tensor! x.ndim={x.ndim}"
        if x.shape[-1] != self.in_features:
            raise ValueError(
                f"Last dimension of input: {x.shape[-1]} does not match in_features: {self.in_features}")

    def forward(self, x: Tensor) -> Tensor:
        self._check(x)
        if x.device == Device.CUDA:
            return self.forward_fused(x, activation=None)
        # CPU: the reference's op sequence  (W @ x^T)^T, then the in-place bias add
        axes = (0, 2, 1) if x.ndim == 3 else None
        out = (self.weight @ x.transpose(axes)).transpose(axes)
        if self.bias is not None:
            b = self.bias.unsqueeze(dim=0)
            if out.data.ndim == 3:
                b = b.unsqueeze(dim=0)
            out += b
        return out

    def forward_fused(self, x: Tensor, activation: Optional[str] = None, x_fused_ctx: Optional[dict] = None) -> Tensor:
        """CUDA: act(X W^T + b) as one kernel and ONE tape node (activation None | "tanh").

        `x_fused_ctx`: set by Sequential when `x` is the PRIVATE output of the preceding fused
        Linear+Tanh node (nobody else can hold it).  This node's dX epilogue then multiplies by
        (1 - x^2) and hands the producer its pre-activation gradient dZ directly, so the producer
        runs two plain GEMMs - tanh' fused into the producer side, no prologue, no extra pass."""
        self._check(x)
        K = backend_for(Device.CUDA).K
        weight, bias = self.weight, self.bias
        lead = tuple(x.shape[:-1])
        x2d = x.data.reshape((-1, self.in_features))
        act = _capi.ACT_TANH if activation == "tanh" else _capi.ACT_NONE
        y2d = K.linear_forward(x2d, weight.data, bias.data if bias is not None else None, act)
        y = y2d.reshape(lead + (self.out_features,))
        need_dx = x.requires_grad
        cache = {}
        ctx = {"grad_is_preact": False}          # flipped by the next fused layer (see x_fused_ctx above)
        chain = x_fused_ctx is not None and need_dx
        if chain:
            x_fused_ctx["grad_is_preact"] = True

        def both(g):
            """dW and dX come out of one C call; the two tape edges share the result."""
            if cache.get("for") is not g:
                y_for_tanh = y2d if (act and not ctx["grad_is_preact"]) else None
                dW, dX = K.linear_backward(g.reshape((-1, self.out_features)), y_for_tanh, x2d, weight.data,
                                           need_dx, dx_preact=chain)
                cache.update({"for": g, "dW": dW, "dX": dX.reshape(x.shape) if dX is not None else None})
            return cache

        deps = [Leaf(value=weight, grad_fn=lambda g: both(g)["dW"])]
        if need_dx:
            deps.append(Leaf(value=x, grad_fn=lambda g: both(g)["dX"]))
        out = Tensor(y, requires_grad=True, dependencies=deps, device=Device.CUDA, dtype=x.dtype)
        out._fused_ctx = ctx if act else None
        return out
