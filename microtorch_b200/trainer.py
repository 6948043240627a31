"""Training-step drivers: the notebook loop as a function, and its data-parallel version.

`train_step` is the reference's loop body (demo.ipynb:5178-5189): zero_grad, forward, loss,
backward, SGD.  `DataParallelTrainer` is the new component BASELINE.json asks for (SURVEY.md
section 8e): one process per GPU, every rank holds a full replica and a contiguous row shard
of the batch, weight gradients are summed with NCCL over NVLink.  Each weight's all-reduce
is queued on the communication stream the moment that weight's gradient is final (a tape
hook - the backward engine finalises a parameter as soon as its last incoming edge has been
processed, and the exchange is gated on the dW GEMM alone, not on the dX GEMM queued behind
it), and `W -= lr * dW_sum` follows on the same stream, so exchange AND update of layer L run
under the backward GEMMs of the layers below it; only the first layer's are exposed.  The
compute stream waits for the communication stream once, before the next reader of the weights.  The loss is scaled by the GLOBAL
element count, so summed rank gradients equal the single-process gradients up to fp32
summation order and no extra 1/world pass is needed.  Bias gradients are identically zero
(Q1) and are never exchanged.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Callable, Optional, Tuple

import numpy as np

from . import _capi
from .nn.loss import Loss
from .nn.module import Module
from .nn.optimizer import SGD
from .tensor import Tensor
from .tensor.device import Device
from .tensor.tensor import _LAZY


def train_step(model: Module, loss_fn: Loss, optimizer: SGD, x: Tensor, target: Tensor,
               pred_map: Optional[Tuple[float, float]] = None, squeeze_dim: Optional[int] = None) -> Tensor:
    """One step of the reference's training loop.  Returns the loss tensor (not synchronised)."""
    model.zero_grad()
    pred = model(x)
    if squeeze_dim is not None:
        pred = pred.squeeze(squeeze_dim)
    if pred_map is not None:
        pred = pred * pred_map[0] + pred_map[1]
    loss = loss_fn(pred, target)
    loss.backward()
    optimizer.step()
    return loss


def _small_mlp_plan(model: Module, loss_fn: Loss, optimizer: SGD, x: Tensor, target: Tensor):
    """(layers, acts) when `model` is a Sequential of Linear [+ Tanh] blocks of the kind the single-launch trainer
    (mt_mlp_fit_f32) runs; None otherwise.  Pure host logic; whether model + batch fit on chip is the library's call."""
    from .nn.activation import Tanh
    from .nn.layer.linear import Linear
    from .nn.sequential import Sequential
    if type(model) is not Sequential or type(optimizer) is not SGD or loss_fn._kind is None or loss_fn.reduction != "mean":
        return None
    if getattr(loss_fn, "world", 1) != 1 or x.device != Device.CUDA or target.device != Device.CUDA or x.ndim != 2:
        return None
    layers, acts, mods, i = [], [], model.modules, 0
    while i < len(mods):
        if type(mods[i]) is not Linear:
            return None
        layers.append(mods[i])
        if i + 1 < len(mods) and type(mods[i + 1]) is Tanh:
            acts.append(_capi.ACT_TANH)
            i += 2
        else:
            acts.append(_capi.ACT_NONE)
            i += 1
    if not 1 <= len(layers) <= _capi.MLP_MAX_LAYERS or layers[0].in_features != x.shape[1]:
        return None
    if any(a.out_features != b.in_features for a, b in zip(layers[:-1], layers[1:])):
        return None
    # the per-step path and the reference assert pred.shape == target.shape ("Input and target must have the same
    # shape", loss.py:29): the single-launch loop must not train on a transposed / differently shaped target.
    # fit() accepts (B,) as well when it squeezes a width-1 prediction.
    out = layers[-1].out_features
    if tuple(target.shape) != (x.shape[0], out) and not (out == 1 and tuple(target.shape) == (x.shape[0],)):
        return None
    return layers, acts


def fit(model: Module, loss_fn: Loss, optimizer: SGD, x: Tensor, target: Tensor, steps: int,
        pred_map: Optional[Tuple[float, float]] = None, squeeze_dim: Optional[int] = None,
        fused: Optional[bool] = None) -> np.ndarray:
    """`steps` iterations of the reference's full-batch loop (demo.ipynb:5178-5189) on fixed x / target;
    returns the float32 loss of every step.

    A small Linear(+Tanh) stack on CUDA runs as ONE kernel launch for the whole loop (mt_mlp_fit_f32: weights,
    activations and gradients never leave the shared memory of a thread-block cluster); anything else runs `train_step` per iteration.  Both are
    device paths.  `fused=True` insists on the single-launch kernel (ValueError if the model does not qualify),
    `fused=False` forces the per-step path.  Afterwards parameters hold the trained weights and `.grad` the last
    step's gradients, exactly as after the loop."""
    plan = None if fused is False else _small_mlp_plan(model, loss_fn, optimizer, x, target)
    if plan is not None:
        # pred is (B, out), or (B,) after squeeze(1) of a width-1 prediction; anything else is the per-step path's
        # business (it raises the reference's AssertionError)
        want = (x.shape[0],) if squeeze_dim is not None else (x.shape[0], plan[0][-1].out_features)
        if tuple(target.shape) != want or (squeeze_dim is not None and plan[0][-1].out_features != 1):
            plan = None
    if plan is None:
        if fused:
            raise ValueError("fit(fused=True): the model / loss / batch do not qualify for the single-launch trainer")
        return np.array([train_step(model, loss_fn, optimizer, x, target, pred_map, squeeze_dim).item()
                         for _ in range(steps)], dtype=np.float32)
    import ctypes as C
    from .cuda.array import DeviceArray
    layers, acts = plan
    for m in (model, *model.modules):                       # the first-call device migration (Q6)
        if not getattr(m, "_device_initialized", False):
            m.to(x.device)
            m._device_initialized = True
    desc = _capi.MlpDesc()
    desc.n_layers = len(layers)
    desc.dims[0] = layers[0].in_features
    grads = []
    for k, layer in enumerate(layers):
        desc.dims[k + 1] = layer.out_features
        desc.act[k] = acts[k]
        wdata = layer.weight.data
        if not wdata.is_dense:
            raise ValueError("fit: weights must be dense arrays")
        desc.weight[k] = wdata.ptr
        desc.bias[k] = layer.bias.data.dense().ptr if layer.bias is not None else None
        g = DeviceArray.empty(wdata.shape)
        grads.append(g)
        desc.grad_weight[k] = g.ptr
    xd, td = x.data.dense(), target.data.dense()
    losses = DeviceArray.empty((max(steps, 1),))
    a, b = (float(pred_map[0]), float(pred_map[1])) if pred_map is not None else (1.0, 0.0)
    scale = float(np.float32(target.size) ** -1)
    status = _capi.lib().mt_mlp_fit_f32(C.byref(desc), xd.ptr, td.ptr, x.shape[0], loss_fn._kind,
                                        1 if pred_map is not None else 0, a, b, float(optimizer.lr), scale,
                                        int(steps), losses.ptr)
    if status == _capi.ERR_UNSUPPORTED:                      # model + batch do not fit in the cluster's shared memory
        if fused:
            raise ValueError("fit(fused=True): " + _capi.lib().mt_last_error().decode(errors="replace"))
        return np.array([train_step(model, loss_fn, optimizer, x, target, pred_map, squeeze_dim).item()
                         for _ in range(steps)], dtype=np.float32)
    _capi.check(status, "mt_mlp_fit_f32")
    if steps > 0:
        for layer, g in zip(layers, grads):
            layer.weight.data._host_value = None
            g._owned = True
            layer.weight.grad = g
            if layer.bias is not None:
                layer.bias.zero_grad()                       # biases get no gradient in the reference (Q1)
    return losses.get()[:steps].astype(np.float32)


def shard_rows(n_rows: int, rank: int, world: int) -> slice:
    """Contiguous row shard [r*n/W, (r+1)*n/W) of a batch (the batch must divide evenly)."""
    if n_rows % world:
        raise ValueError(f"global batch {n_rows} is not divisible by world size {world}")
    per = n_rows // world
    return slice(rank * per, (rank + 1) * per)


def nccl_library_path() -> Optional[str]:
    """libnccl.so.2 shipped in the `nvidia-nccl-cu12` wheel (no torch import needed)."""
    env = os.environ.get("MICROTORCH_B200_NCCL")
    if env:
        return env
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            for root in spec.submodule_search_locations:
                cand = os.path.join(root, "lib", "libnccl.so.2")
                if os.path.exists(cand):
                    return cand
    except Exception:
        pass
    return None


class Communicator:
    """NCCL communicator of this process (rank r of `world`), created through the C ABI.

    `exchange_id` moves the 128-byte NCCL unique id from rank 0 to everyone: any callable
    `bytes|None -> bytes` (torch.distributed broadcast over gloo in bench.py, a file in tests)."""

    def __init__(self, rank: int, world: int, exchange_id: Callable[[Optional[bytes]], bytes],
                 allgather: Optional[Callable[[bytes], list]] = None, peer_arena_bytes: Optional[int] = None) -> None:
        """`allgather` (bytes -> list of every rank's bytes, in rank order) enables the peer-memory exchange
        (csrc/mt_p2p.cu): every rank allocates an arena, the CUDA IPC handles are all-gathered, and weight gradients
        are summed and applied through NVLink loads / stores instead of ncclAllReduce.  It is used only if EVERY rank
        could map every peer; MICROTORCH_B200_P2P=0 keeps NCCL."""
        self.rank, self.world = rank, world
        self.lib = _capi.lib()
        _capi.require_device()
        path = nccl_library_path()
        _capi.check(self.lib.mt_comm_load(path.encode() if path else None), "mt_comm_load")
        uid = None
        if rank == 0:
            buf = C.create_string_buffer(128)
            _capi.check(self.lib.mt_comm_unique_id(buf), "mt_comm_unique_id")
            uid = buf.raw
        uid = exchange_id(uid)
        _capi.check(self.lib.mt_comm_init_rank(uid, rank, world), "mt_comm_init_rank")
        self.p2p = False
        self.p2p_fallbacks = 0
        if allgather is not None and 2 <= world <= 8 and os.environ.get("MICROTORCH_B200_P2P", "1") != "0":
            if peer_arena_bytes is None:
                peer_arena_bytes = int(os.environ.get("MICROTORCH_B200_P2P_ARENA_MB", "4096")) << 20
            handle = C.create_string_buffer(64)
            ok = self.lib.mt_p2p_create(peer_arena_bytes, handle) == 0
            handles = allgather(bytes([1 if ok else 0]) + handle.raw)
            if all(len(h) == 65 and h[0] == 1 for h in handles):
                ok = self.lib.mt_p2p_open(rank, world, b"".join(h[1:] for h in handles)) == 0
            else:
                ok = False
            votes = allgather(bytes([1 if ok else 0]))
            self.p2p = all(v == b"\x01" for v in votes)
            if not self.p2p:
                self.lib.mt_p2p_destroy()

    def allreduce_(self, arr) -> None:
        """In-place sum over ranks, asynchronous on the communication stream."""
        _capi.check(self.lib.mt_allreduce_sum_f32(arr.ptr, arr.size), "mt_allreduce_sum_f32")

    def allreduce_sgd_(self, param, grad, lr: float, key: Optional[int] = None) -> None:
        """grad <- sum over ranks; param -= lr * grad - both on the communication stream.  With a peer arena and a `key`
        (the same integer on every rank for the same parameter) the exchange goes through NVLink peer memory
        (mt_p2p_allreduce_sgd_f32); tensors the arena cannot take, and communicators without one, use ncclAllReduce
        (mt_allreduce_sgd_f32) - a decision every rank takes identically."""
        if self.p2p and key is not None:
            rc = self.lib.mt_p2p_allreduce_sgd_f32(param.ptr, grad.ptr, grad.size, float(lr), int(key))
            if rc == 0:
                param._host_value = None
                return
            if rc != _capi.ERR_UNSUPPORTED:
                _capi.check(rc, "mt_p2p_allreduce_sgd_f32")
            self.p2p_fallbacks += 1
        _capi.check(self.lib.mt_allreduce_sgd_f32(param.ptr, grad.ptr, grad.size, float(lr)), "mt_allreduce_sgd_f32")
        param._host_value = None

    def wait(self) -> None:
        _capi.check(self.lib.mt_comm_wait(), "mt_comm_wait")
        if self.p2p:
            code = C.c_uint(0)
            self.lib.mt_p2p_error(C.byref(code))
            if code.value:
                raise RuntimeError(f"peer-memory exchange timed out: phase {code.value & 0xff}, waiting for rank "
                                   f"{(code.value >> 8) & 0xff}, slot {code.value >> 16}")

    def close(self) -> None:
        if self.p2p:
            self.lib.mt_p2p_destroy()
            self.p2p = False
        self.lib.mt_comm_destroy()


class GlooCommunicator:
    """Same interface over torch.distributed (gloo) for Device.CPU replicas: exercises the
    sharding / loss-scaling / hook logic of the trainer without a GPU (world_size-2 tests)."""

    def __init__(self) -> None:
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def allreduce_(self, arr: np.ndarray) -> None:
        import torch
        t = torch.from_numpy(arr.reshape(-1))
        self.dist.all_reduce(t)

    def wait(self) -> None:
        pass

    def close(self) -> None:
        pass


class DataParallelTrainer:
    _created = 0

    def __init__(self, model: Module, loss_fn: Loss, lr: float, comm: Optional[Communicator] = None,
                 pred_map: Optional[Tuple[float, float]] = None, squeeze_dim: Optional[int] = None,
                 overlap: bool = True) -> None:
        """`overlap=False` is the measurement baseline (tools/dp_timeline.py): every gradient is exchanged only after
        the whole backward pass has been queued and all parameters are updated by the one fused SGD launch - what
        round 1 effectively did (its hooks fired after the last backward GEMM)."""
        self.model, self.loss_fn = model, loss_fn
        self.overlap = overlap
        self._deferred = []
        self.comm = comm
        self.world = comm.world if comm is not None else 1
        self.loss_fn.world = self.world            # mean over the GLOBAL batch
        self.optimizer = SGD(model.parameters, lr=lr)
        self.pred_map, self.squeeze_dim = pred_map, squeeze_dim
        self.exchanged_bytes = 0
        self.steps = 0
        self.hook_order = []          # id(param) in the order the gradients became final (last step)
        self._updated = set()         # parameters already updated behind their all-reduce in this step
        # names of the parameters for the peer-memory exchange: (trainer number) << 32 | index in model.parameters().
        # Trainers with a communicator are created in the same order on every rank, so the keys agree across ranks.
        self._trainer_no = 0
        if comm is not None:
            DataParallelTrainer._created += 1
            self._trainer_no = DataParallelTrainer._created
        self._keys = {}

    def _on_grad_final(self, param: Tensor) -> None:
        g = param._grad
        if g is None or g is _LAZY:
            return
        if not getattr(g, "is_dense", True):
            g = param.grad
        if isinstance(g, np.ndarray) and not g.flags.c_contiguous:
            g = np.ascontiguousarray(g)
            param._grad = g
        self.hook_order.append(id(param))
        if not self.overlap:
            self._deferred.append(g)
            self.exchanged_bytes += g.nbytes
            return
        fused_update = getattr(self.comm, "allreduce_sgd_", None)
        if fused_update is not None and param.device == Device.CUDA and param.data.is_dense and g.size == param.data.size:
            if getattr(self.comm, "p2p", False):
                fused_update(param.data, g, self.optimizer.lr, self._keys.get(id(param)))
            else:
                fused_update(param.data, g, self.optimizer.lr)
            self._updated.add(id(param))
        else:
            self.comm.allreduce_(g)
        self.exchanged_bytes += g.nbytes

    def step(self, x: Tensor, target: Tensor) -> Tensor:
        """One data-parallel step on this rank's shard.  Returns the LOCAL partial loss (its
        sum over ranks is the global mean loss)."""
        model = self.model
        model.zero_grad()
        pred = model(x)
        if self.squeeze_dim is not None:
            pred = pred.squeeze(self.squeeze_dim)
        if self.pred_map is not None:
            pred = pred * self.pred_map[0] + self.pred_map[1]
        loss = self.loss_fn(pred, target)
        if self.comm is not None and self.world > 1:
            # Parameter objects are replaced on the first forward (Q6): hook the current ones
            self._keys = {}
            for i, p in enumerate(model.parameters()):
                p._grad_hook = self._on_grad_final
                self._keys[id(p)] = (self._trainer_no << 32) | i
            self.hook_order = []
            self._updated = set()
        loss.backward()
        if self.comm is not None and self.world > 1:
            for g in self._deferred:
                self.comm.allreduce_(g)
            self._deferred = []
            self.comm.wait()
        self.optimizer.step(skip=self._updated)
        self.steps += 1
        return loss

    # -- resume: every rank holds the same replica, so rank 0 writes, all ranks read
    def save_checkpoint(self, path) -> None:
        from .checkpoint import save_checkpoint
        if self.comm is None or self.comm.rank == 0:
            save_checkpoint(path, self.model, step=self.steps, lr=self.optimizer.lr, extra={"world": self.world})

    def load_checkpoint(self, path) -> dict:
        from .checkpoint import load_checkpoint
        header = load_checkpoint(path, self.model)
        self.steps = int(header.get("step", 0))
        if header.get("lr") is not None:
            self.optimizer.lr = float(header["lr"])
        return header

    def global_loss(self, local_loss: Tensor) -> float:
        """Sum of the rank partial losses (host value; synchronises)."""
        if self.comm is None or self.world == 1:
            return float(local_loss.item())
        if local_loss.device == Device.CPU:
            buf = np.array(local_loss.data, dtype=np.float32).reshape((1,))
            self.comm.allreduce_(buf)
            return float(buf[0])
        buf = local_loss.data.copy().reshape((1,))
        self.comm.allreduce_(buf)
        self.comm.wait()
        return float(buf.get()[0])
