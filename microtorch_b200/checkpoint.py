"""Checkpoints for the training path.

The reference persists single tensors only (`Tensor.save/load`, a pickle of the host arrays,
src/tensor/tensor.py:312-365 - mirrored in tensor/tensor.py here, device tensors included).  A model
has no on-disk form there, so this module adds the small missing piece the trainer needs to resume:

    state_dict(module)              {"modules.0.weight": ndarray, ...}   host copies, enumeration order
    load_state_dict(module, state)  writes INTO the existing Parameter buffers (device or host), so
                                    optimizers / trainers / captured graphs that hold them stay valid
    save_checkpoint / load_checkpoint   one .npz (no pickle), written atomically, plus step / lr / extras

Names follow the attribute path that `Module.parameters()` walks (the same order SGD sees), so a state
saved from a CPU model loads into a CUDA model and back.  In data-parallel runs every rank holds the same
replica: rank 0 writes, every rank reads (trainer.DataParallelTrainer.save_checkpoint / load_checkpoint).
"""

from __future__ import annotations

import json
import os
import tempfile
from pathlib import Path
from typing import Any, Dict, Iterator, Optional, Tuple, Union

import numpy as np

from .nn.module import Module
from .nn.param import Parameter
from .tensor.device import Device

FORMAT = "microtorch_b200.checkpoint.v1"


def named_parameters(module: Module, prefix: str = "") -> Iterator[Tuple[str, Parameter]]:
    """(attribute path, Parameter) in exactly the order of `module.parameters()`."""
    for name, value in module.__dict__.items():
        path = f"{prefix}{name}"
        if isinstance(value, Parameter):
            yield path, value
        elif isinstance(value, Module):
            yield from named_parameters(value, path + ".")
        elif isinstance(value, (tuple, list)):
            for i, child in enumerate(value):
                if isinstance(child, Module):
                    yield from named_parameters(child, f"{path}.{i}.")


def _check_enumeration(module: Module, named) -> None:
    listed = [id(p) for p in module.parameters()]
    if listed != [id(p) for _, p in named]:
        raise RuntimeError("named_parameters() does not match Module.parameters(): the module overrides parameters() "
                           "in a way this walker does not know")


def state_dict(module: Module) -> Dict[str, np.ndarray]:
    named = list(named_parameters(module))
    _check_enumeration(module, named)
    out: Dict[str, np.ndarray] = {}
    for name, p in named:
        out[name] = p.data.get() if p.device == Device.CUDA else np.array(p.data, copy=True)
    return out


def load_state_dict(module: Module, state: Dict[str, np.ndarray], strict: bool = True) -> None:
    named = list(named_parameters(module))
    _check_enumeration(module, named)
    have = {name for name, _ in named}
    missing, unexpected = sorted(have - set(state)), sorted(set(state) - have)
    if strict and (missing or unexpected):
        raise KeyError(f"state does not match the module: missing {missing}, unexpected {unexpected}")
    for name, p in named:
        if name not in state:
            continue
        value = np.asarray(state[name])
        if tuple(value.shape) != tuple(p.shape):
            raise ValueError(f"{name}: shape {tuple(value.shape)} does not match the parameter's {tuple(p.shape)}")
        value = np.ascontiguousarray(value, dtype=np.float32) if p.device == Device.CUDA else value.astype(p.data.dtype)
        p.data[...] = value          # into the existing buffer: H2D + scatter on CUDA, a host copy on CPU


def save_checkpoint(path: Union[str, Path], module: Module, step: int = 0, lr: Optional[float] = None,
                    extra: Optional[Dict[str, Any]] = None) -> None:
    """Write `path` (.npz) atomically: parameters by name plus a JSON header (format, step, lr, extra)."""
    path = Path(path)
    arrays = state_dict(module)
    header = {"format": FORMAT, "step": int(step), "lr": None if lr is None else float(lr), "extra": extra or {},
              "names": list(arrays)}
    fd, tmp = tempfile.mkstemp(dir=str(path.parent), prefix=path.name + ".", suffix=".tmp")
    try:
        with os.fdopen(fd, "wb") as f:
            np.savez(f, __header__=np.frombuffer(json.dumps(header).encode(), dtype=np.uint8),
                     **{f"p{i}": a for i, a in enumerate(arrays.values())})
        os.replace(tmp, path)
    except BaseException:
        if os.path.exists(tmp):
            os.unlink(tmp)
        raise


def read_checkpoint(path: Union[str, Path]) -> Tuple[Dict[str, np.ndarray], Dict[str, Any]]:
    path = Path(path)
    if not path.exists():
        raise FileNotFoundError(f"No checkpoint found at {path}")
    try:
        with np.load(path, allow_pickle=False) as z:
            header = json.loads(bytes(z["__header__"]).decode())
            if header.get("format") != FORMAT:
                raise ValueError(f"unknown checkpoint format {header.get('format')!r}")
            arrays = {name: z[f"p{i}"] for i, name in enumerate(header["names"])}
    except (KeyError, OSError, json.JSONDecodeError, ValueError) as e:
        raise ValueError(f"Invalid checkpoint file at {path}: {e}") from e
    return arrays, header


def load_checkpoint(path: Union[str, Path], module: Module, strict: bool = True) -> Dict[str, Any]:
    """Load parameters into `module` in place; returns the header (step, lr, extra)."""
    arrays, header = read_checkpoint(path)
    load_state_dict(module, arrays, strict=strict)
    return header
