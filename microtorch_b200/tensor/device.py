"""Device / dtype dispatch - the seam the B200 backend plugs into.

Mirrors the interface of the reference's src/tensor/device.py (same names, argument meaning
and error behaviour): `Device`, `DType`, `CUDA_AVAILABLE`, `check_cuda`, `_device`,
`_tensor`, `get_dtype`.  The one change is what sits behind Device.CUDA: where the reference
returns the `cupy` module (device.py:74-84), this returns `microtorch_b200.cuda`, the array
namespace backed by libmicrotorch_b200.so.  CUDA_AVAILABLE (device.py:7-11) is "the library
loads and a device is visible"; there is no fallback if it is not.
"""

from __future__ import annotations

from enum import Enum
from typing import Union

import numpy as np

from .. import cuda as _cuda_ns

CUDA_AVAILABLE: bool = _cuda_ns.is_available()

Scalar = Union[int, float]
Vector = Union[np.ndarray, "_cuda_ns.DeviceArray"]


class DType(Enum):
    """Supported element types (the CUDA backend computes in FLOAT32 only)."""
    FLOAT64 = "float64"
    FLOAT32 = "float32"
    INT64 = "int64"
    INT32 = "int32"
    INT16 = "int16"
    INT8 = "int8"


class Device(str, Enum):
    CPU = "cpu"
    CUDA = "cuda"


def check_cuda() -> bool:
    return CUDA_AVAILABLE


def _device(device: Union[Device, str]) -> Device:
    """Validate a device spec.  ImportError for an unknown name or for "cuda" without a
    usable B200 backend (the reference raises ImportError here too, device.py:65-69)."""
    if isinstance(device, Device):
        return device
    name = device.lower() if isinstance(device, str) else device
    if name not in (Device.CPU.value, Device.CUDA.value):
        raise ImportError("Device must be either 'cpu' or 'cuda'.")
    if name == Device.CUDA.value and not CUDA_AVAILABLE:
        raise ImportError("CUDA operations requested but the B200 backend (libmicrotorch_b200.so + a visible "
                          "GPU) is not available!")
    return Device(name)


def _tensor(device: Union[Device, str] = Device.CPU):
    """The array namespace for `device`: NumPy on CPU, microtorch_b200.cuda on CUDA."""
    return _cuda_ns if _device(device) == Device.CUDA else np


_DTYPE_NAMES = {d: d.value for d in DType}


def get_dtype(device: Device, dtype: DType):
    """The namespace's scalar type for a DType (np.float32, ...)."""
    if dtype not in _DTYPE_NAMES:
        raise ValueError(f"Unsupported dtype '{dtype}' for device '{device}'")
    return getattr(_tensor(device), _DTYPE_NAMES[dtype])
