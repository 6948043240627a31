"""Drop-in aliasing: make `import src...` (the reference's package name) resolve to this package.

    import microtorch_b200.compat as compat
    compat.install_as_src()
    from src.tensor import Tensor            # -> microtorch_b200.tensor.Tensor
    from src.nn.optimizer import SGD

Every module path of the reference that user code or its tests import has a counterpart of
the same name here (src.tensor.{device,tensor,types,ops.*}, src.nn.{module,param,sequential,
activation,loss,optimizer,layer.linear}).
"""

from __future__ import annotations

import importlib
import sys

_MODULES = [
    "tensor", "tensor.device", "tensor.tensor", "tensor.types", "tensor.backend",
    "tensor.ops", "tensor.ops.base", "tensor.ops.overload", "tensor.ops.math", "tensor.ops.reduce",
    "tensor.ops.elementwise",
    "nn", "nn.module", "nn.param", "nn.sequential", "nn.activation", "nn.loss", "nn.optimizer",
    "nn.layer", "nn.layer.linear",
]


def install_as_src(name: str = "src") -> None:
    """Register `name` and its submodules in sys.modules as aliases of microtorch_b200."""
    root = importlib.import_module("microtorch_b200")
    existing = sys.modules.get(name)
    if existing is not None and existing is not root and getattr(existing, "__microtorch_b200_alias__", False) is False \
            and getattr(existing, "__file__", None) and "microtorch_b200" not in str(existing.__file__) \
            and not str(existing.__file__).endswith("_src_alias/src/__init__.py"):
        raise ImportError(f"a different package named '{name}' is already imported from {existing.__file__}")
    sys.modules[name] = root if existing is None else existing
    for sub in _MODULES:
        mod = importlib.import_module(f"microtorch_b200.{sub}")
        sys.modules[f"{name}.{sub}"] = mod
        parent, _, leaf = sub.rpartition(".")
        holder = sys.modules[f"{name}.{parent}"] if parent else sys.modules[name]
        setattr(holder, leaf, mod)
