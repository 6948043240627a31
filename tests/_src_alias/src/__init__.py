"""Test shim: a package literally named `src` that aliases itself to microtorch_b200, so the
reference's own unit tests (which do `from src.tensor import Tensor`) run against this repo."""
__microtorch_b200_alias__ = True
import microtorch_b200.compat as _compat

_compat.install_as_src(__name__)
