"""CPU oracle for the MicroTorch training-step hot path.  TEST INFRASTRUCTURE ONLY.

This package is a NumPy restatement of the algorithm the reference
(nickovchinnikov/microtorch, NumPy backend) runs for one MLP training step:
broadcasting elementwise forward/backward, sum/mean, Linear (matmul + in-place
bias), Tanh, MSE / "CrossEntropy" / BCE losses, recursive backward, SGD.
Every function cites the reference file:line it follows (paths relative to
/root/reference).

Rules (enforced by tests/test_no_oracle_in_product.py):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
    ``cpu_baseline`` / ``--impl reference`` legs may import this package;
  * nothing under ``microtorch_b200/`` imports it - the product path fails
    loudly when the CUDA extension is missing, it never falls back to this.

Parity pin: the oracle is checked bit-for-bit against the reference itself,
imported from /root/reference in the build container (tests/test_oracle_vs_reference.py),
against golden vectors generated from the reference (tests/golden/*.npz, made by
tests/golden/make_golden.py) and against the reference's own published loss stream
(demo.ipynb output, steps 0..100 -> tests/golden/c1a_notebook_losses.json).
"""

from .tape import Var, unbroadcast  # noqa: F401
from . import nnref, workloads  # noqa: F401
