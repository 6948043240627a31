"""Random op-graph programs (tests/fuzz_graphs.py) on the CUDA path: every op of the reference's Tensor surface in
random compositions - broadcast layouts, transposed / sliced / re-shaped views as operands, tuple-axis reductions,
integer-array indexing, 1-D and 2-D matmul - forward values, loss and every leaf gradient.

Compared against (a) the fixtures made by running the REFERENCE (tests/golden/fuzz_ref.npz) and (b) this package's
NumPy path (pinned to the live reference by tests/test_fuzz_cpu.py) on every program the generating script found robust
(reference and NumPy path both within 3e-6 of float64).  Tolerance: 5e-5 norm-wise per array - a dozen composed ops, each
within the north_star's 1e-5; a structural fault (a stride, a broadcast axis, a missed accumulation) is O(1)."""
import os

import numpy as np
import pytest

import fuzz_graphs as F
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

TOL = 5e-5


@pytest.fixture(scope="module")
def Tensor():
    from microtorch_b200 import _capi
    _capi.require_device()
    from microtorch_b200.tensor import Tensor
    return Tensor


def test_cuda_path_matches_reference_fixtures(Tensor):
    fx = np.load(os.path.join(GOLDEN, "fuzz_ref.npz"))
    seeds = sorted(int(k.split("_")[1]) for k in fx.files if k.startswith("digest_"))
    assert len(seeds) >= 100
    for seed in seeds:
        prog = F.make_program(seed)
        ref = {"values": [fx[f"{seed}_v{i}"] for i in range(len(prog.terminals))],
               "loss": float(fx[f"{seed}_loss"]),
               "grads": [fx[f"{seed}_g{i}"] for i in range(len(prog.leaves))]}
        got = F.run_program(prog, Tensor, "cuda")
        F.compare(got, ref, TOL, f"seed {seed}\n{prog.describe()}")


def test_cuda_path_matches_cpu_path_on_robust_programs(Tensor):
    fx = np.load(os.path.join(GOLDEN, "fuzz_ref.npz"))
    seeds = [int(s) for s in fx["robust_seeds"][:600]]
    for seed in seeds:
        prog = F.make_program(seed)
        ref = F.run_program(prog, Tensor, "cpu")
        got = F.run_program(prog, Tensor, "cuda")
        F.compare(got, ref, TOL, f"seed {seed}\n{prog.describe()}")


def test_cuda_path_matches_cpu_path_on_large_programs(Tensor):
    """The large variant (extents 1 ... 1024 with odd ones in between, up to 4 M elements per node): the vectorised and
    grid-stride elementwise paths, split reductions, strided copies of big views and - through matmul - the SIMT, mid-size
    and persistent tensor-core engines, all inside one autograd graph.  1e-4 norm-wise where a program multiplies
    matrices (the north_star's GEMM tolerance), 5e-5 otherwise."""
    fx = np.load(os.path.join(GOLDEN, "fuzz_ref.npz"))
    seeds = [int(s) for s in fx["robust_seeds_large"]]
    assert len(seeds) >= 120
    for seed in seeds:
        prog = F.make_program(seed, **F.LARGE)
        ref = F.run_program(prog, Tensor, "cpu")
        got = F.run_program(prog, Tensor, "cuda")
        tol = 1e-4 if any(op == "matmul" for op, _, _ in prog.ops) else TOL
        F.compare(got, ref, tol, f"large seed {seed}\n{prog.describe()}")


def test_cuda_path_matches_cpu_path_on_select_family_programs(Tensor):
    """The select-family variant: comparisons feeding where / clip / threshold / masked_fill / sign (reference
    tensor.py:467-503 - on CuPy the reference builds those masks and constants on the CPU, Q8, so this family never ran on
    its CUDA path; here it runs on the device) mixed into the same random graphs.  The generator keeps every compared
    element 1e-4 clear of its threshold, so which branch wins does not hang on an ulp."""
    fx = np.load(os.path.join(GOLDEN, "fuzz_ref.npz"))
    seeds = [int(s) for s in fx["robust_seeds_select"][:400]]
    assert len(seeds) >= 350
    for seed in seeds:
        prog = F.make_program(seed, select=True)
        ref = F.run_program(prog, Tensor, "cpu")
        got = F.run_program(prog, Tensor, "cuda")
        F.compare(got, ref, TOL, f"select seed {seed}\n{prog.describe()}")
