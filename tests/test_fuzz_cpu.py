"""Random op-graph programs (tests/fuzz_graphs.py): this package's NumPy path against the reference.

* everywhere: against the fixtures made by running the reference (tests/golden/make_fuzz_golden.py -> fuzz_ref.npz),
  and the generator itself against the program digests stored with them;
* in the build container: against the live reference on 1500 fresh programs (forward values, loss and every leaf
  gradient; 2e-5 norm-wise on programs that a float64 evaluation shows to be free of cancellation)."""
import hashlib
import importlib.util
import os

import numpy as np
import pytest

import fuzz_graphs as F
from conftest import GOLDEN, needs_reference


def fixture():
    return np.load(os.path.join(GOLDEN, "fuzz_ref.npz"))


def stored_seeds(fx):
    return sorted(int(k.split("_")[1]) for k in fx.files if k.startswith("digest_"))      # ("large_" / "select_digest_*" do not match)


def reference_result(fx, seed, prog):
    return {"values": [fx[f"{seed}_v{i}"] for i in range(len(prog.terminals))],
            "loss": float(fx[f"{seed}_loss"]),
            "grads": [fx[f"{seed}_g{i}"] for i in range(len(prog.leaves))]}


def test_generator_reproduces_the_stored_programs():
    fx = fixture()
    seeds = stored_seeds(fx)
    assert len(seeds) >= 100
    for seed in seeds:
        prog = F.make_program(seed)
        assert hashlib.sha1(prog.describe().encode()).hexdigest() == str(fx[f"digest_{seed}"]), f"program {seed} changed"
    for seed in fx["robust_seeds_large"][:20]:
        prog = F.make_program(int(seed), **F.LARGE)
        assert hashlib.sha1(prog.describe().encode()).hexdigest() == str(fx[f"large_digest_{int(seed)}"]), f"large program {seed} changed"
    for key in [k for k in fx.files if k.startswith("select_digest_")]:
        seed = int(key.split("_")[2])
        prog = F.make_program(seed, select=True)
        assert hashlib.sha1(prog.describe().encode()).hexdigest() == str(fx[key]), f"select program {seed} changed"
    sel_ops = {op for seed in fx["robust_seeds_select"][:300] for op, _, _ in F.make_program(int(seed), select=True).ops}
    assert {n for n, _ in F._SELECT_OPS} <= sel_ops
    ops = {op for seed in fx["robust_seeds"][:400] for op, _, _ in F.make_program(int(seed)).ops}
    assert ops == set(F._NAMES), f"ops never generated: {set(F._NAMES) - ops}"


def test_cpu_path_matches_reference_fixtures():
    from microtorch_b200.tensor import Tensor
    fx = fixture()
    for seed in stored_seeds(fx):
        prog = F.make_program(seed)
        got = F.run_program(prog, Tensor, "cpu")
        F.compare(got, reference_result(fx, seed, prog), 2e-5, f"seed {seed}\n{prog.describe()}")


@needs_reference
def test_cpu_path_matches_live_reference_on_fresh_programs():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    mg.import_reference()
    import src.tensor.tensor as ref_mod
    from microtorch_b200.tensor import Tensor
    from microtorch_b200.tensor.device import DType
    compared = 0
    for seed in range(100000, 101500):
        prog = F.make_program(seed)
        ref = F.run_program(prog, ref_mod.Tensor, "cpu")         # the generator avoids what the reference cannot run
        got = F.run_program(prog, Tensor, "cpu")
        if not (F.well_conditioned(prog, Tensor, ref, 3e-6, DType.FLOAT64)
                and F.well_conditioned(prog, Tensor, got, 3e-6, DType.FLOAT64)):
            continue
        F.compare(got, ref, 2e-5, f"seed {seed}\n{prog.describe()}")
        compared += 1
    assert compared >= 1400, compared
    compared = 0
    for seed in range(100000, 101000):                          # the select-family variant
        prog = F.make_program(seed, select=True)
        ref = F.run_program(prog, ref_mod.Tensor, "cpu")
        got = F.run_program(prog, Tensor, "cpu")
        if not (F.well_conditioned(prog, Tensor, ref, 3e-6, DType.FLOAT64)
                and F.well_conditioned(prog, Tensor, got, 3e-6, DType.FLOAT64)):
            continue
        F.compare(got, ref, 2e-5, f"select seed {seed}\n{prog.describe()}")
        compared += 1
    assert compared >= 930, compared
