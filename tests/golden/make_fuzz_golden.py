#!/usr/bin/env python
"""Generate tests/golden/fuzz_ref.npz by RUNNING THE REFERENCE on the random op-graph programs of tests/fuzz_graphs.py.

Build container only (needs /root/reference):   python tests/golden/make_fuzz_golden.py

For seeds 0 .. N_SEEDS-1 a program is kept ("robust") when the reference's float32 result AND this package's NumPy-path
float32 result both sit within 3e-6 (norm-wise) of a float64 evaluation and within 2e-5 of each other - two different
accumulation orders agree, so what a third backend is compared on is arithmetic, not cancellation noise.  Stored:
  robust_seeds        every robust seed (the GPU test runs the CUDA path against the CPU path on all of them)
  digest_<seed>       sha1 of the program text (generator drift shows up as a digest mismatch, not as a numeric one)
  <seed>_loss / <seed>_v<i> / <seed>_g<i>   the REFERENCE's loss, terminal values and leaf gradients for the robust
                      programs with at most MAX_ELEMS_STORED result elements (first N_STORED of them)
  robust_seeds_large / large_digest_<seed>   the same filter over the LARGE variant (fuzz_graphs.LARGE: extents up to
                      1024, up to 4 M elements per node); seeds and program digests only
  robust_seeds_select / select_digest_<seed>  the same for the select-family variant (make_program(select=True))
"""
import hashlib
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.dont_write_bytecode = True

N_SEEDS = 1200
N_SEEDS_LARGE = 160
N_SEEDS_SELECT = 500
N_STORED = 160
MAX_ELEMS_STORED = 3000


def digest(prog) -> str:
    return hashlib.sha1(prog.describe().encode()).hexdigest()


def main():
    import fuzz_graphs as F
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    mg.import_reference()                      # the unmodified reference as `src` + shim S1 (0-d gradients)
    import src.tensor.tensor as ref_mod
    from microtorch_b200.tensor import Tensor
    from microtorch_b200.tensor.device import DType

    out, robust, stored = {}, [], 0
    for seed in range(N_SEEDS):
        prog = F.make_program(seed)
        ref = F.run_program(prog, ref_mod.Tensor, "cpu")
        got = F.run_program(prog, Tensor, "cpu")
        exact = F.run_program(prog, Tensor, "cpu", dtype=DType.FLOAT64)
        try:
            F.compare(ref, exact, 3e-6, "reference vs float64")
            F.compare(got, exact, 3e-6, "package vs float64")
            F.compare(got, ref, 2e-5, "package vs reference")
        except AssertionError:
            continue
        robust.append(seed)
        n_elems = sum(np.size(v) for v in ref["values"]) + sum(np.size(g) for g in ref["grads"])
        if stored < N_STORED and n_elems <= MAX_ELEMS_STORED:
            stored += 1
            out[f"digest_{seed}"] = np.array(digest(prog))
            out[f"{seed}_loss"] = np.float32(ref["loss"])
            for i, v in enumerate(ref["values"]):
                out[f"{seed}_v{i}"] = np.asarray(v, dtype=np.float32)
            for i, g in enumerate(ref["grads"]):
                out[f"{seed}_g{i}"] = np.asarray(g, dtype=np.float32)
    out["robust_seeds"] = np.array(robust, dtype=np.int64)
    large = []
    for seed in range(N_SEEDS_LARGE):          # the large variant: seeds only (the GPU test compares CUDA with the CPU path)
        prog = F.make_program(seed, **F.LARGE)
        ref = F.run_program(prog, ref_mod.Tensor, "cpu")
        got = F.run_program(prog, Tensor, "cpu")
        exact = F.run_program(prog, Tensor, "cpu", dtype=DType.FLOAT64)
        try:
            F.compare(ref, exact, 3e-6, "reference vs float64")
            F.compare(got, exact, 3e-6, "package vs float64")
            F.compare(got, ref, 2e-5, "package vs reference")
        except AssertionError:
            continue
        large.append(seed)
        out[f"large_digest_{seed}"] = np.array(digest(prog))
    out["robust_seeds_large"] = np.array(large, dtype=np.int64)
    sel = []
    for seed in range(N_SEEDS_SELECT):         # the select-family variant (clip / threshold / where / masked_fill / sign): seeds only
        prog = F.make_program(seed, select=True)
        ref = F.run_program(prog, ref_mod.Tensor, "cpu")
        got = F.run_program(prog, Tensor, "cpu")
        exact = F.run_program(prog, Tensor, "cpu", dtype=DType.FLOAT64)
        try:
            F.compare(ref, exact, 3e-6, "reference vs float64")
            F.compare(got, exact, 3e-6, "package vs float64")
            F.compare(got, ref, 2e-5, "package vs reference")
        except AssertionError:
            continue
        sel.append(seed)
        if len(sel) <= 20:
            out[f"select_digest_{seed}"] = np.array(digest(prog))
    out["robust_seeds_select"] = np.array(sel, dtype=np.int64)
    path = os.path.join(HERE, "fuzz_ref.npz")
    np.savez_compressed(path, **out)
    print(f"{len(robust)} robust programs of {N_SEEDS}, {stored} with reference outputs; {len(large)} large of {N_SEEDS_LARGE}, {len(sel)} select of {N_SEEDS_SELECT} -> {path} "
          f"({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
