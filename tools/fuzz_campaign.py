#!/usr/bin/env python
"""A one-off, larger random op-graph campaign on the CUDA path (run under gpurun): fresh seeds that no fixture holds.

    python tools/fuzz_campaign.py [--start 200000] [--count 6000] [--large 600] > profiles/rNN_fuzz_campaign.json

Every program of tests/fuzz_graphs.py from the seed range is run on this package's NumPy path (float32 and float64) and
on the CUDA path; programs whose float32 NumPy result is more than 3e-6 (norm-wise) from the float64 one are counted as
ill-conditioned and skipped, the rest must agree between CUDA and NumPy within the tolerance of tests/test_fuzz_gpu.py
(5e-5, 1e-4 with matmul in the large variant).  Prints one JSON summary; mismatching programs are listed with their text.
"""
import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--start", type=int, default=200000)
    ap.add_argument("--count", type=int, default=6000)
    ap.add_argument("--large", type=int, default=600)
    args = ap.parse_args()
    import fuzz_graphs as F
    from microtorch_b200 import _capi
    _capi.require_device(0)
    from microtorch_b200.tensor import Tensor
    from microtorch_b200.tensor.device import DType
    out = {"start": args.start}
    for tag, n, kw in (("small", args.count, {}), ("large", args.large, F.LARGE)):
        t0 = time.time()
        ok = ill = 0
        bad, ops = [], {}
        for seed in range(args.start, args.start + n):
            prog = F.make_program(seed, **kw)
            ref = F.run_program(prog, Tensor, "cpu")
            if not F.well_conditioned(prog, Tensor, ref, 3e-6, DType.FLOAT64):
                ill += 1
                continue
            tol = 1e-4 if (kw and any(op == "matmul" for op, _, _ in prog.ops)) else 5e-5
            try:
                got = F.run_program(prog, Tensor, "cuda")
                F.compare(got, ref, tol, f"seed {seed}")
                ok += 1
                for op, _, _ in prog.ops:
                    ops[op] = ops.get(op, 0) + 1
            except Exception as e:                      # a mismatch OR an exception of the CUDA path: both are findings
                bad.append({"seed": seed, "error": f"{type(e).__name__}: {str(e)[:300]}", "program": prog.describe()})
        out[tag] = {"programs": n, "agree": ok, "ill_conditioned_skipped": ill, "findings": bad, "seconds": round(time.time() - t0, 1),
                    "ops_executed": dict(sorted(ops.items()))}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
