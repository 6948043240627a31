#!/usr/bin/env python
"""Tile order / L2-hint experiment on the tcgen05 GEMMs of BASELINE config #2 (run under gpurun).

VERDICT r1: the persistent loop walked tiles n-fastest and re-fetched operand panels every wave - DRAM traffic 3.0x
the algorithmic bytes on average (dW 6.2x).  This script times the five big GEMMs of a C2 step under

    group_n  n-blocks per column group (-1 = row-major = the round-1 order, 0 = library's choice, 4 / 8 explicit)
    hints    bit 0 = L2 eviction-priority hints on the TMA loads, bit 1 = streaming C stores

    python tools/gemm_l2.py                 # CUDA-event timings, JSON lines
    python tools/gemm_l2.py --once          # one launch per (shape, setting): the run to put under ncu
                                            #   ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
"""
import argparse
import ctypes as C
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np  # noqa: E402
from microtorch_b200 import _capi  # noqa: E402
from capi_util import Dev  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--iters", type=int, default=8)
    ap.add_argument("--settings", default="-1:0,8:0,8:2,8:4,8:6,6:2,4:0,4:2")
    ap.add_argument("--rows", type=int, default=65536)
    args = ap.parse_args()
    lib = _capi.lib()
    _capi.require_device(0)
    for f in ("mt_gemm_tc_set_group_n", "mt_gemm_tc_set_hints"):
        getattr(lib, f).argtypes = [C.c_int]
    lib.mt_gemm_set_engine(1)
    B = args.rows
    rng = np.random.default_rng(0)
    X1 = Dev(rng.standard_normal((B, 1024), dtype=np.float32))
    X2 = Dev(np.tanh(rng.standard_normal((B, 4096), dtype=np.float32)))
    W1 = Dev((rng.standard_normal((4096, 1024), dtype=np.float32) * 0.05).astype(np.float32))
    W2 = Dev((rng.standard_normal((4096, 4096), dtype=np.float32) * 0.05).astype(np.float32))
    dZ = Dev(rng.standard_normal((B, 4096), dtype=np.float32))
    Y = Dev(shape=(B, 4096))
    dX = Dev(shape=(B, 4096))
    dW1 = Dev(shape=(4096, 1024))
    dW2 = Dev(shape=(4096, 4096))
    e0, e1 = C.c_void_p(), C.c_void_p()
    lib.mt_event_create(C.byref(e0))
    lib.mt_event_create(C.byref(e1))
    ops = {
        # name: (callable, M, N, K of the GEMM, algorithmic bytes: read A + read B + write C)
        "fwd1_65536x4096x1024": (lambda: lib.mt_linear_fwd_f32(Y.ptr, X1.ptr, W1.ptr, None, B, 4096, 1024, 1),
                                 2.0 * B * 4096 * 1024, 4.0 * (B * 1024 + 4096 * 1024 + B * 4096)),
        "fwd2_65536x4096x4096": (lambda: lib.mt_linear_fwd_f32(Y.ptr, X2.ptr, W2.ptr, None, B, 4096, 4096, 1),
                                 2.0 * B * 4096 * 4096, 4.0 * (B * 4096 + 4096 * 4096 + B * 4096)),
        "dW2_4096x4096x65536": (lambda: lib.mt_linear_bwd_f32(None, dW2.ptr, dZ.ptr, None, X2.ptr, W2.ptr, B, 4096, 4096, 0),
                                2.0 * B * 4096 * 4096, 4.0 * (2 * B * 4096 + 4096 * 4096)),
        "dW1_4096x1024x65536": (lambda: lib.mt_linear_bwd_f32(None, dW1.ptr, dZ.ptr, None, X1.ptr, W1.ptr, B, 4096, 1024, 0),
                                2.0 * B * 4096 * 1024, 4.0 * (B * 4096 + B * 1024 + 4096 * 1024)),
    }

    def dx_only():
        # dX2 = dZ . W2 (+ the tanh' epilogue of the chain fusion reads X2): mt_linear_bwd_f32 always runs dW too, so the
        # bare product goes through the generic matmul entry point (same kernel, operand B MN-major)
        return lib.mt_matmul_f32(dX.ptr, dZ.ptr, W2.ptr, 1, B, 4096, 4096, 0, 4096, 1, 0, 4096, 1)
    ops["dX2_65536x4096x4096"] = (dx_only, 2.0 * B * 4096 * 4096, 4.0 * (2 * B * 4096 + 4096 * 4096))

    settings = [tuple(int(v) for v in st.split(":")) for st in args.settings.split(",")]

    def apply(gn, hints):
        lib.mt_gemm_tc_set_group_n(gn)
        lib.mt_gemm_tc_set_hints(hints)

    if args.once:
        for gn, hints in settings:
            apply(gn, hints)
            for name, (fn, flops, alg) in ops.items():
                _capi.check(fn(), name)
                _capi.check(lib.mt_sync())
                print(json.dumps({"op": name, "group_n": gn, "hints": hints, "alg_bytes": alg}), flush=True)
    else:
        # INTERLEAVED: under the power cap the clock sinks while the GPU warms up, so timing one setting after the other
        # favours whichever comes first; every repetition visits every (setting, op) once instead
        times = {}
        for rep in range(args.iters + 2):
            for gn, hints in settings:
                apply(gn, hints)
                for name, (fn, flops, alg) in ops.items():
                    lib.mt_event_record(e0)
                    _capi.check(fn(), name)
                    lib.mt_event_record(e1)
                    lib.mt_event_sync(e1)
                    ms = C.c_float()
                    lib.mt_event_elapsed_ms(e0, e1, C.byref(ms))
                    if rep >= 2:
                        times.setdefault((gn, hints, name), []).append(ms.value)
        for (gn, hints, name), ts in times.items():
            ts.sort()
            flops, alg = ops[name][1], ops[name][2]
            print(json.dumps({"op": name, "group_n": gn, "hints": hints, "ms_min": ts[0], "ms_med": ts[len(ts) // 2],
                              "tflops_at_med": flops / ts[len(ts) // 2] / 1e9, "alg_gb": alg / 1e9}), flush=True)
    lib.mt_gemm_tc_set_group_n(0)
    lib.mt_gemm_tc_set_hints(3)


if __name__ == "__main__":
    main()
