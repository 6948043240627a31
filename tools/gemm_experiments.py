#!/usr/bin/env python
"""Timing experiments on the tcgen05 GEMM (run under gpurun): which stage bounds the pipeline?
Flags: 1 = converters skip their work, 2 = only the hi*hi MMA is issued, 4 = no MMA issued.
Results with flags != 0 are numerically wrong on purpose; only the time matters."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np  # noqa: E402
from microtorch_b200 import _capi  # noqa: E402
from capi_util import Dev  # noqa: E402

lib = _capi.lib()
_capi.require_device(0)
for f in ("mt_gemm_tc_set_dbg", "mt_gemm_tc_set_pair", "mt_gemm_tc_set_chunk_kb"):
    getattr(lib, f).argtypes = [C.c_int]
M, N, K = 65536, 4096, 4096
rng = np.random.default_rng(0)
X = Dev(rng.standard_normal((M, K), dtype=np.float32))
W = Dev((rng.standard_normal((N, K), dtype=np.float32) * 0.05).astype(np.float32))
dY = Dev(rng.standard_normal((M, N), dtype=np.float32))
Y = Dev(shape=(M, N))
dW = Dev(shape=(N, K))
dX = Dev(shape=(M, K))
lib.mt_gemm_set_engine(1)
e0, e1 = C.c_void_p(), C.c_void_p()
lib.mt_event_create(C.byref(e0))
lib.mt_event_create(C.byref(e1))


def time_it(fn, iters=4):
    fn()
    ts = []
    for _ in range(iters):
        lib.mt_event_record(e0)
        fn()
        lib.mt_event_record(e1)
        lib.mt_event_sync(e1)
        ms = C.c_float()
        lib.mt_event_elapsed_ms(e0, e1, C.byref(ms))
        ts.append(ms.value)
    return min(ts), sorted(ts)[len(ts) // 2]


def fwd():
    _capi.check(lib.mt_linear_fwd_f32(Y.ptr, X.ptr, W.ptr, None, M, N, K, 0))


def dw():
    _capi.check(lib.mt_linear_bwd_f32(None, dW.ptr, dY.ptr, None, X.ptr, W.ptr, M, N, K, 0))


flops = 2.0 * M * N * K
lib.mt_gemm_tc_set_rings.argtypes = [C.c_int, C.c_int]
mode = sys.argv[1] if len(sys.argv) > 1 else "rings"
lib.mt_gemm_tc_set_splitk.argtypes = [C.c_int]
Yact = Dev(np.tanh(rng.standard_normal((M, N), dtype=np.float32)).astype(np.float32))


def dw_pro():
    _capi.check(lib.mt_linear_bwd_f32(None, dW.ptr, dY.ptr, Yact.ptr, X.ptr, W.ptr, M, N, K, 0))


def bwd_pro():   # dW + dX, both with the fused tanh' prologue
    _capi.check(lib.mt_linear_bwd_f32(dX.ptr, dW.ptr, dY.ptr, Yact.ptr, X.ptr, W.ptr, M, N, K, 0))


if mode == "final":
    lib.mt_gemm_tc_set_pair(1)
    lib.mt_gemm_tc_set_chunk_kb(8)
    lib.mt_gemm_tc_set_dbg(0)
    for splitk in (1, 0):
        for raw, lo in ((0, 2), (0, 3)):
            lib.mt_gemm_tc_set_splitk(splitk)
            lib.mt_gemm_tc_set_rings(raw, lo)
            for name, fn, nf in (("fwd", fwd, 1), ("dw", dw, 1), ("dw_pro", dw_pro, 1), ("dw_pro+dx_pro", bwd_pro, 2)):
                mn, med = time_it(fn)
                print(json.dumps({"op": name, "splitk": splitk, "raw": raw, "lo": lo, "ms_min": mn, "ms_med": med,
                                  "tflops_logical_at_min": nf * flops / mn / 1e9}), flush=True)
    sys.exit(0)
if mode == "rings":
    # ring-depth sweep: (raw, lo); raw=0 -> whatever fits
    for pair, rings in ((1, [(0, 2), (0, 3), (0, 4), (3, 3), (2, 5)]), (0, [(0, 1), (0, 2)])):
        for raw, lo in rings:
            for dbg in (0, 4, 5):
                lib.mt_gemm_tc_set_pair(pair)
                lib.mt_gemm_tc_set_chunk_kb(8)
                lib.mt_gemm_tc_set_rings(raw, lo)
                lib.mt_gemm_tc_set_dbg(dbg)
                for name, fn in (("fwd", fwd), ("dw", dw)):
                    mn, med = time_it(fn)
                    print(json.dumps({"op": name, "pair": pair, "raw": raw, "lo": lo, "dbg": dbg, "ms_min": mn, "ms_med": med,
                                      "tflops_logical_at_min": flops / mn / 1e9}), flush=True)
else:
    for pair in (1, 0):
        for dbg in (0, 1, 2, 3, 4, 5):
            lib.mt_gemm_tc_set_pair(pair)
            lib.mt_gemm_tc_set_chunk_kb(8)
            lib.mt_gemm_tc_set_dbg(dbg)
            for name, fn in (("fwd", fwd), ("dw", dw)):
                mn, med = time_it(fn)
                print(json.dumps({"op": name, "pair": pair, "dbg": dbg, "ms_min": mn, "ms_med": med,
                                  "tflops_logical_at_min": flops / mn / 1e9}), flush=True)
