#!/usr/bin/env python
"""Bring-up of the tcgen05 "mix" mode (TF32 hi.hi + two BF16 correction products): accuracy vs float64 and time vs the
3xTF32 mode on forward-shaped (both operands K-major) Linear GEMMs.  Run under gpurun; one JSON line per case."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def main():
    from microtorch_b200 import _capi
    _capi.require_device(0)
    lib = _capi.lib()
    from microtorch_b200.cuda.array import DeviceArray
    lib.mt_gemm_tc_set_mix.argtypes = [C.c_int]
    e0, e1 = C.c_void_p(), C.c_void_p()
    lib.mt_event_create(C.byref(e0))
    lib.mt_event_create(C.byref(e1))
    rng = np.random.default_rng(0)

    def timed(fn, iters):
        for _ in range(3):
            fn()
        lib.mt_sync()
        lib.mt_event_record(e0)
        for _ in range(iters):
            fn()
        lib.mt_event_record(e1)
        lib.mt_event_sync(e1)
        ms = C.c_float()
        lib.mt_event_elapsed_ms(e0, e1, C.byref(ms))
        return ms.value / iters

    lib.mt_gemm_set_engine(1)
    for (M, N, K, check) in ((512, 256, 64, True), (2048, 512, 512, True), (4096, 1024, 4096, True), (1000, 260, 100, True),
                             (65536, 4096, 1024, False), (65536, 4096, 4096, False)):
        X = np.tanh(rng.standard_normal((M, K))).astype(np.float32)
        W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
        b = rng.standard_normal(N).astype(np.float32)
        Xd, Wd, bd, Y = DeviceArray.from_host(X), DeviceArray.from_host(W), DeviceArray.from_host(b), DeviceArray.empty((M, N))
        ref = (X.astype(np.float64) @ W.astype(np.float64).T + b) if check else None
        rec = {"M": M, "N": N, "K": K}
        for mix in (0, 1):
            lib.mt_gemm_tc_set_mix(mix)
            fn = lambda: _capi.check(lib.mt_linear_fwd_f32(Y.ptr, Xd.ptr, Wd.ptr, bd.ptr, M, N, K, 0))
            fn()
            if check:
                got = Y.get().astype(np.float64)
                rec[f"err_mix{mix}"] = float(np.abs(got - ref).max() / np.abs(ref).max())
            ms = timed(fn, 5 if M >= 65536 else 30)
            rec[f"ms_mix{mix}"] = round(ms, 4)
            rec[f"tflops_mix{mix}"] = round(2.0 * M * N * K / ms / 1e9, 1)
        lib.mt_gemm_tc_set_mix(0)
        print(json.dumps(rec), flush=True)
    # backward: dW = dZ^T X (both operands MN-major), dX = dZ W (B MN-major), with the accumulate / tanh' epilogues
    for (M, N, K, check) in ((2048, 512, 512, True), (4096, 1024, 1028, True), (1000, 260, 100, True),
                             (65536, 4096, 1024, False), (65536, 4096, 4096, False)):
        X = np.tanh(rng.standard_normal((M, K))).astype(np.float32)
        W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
        dZ = rng.standard_normal((M, N)).astype(np.float32)
        Xd, Wd, dZd = DeviceArray.from_host(X), DeviceArray.from_host(W), DeviceArray.from_host(dZ)
        dX, dW = DeviceArray.empty((M, K)), DeviceArray.empty((N, K))
        rec = {"op": "bwd", "M": M, "N": N, "K": K}
        if check:
            dWr = dZ.astype(np.float64).T @ X.astype(np.float64)
            dXr = (dZ.astype(np.float64) @ W.astype(np.float64)) * (1 - X.astype(np.float64) ** 2)
        for mix in (0, 1):
            lib.mt_gemm_tc_set_mix(mix)
            fn = lambda: _capi.check(lib.mt_linear_bwd_f32(dX.ptr, dW.ptr, dZd.ptr, None, Xd.ptr, Wd.ptr, M, N, K, 2))
            fn()
            if check:
                rec[f"err_dW_mix{mix}"] = float(np.abs(dW.get() - dWr).max() / np.abs(dWr).max())
                rec[f"err_dX_mix{mix}"] = float(np.abs(dX.get() - dXr).max() / np.abs(dXr).max())
            ms = timed(fn, 3 if M >= 65536 else 30)
            fn_dw = lambda: _capi.check(lib.mt_linear_bwd_f32(None, dW.ptr, dZd.ptr, None, Xd.ptr, Wd.ptr, M, N, K, 0))
            ms_dw = timed(fn_dw, 3 if M >= 65536 else 30)
            rec[f"ms_mix{mix}"] = round(ms, 4)
            rec[f"ms_dW_mix{mix}"] = round(ms_dw, 4)
            rec[f"ms_dX_mix{mix}"] = round(ms - ms_dw, 4)
            rec[f"tflops_mix{mix}"] = round(4.0 * M * N * K / ms / 1e9, 1)
        lib.mt_gemm_tc_set_mix(0)
        print(json.dumps(rec), flush=True)
    lib.mt_gemm_set_engine(-1)


if __name__ == "__main__":
    main()
