#!/usr/bin/env python
"""A/B timing of the five tcgen05 GEMMs of one BASELINE config #2 step with the BF16-correction ("mix") mode off / on,
interleaved (off, on, off, on, ...) so that clock drift under the power cap hits both arms alike.  Run under gpurun."""
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
    B = 65536

    def dev(*shape):
        return DeviceArray.from_host(rng.standard_normal(shape).astype(np.float32) * 0.05)

    X0, X1 = dev(B, 1024), dev(B, 4096)
    W1, W2 = dev(4096, 1024), dev(4096, 4096)
    dZ = dev(B, 4096)
    Y = DeviceArray.empty((B, 4096))
    dX = DeviceArray.empty((B, 4096))
    dW1, dW2 = DeviceArray.empty((4096, 1024)), DeviceArray.empty((4096, 4096))
    cases = {
        "fwd_L1": lambda: lib.mt_linear_fwd_f32(Y.ptr, X0.ptr, W1.ptr, None, B, 4096, 1024, 1),
        "fwd_L2": lambda: lib.mt_linear_fwd_f32(Y.ptr, X1.ptr, W2.ptr, None, B, 4096, 4096, 1),
        "dW_L2": lambda: lib.mt_linear_bwd_f32(None, dW2.ptr, dZ.ptr, None, X1.ptr, W2.ptr, B, 4096, 4096, 0),
        "dW_L1": lambda: lib.mt_linear_bwd_f32(None, dW1.ptr, dZ.ptr, None, X0.ptr, W1.ptr, B, 4096, 1024, 0),
        "dWdX_L2": lambda: lib.mt_linear_bwd_f32(dX.ptr, dW2.ptr, dZ.ptr, None, X1.ptr, W2.ptr, B, 4096, 4096, 2),
    }

    def once(fn, iters=3):
        lib.mt_event_record(e0)
        for _ in range(iters):
            _capi.check(fn())
        lib.mt_event_record(e1)
        lib.mt_event_sync(e1)
        ms = C.c_float()
        lib.mt_event_elapsed_ms(e0, e1, C.byref(ms))
        return ms.value / iters

    for name, fn in cases.items():
        t = {0: [], 1: []}
        for mix in (0, 1):
            lib.mt_gemm_tc_set_mix(mix)
            once(fn, 2)
        for _ in range(6):
            for mix in (0, 1):
                lib.mt_gemm_tc_set_mix(mix)
                t[mix].append(once(fn))
        lib.mt_gemm_tc_set_mix(0)
        print(json.dumps({"gemm": name, "ms_3xtf32_min": round(min(t[0]), 3), "ms_mix_min": round(min(t[1]), 3),
                          "ms_3xtf32_med": round(sorted(t[0])[3], 3), "ms_mix_med": round(sorted(t[1])[3], 3),
                          "speedup_med": round(sorted(t[0])[3] / sorted(t[1])[3], 3)}), flush=True)


if __name__ == "__main__":
    main()
