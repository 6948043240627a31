#!/usr/bin/env python
"""Ring-depth sweep of the tcgen05 GEMM (mix mode) on the two big GEMM shapes of config #2.  Run under gpurun."""
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
    lib.mt_gemm_tc_set_rings.argtypes = [C.c_int, C.c_int]
    lib.mt_gemm_tc_set_chunk_kb.argtypes = [C.c_int]
    e0, e1 = C.c_void_p(), C.c_void_p()
    lib.mt_event_create(C.byref(e0))
    lib.mt_event_create(C.byref(e1))
    rng = np.random.default_rng(0)
    B = 65536
    dev = lambda *shape: DeviceArray.from_host(rng.standard_normal(shape).astype(np.float32) * 0.05)
    X1, W2, dZ = dev(B, 4096), dev(4096, 4096), dev(B, 4096)
    Y, dX, dW2 = DeviceArray.empty((B, 4096)), DeviceArray.empty((B, 4096)), DeviceArray.empty((4096, 4096))
    cases = {
        "fwd_L2": lambda: lib.mt_linear_fwd_f32(Y.ptr, X1.ptr, W2.ptr, None, B, 4096, 4096, 1),
        "dW_L2": lambda: lib.mt_linear_bwd_f32(None, dW2.ptr, dZ.ptr, None, X1.ptr, W2.ptr, B, 4096, 4096, 0),
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

    configs = [(0, 3, 8), (0, 2, 8), (0, 4, 8), (4, 4, 8), (5, 2, 8), (0, 3, 4), (0, 3, 16)]
    for name, fn in cases.items():
        times = {c: [] for c in configs}
        for rnd in range(4):
            for c in configs:
                lib.mt_gemm_tc_set_rings(c[0], c[1])
                lib.mt_gemm_tc_set_chunk_kb(c[2])
                if rnd == 0:
                    once(fn, 1)
                times[c].append(once(fn))
        lib.mt_gemm_tc_set_rings(0, 3)
        lib.mt_gemm_tc_set_chunk_kb(8)
        print(json.dumps({"gemm": name, "ms_min_by_raw_lo_chunk": {f"{c[0]}/{c[1]}/{c[2]}": round(min(v), 3) for c, v in times.items()}}),
              flush=True)


if __name__ == "__main__":
    main()
