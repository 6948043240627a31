#!/usr/bin/env python
"""Where should Linear GEMMs switch between the SIMT fp32 tile, the mid-size tcgen05 kernel and the persistent tcgen05 kernel?  Times mt_linear_fwd_f32 on both
engines over a grid of (M, N, K) and prints one JSON line per shape (run under gpurun).  The dispatch threshold in
mt_gemm_tc.cu (gemm_tc_try) comes from this table."""
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
    e0, e1 = C.c_void_p(), C.c_void_p()
    lib.mt_event_create(C.byref(e0))
    lib.mt_event_create(C.byref(e1))
    rng = np.random.default_rng(0)

    def timed(fn, iters=30):
        for _ in range(5):
            fn()
        lib.mt_sync()
        lib.mt_event_record(e0)
        for _ in range(iters):
            fn()
        lib.mt_event_record(e1)
        lib.mt_event_sync(e1)
        ms = C.c_float()
        lib.mt_event_elapsed_ms(e0, e1, C.byref(ms))
        return ms.value / iters * 1e3          # us

    shapes = []
    for m in (256, 1024, 4096, 16384):
        for n, k in ((64, 64), (128, 128), (256, 256), (512, 512), (1024, 1024), (256, 1024), (1024, 256)):
            shapes.append((m, n, k))
    for (M, N, K) in shapes:
        X = DeviceArray.from_host(rng.standard_normal((M, K)).astype(np.float32))
        Wt = DeviceArray.from_host((rng.standard_normal((N, K)) * 0.05).astype(np.float32))
        b = DeviceArray.from_host(rng.standard_normal(N).astype(np.float32))
        Y = DeviceArray.empty((M, N))
        rec = {"M": M, "N": N, "K": K, "mmacs": M * N * K / 1e6}
        for name, eng in (("simt", 0), ("tcgen05", 1), ("tcgen05_mid", 2)):
            lib.mt_gemm_set_engine(eng)
            try:
                fn = lambda: _capi.check(lib.mt_linear_fwd_f32(Y.ptr, X.ptr, Wt.ptr, b.ptr, M, N, K, 1))
                rec[name + "_us"] = round(timed(fn), 2)
            except Exception as ex:        # shape refused by the forced engine
                rec[name + "_us"] = None
            finally:
                lib.mt_gemm_set_engine(-1)
        if rec["simt_us"] and rec["tcgen05_us"]:
            rec["tc_speedup"] = round(rec["simt_us"] / rec["tcgen05_us"], 2)
        if rec["simt_us"] and rec.get("tcgen05_mid_us"):
            rec["mid_speedup"] = round(min(rec["simt_us"], rec["tcgen05_us"] or 1e30) / rec["tcgen05_mid_us"], 2)
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
