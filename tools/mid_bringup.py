#!/usr/bin/env python
"""Bring-up and crossover measurement of the mid-size tcgen05 kernel (mt_gemm_tc_mid.cu).

    python tools/mid_bringup.py [check] [time] [--out gpurun_out/mid.jsonl]

`check`: forward (bias + tanh), dW, dX (with the producer's tanh' epilogue) and bare strided matmuls through the forced
mid kernel (engine 2) against float64, for every precision level and every number of K slices.
`time` : the same Linear forward on the SIMT tile (engine 0), the persistent CTA-pair kernel (1) and the mid kernel (2)
over a grid of shapes, back-to-back launches timed with CUDA events.  One JSON line per case (run under gpurun)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np  # noqa: E402


def main():
    from microtorch_b200 import _capi
    from capi_util import Dev
    _capi.require_device(0)
    lib = _capi.lib()
    lib.mt_gemm_tc_debug_word.restype = C.POINTER(C.c_int)
    out_path = None
    if "--out" in sys.argv:
        out_path = sys.argv[sys.argv.index("--out") + 1]
    out = open(out_path, "a") if out_path else None

    def emit(rec):
        line = json.dumps(rec)
        print(line, flush=True)
        if out:
            out.write(line + "\n")
            out.flush()

    def err(got, ref):
        return float(np.max(np.abs(got.astype(np.float64) - ref)) / max(np.max(np.abs(ref)), 1e-30))

    def dbg_word():
        p = lib.mt_gemm_tc_debug_word()
        return [hex(p[i]) for i in range(4)] if p else None

    def mid_count():
        n = C.c_uint64()
        lib.mt_gemm_tc_mid_launches(C.byref(n))
        return n.value

    if "resident" in sys.argv or "sweep" in sys.argv:
        res = (C.c_longlong * 16)()
        _capi.check(lib.mt_gemm_tc_mid_resident(res))
        emit({"case": "resident_ctas_by_cluster_size", "tile_128x128": list(res[:8]), "tile_pair": list(res[8:])})

    if "check" in sys.argv:
        shapes = [(256, 128, 64), (1024, 512, 768), (1000, 260, 100), (129, 68, 4100), (640, 1100, 516), (2049, 260, 264),
                  (1024, 1024, 1024), (4096, 256, 1024)]
        for (M, N, K) in shapes:
            rng = np.random.default_rng(M + N + K)
            X = np.tanh(rng.standard_normal((M, K))).astype(np.float32)
            W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
            b = rng.standard_normal(N).astype(np.float32)
            dY = rng.standard_normal((M, N)).astype(np.float32)
            Xd, Wd, bd, dYd = Dev(X), Dev(W), Dev(b), Dev(dY)
            ref_y = np.tanh(X.astype(np.float64) @ W.astype(np.float64).T + b)
            ref_dw = dY.astype(np.float64).T @ X.astype(np.float64)
            ref_dx = (dY.astype(np.float64) @ W.astype(np.float64)) * (1 - X.astype(np.float64) ** 2)
            for level, ncta, splits in [(lv, nc, sp) for lv in (0, 1, 2) for nc in (0, 1, 2) for sp in (0, 1, 2, 4, 8)
                                        if not (nc == 0 and sp) and not (nc == 2 and sp == 8)] + [(0, 1, 3), (1, 1, 5), (0, 1, 7), (1, 2, 3)]:
                if True:
                    rec = {"case": "check", "M": M, "N": N, "K": K, "precision": level, "tile": ncta, "splits": splits}
                    try:
                        lib.mt_gemm_set_precision(level)
                        lib.mt_gemm_tc_mid_set(1, 0, 0, splits)
                        lib.mt_gemm_tc_mid_set_tile(ncta, 0)
                        lib.mt_gemm_set_engine(2)
                        n0 = mid_count()
                        Y, dX = Dev(shape=(M, N)), Dev(shape=(M, K))
                        dW = Dev(np.full((N, K), 0.5, np.float32))
                        _capi.check(lib.mt_linear_fwd_f32(Y.ptr, Xd.ptr, Wd.ptr, bd.ptr, M, N, K, 1))
                        _capi.check(lib.mt_linear_bwd_f32(dX.ptr, dW.ptr, dYd.ptr, None, Xd.ptr, Wd.ptr, M, N, K, 1 | 2))
                        _capi.check(lib.mt_sync())
                        rec["launches"] = mid_count() - n0
                        rec["fwd"] = err(Y.get(), ref_y)
                        rec["dW"] = err(dW.get(), ref_dw + 0.5)
                        rec["dX"] = err(dX.get(), ref_dx)
                        rec["ok"] = bool(rec["launches"] == 3 and max(rec["fwd"], rec["dW"], rec["dX"]) < 2e-5)
                    except Exception as ex:  # noqa: BLE001
                        rec["error"] = str(ex)[:300]
                        rec["dbg"] = dbg_word()
                        emit(rec)
                        return 1
                    finally:
                        lib.mt_gemm_set_engine(-1)
                        lib.mt_gemm_set_precision(1)
                        lib.mt_gemm_tc_mid_set(1, 0, 0, 0)
                        lib.mt_gemm_tc_mid_set_tile(0, 0)
                    emit(rec)
        # bare matmul with transposed views: A MN-major x B K-major, the one major pair Linear never produces
        M, N, K = 512, 384, 640
        rng = np.random.default_rng(5)
        At = rng.standard_normal((K, M)).astype(np.float32)      # A(m,k) = At[k,m]
        Bt = rng.standard_normal((N, K)).astype(np.float32)      # B(k,n) = Bt[n,k]
        Ad, Bd, Cd = Dev(At), Dev(Bt), Dev(shape=(M, N))
        rec = {"case": "check_matmul_mn_k", "M": M, "N": N, "K": K}
        try:
            lib.mt_gemm_set_engine(2)
            _capi.check(lib.mt_matmul_f32(Cd.ptr, Ad.ptr, Bd.ptr, 1, M, N, K, 0, 1, M, 0, 1, K))
            _capi.check(lib.mt_sync())
            rec["err"] = err(Cd.get(), At.astype(np.float64).T @ Bt.astype(np.float64).T)
            rec["ok"] = bool(rec["err"] < 2e-5)
        except Exception as ex:  # noqa: BLE001
            rec["error"] = str(ex)[:300]
            rec["dbg"] = dbg_word()
        finally:
            lib.mt_gemm_set_engine(-1)
        emit(rec)

    if "time" in sys.argv:
        e0, e1 = C.c_void_p(), C.c_void_p()
        lib.mt_event_create(C.byref(e0))
        lib.mt_event_create(C.byref(e1))

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

        rng = np.random.default_rng(0)
        shapes = [(256, 256, 256), (256, 512, 512), (1024, 128, 256), (1024, 256, 256), (512, 512, 512), (1024, 512, 512),
                  (1024, 1024, 256), (1024, 256, 1024), (4096, 128, 128), (4096, 256, 256), (2048, 512, 512),
                  (1024, 1024, 1024), (4096, 512, 512), (16384, 256, 256), (2048, 1024, 1024), (2048, 2048, 512),
                  (8192, 512, 512), (4096, 1024, 1024), (16384, 512, 512), (2048, 2048, 1024), (2048, 2048, 2048),
                  (8192, 1024, 1024), (4096, 4096, 1024), (16384, 1024, 1024)]
        levels = (0, 1) if "--levels" in sys.argv else (1,)
        for level in levels:
            lib.mt_gemm_set_precision(level)
            for (M, N, K) in shapes:
                X = Dev(rng.standard_normal((M, K)).astype(np.float32))
                Wt = Dev((rng.standard_normal((N, K)) * 0.05).astype(np.float32))
                b = Dev(rng.standard_normal(N).astype(np.float32))
                Y = Dev(shape=(M, N))
                rec = {"case": "time_fwd", "precision": level, "M": M, "N": N, "K": K, "log2_macs": round(float(np.log2(M * N * K)), 2)}
                for name, eng in (("simt", 0), ("pair", 1), ("mid", 2), ("auto", -1)):
                    lib.mt_gemm_set_engine(eng)
                    try:
                        fn = lambda: _capi.check(lib.mt_linear_fwd_f32(Y.ptr, X.ptr, Wt.ptr, b.ptr, M, N, K, 1))  # noqa: E731
                        rec[name + "_us"] = round(timed(fn), 2)
                    except Exception as ex:  # noqa: BLE001
                        rec[name + "_us"] = None
                        rec[name + "_error"] = str(ex)[:200]
                        rec["dbg"] = dbg_word()
                    finally:
                        lib.mt_gemm_set_engine(-1)
                if rec.get("mid_us"):
                    rec["mid_tflops"] = round(2.0 * M * N * K / rec["mid_us"] / 1e6, 1)
                    best_other = min(v for v in (rec.get("simt_us"), rec.get("pair_us")) if v)
                    rec["mid_speedup_vs_best_other"] = round(best_other / rec["mid_us"], 2)
                if rec.get("auto_us"):
                    rec["auto_tflops"] = round(2.0 * M * N * K / rec["auto_us"] / 1e6, 1)
                emit(rec)
                for d in (X, Wt, b, Y):
                    d.free()
        lib.mt_gemm_set_precision(1)

    if "sweep" in sys.argv:
        # forced K-slice counts on a few shapes: device time per launch (events around back-to-back launches) and the
        # host's own cost per call (wall clock of queueing 200 launches without synchronising)
        import time
        e0, e1 = C.c_void_p(), C.c_void_p()
        lib.mt_event_create(C.byref(e0))
        lib.mt_event_create(C.byref(e1))
        rng = np.random.default_rng(1)
        iters = 8 if "--few" in sys.argv else 40
        for (M, N, K) in [(1024, 1024, 32), (1024, 1024, 1024), (1024, 256, 1024), (1024, 512, 512), (4096, 512, 512), (2048, 2048, 512), (2048, 2048, 2048)]:
            X = Dev(rng.standard_normal((M, K)).astype(np.float32))
            Wt = Dev((rng.standard_normal((N, K)) * 0.05).astype(np.float32))
            b = Dev(rng.standard_normal(N).astype(np.float32))
            Y = Dev(shape=(M, N))
            for level in (0, 1):
                lib.mt_gemm_set_precision(level)
                for ncta, splits in ((0, 0), (1, 1), (1, 2), (1, 3), (1, 4), (1, 6), (1, 8), (2, 1), (2, 2), (2, 3), (2, 4)):
                    lib.mt_gemm_tc_mid_set(1, 0, 0, splits)
                    lib.mt_gemm_tc_mid_set_tile(ncta, 16 if "--nopre" in sys.argv else 0)
                    lib.mt_gemm_set_engine(2)
                    fn = lambda: _capi.check(lib.mt_linear_fwd_f32(Y.ptr, X.ptr, Wt.ptr, b.ptr, M, N, K, 1))  # noqa: E731
                    for _ in range(3):
                        fn()
                    lib.mt_sync()
                    lib.mt_event_record(e0)
                    t0 = time.perf_counter()
                    for _ in range(iters):
                        fn()
                    t1 = time.perf_counter()
                    lib.mt_event_record(e1)
                    lib.mt_event_sync(e1)
                    ms = C.c_float()
                    lib.mt_event_elapsed_ms(e0, e1, C.byref(ms))
                    lib.mt_gemm_set_engine(-1)
                    try:
                        _capi.check(lib.mt_sync(), "mt_sync")
                    except Exception as ex:  # noqa: BLE001
                        emit({"case": "sweep_error", "M": M, "N": N, "K": K, "precision": level, "tile": ncta, "splits": splits,
                              "error": str(ex)[:300], "dbg": dbg_word()})
                        return 1
                    us = max(ms.value / iters * 1e3, 1e-3)
                    emit({"case": "sweep", "M": M, "N": N, "K": K, "precision": level, "tile": ncta, "splits": splits,
                          "us_per_launch": round(us, 2), "host_us_per_call": round((t1 - t0) / iters * 1e6, 2),
                          "tflops": round(2.0 * M * N * K / us / 1e6, 1)})
            lib.mt_gemm_tc_mid_set(1, 0, 0, 0)
            lib.mt_gemm_tc_mid_set_tile(0, 0)
            lib.mt_gemm_set_precision(1)

    if "stamps" in sys.argv:
        # phase clocks of CTA (0,0,0) (mt_gemm_tc_mid_set_stamps): where the fixed cost of a launch goes
        rng = np.random.default_rng(2)
        st = Dev(np.zeros(16, np.int64))
        lib.mt_gemm_tc_mid_set_stamps(C.c_void_p(st.ptr))
        names = ["entry", "prologue_done", "first_kb_converted", "last_kb_converted", "last_chunk_drained", "staged",
                 "cluster_barrier_1", "fold_epilogue_done", "cluster_barrier_2"]
        for (M, N, K, ncta, splits, level) in [(1024, 1024, 32, 1, 1, 0), (1024, 1024, 1024, 1, 2, 0), (1024, 1024, 1024, 1, 2, 1),
                                               (1024, 1024, 1024, 2, 2, 0), (1024, 1024, 1024, 2, 4, 0), (1024, 512, 512, 1, 4, 0),
                                               (2048, 2048, 2048, 2, 1, 0)]:
            X = Dev(rng.standard_normal((M, K)).astype(np.float32))
            Wt = Dev((rng.standard_normal((N, K)) * 0.05).astype(np.float32))
            b = Dev(rng.standard_normal(N).astype(np.float32))
            Y = Dev(shape=(M, N))
            lib.mt_gemm_set_precision(level)
            lib.mt_gemm_tc_mid_set(1, 0, 0, splits)
            lib.mt_gemm_tc_mid_set_tile(ncta, 0)
            lib.mt_gemm_set_engine(2)
            for _ in range(3):
                _capi.check(lib.mt_linear_fwd_f32(Y.ptr, X.ptr, Wt.ptr, b.ptr, M, N, K, 1))
            _capi.check(lib.mt_sync())
            lib.mt_gemm_set_engine(-1)
            t = st.get()
            rec = {"case": "stamps", "M": M, "N": N, "K": K, "tile": ncta, "splits": splits, "precision": level,
                   "cycles_since_entry": {names[i]: int(t[i] - t[0]) for i in range(1, 9)}}
            emit(rec)
        lib.mt_gemm_tc_mid_set_stamps(None)
        lib.mt_gemm_tc_mid_set(1, 0, 0, 0)
        lib.mt_gemm_tc_mid_set_tile(0, 0)
        lib.mt_gemm_set_precision(1)

    if "ncu" in sys.argv:
        # a handful of single launches for `ncu --set full -k regex:gemm_mid` (forward, both precisions; dW; dX)
        rng = np.random.default_rng(3)
        for (M, N, K) in [(1024, 1024, 1024), (2048, 1024, 1024), (1024, 512, 512)]:
            X, W, b, dY = (rng.standard_normal((M, K)).astype(np.float32), (rng.standard_normal((N, K)) * 0.05).astype(np.float32),
                           rng.standard_normal(N).astype(np.float32), rng.standard_normal((M, N)).astype(np.float32))
            Xd, Wd, bd, dYd = Dev(X), Dev(W), Dev(b), Dev(dY)
            Y, dX, dW = Dev(shape=(M, N)), Dev(shape=(M, K)), Dev(shape=(N, K))
            for level in (0, 1):
                lib.mt_gemm_set_precision(level)
                for _ in range(2):
                    _capi.check(lib.mt_linear_fwd_f32(Y.ptr, Xd.ptr, Wd.ptr, bd.ptr, M, N, K, 1))
            lib.mt_gemm_set_precision(1)
            for _ in range(2):
                _capi.check(lib.mt_linear_bwd_f32(dX.ptr, dW.ptr, dYd.ptr, None, Xd.ptr, Wd.ptr, M, N, K, 2))
            _capi.check(lib.mt_sync())
    return 0


if __name__ == "__main__":
    sys.exit(main())
