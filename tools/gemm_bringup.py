#!/usr/bin/env python
"""Hardware bring-up / characterisation of the tcgen05 3xTF32 GEMM through the C ABI (run under gpurun).

Each case runs in its own subprocess with a timeout: a device trap (bounded mbarrier wait)
or a wrong descriptor must not take the remaining cases down.  Prints one JSON line per case:
error of the tensor-core result against float64 NumPy (on a slice for the big cases), against
the SIMT engine, and logical TFLOP/s (CUDA events, L2 not flushed: operands exceed L2 anyway).
"""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

# name, M, N, K (Linear convention: X[M,K], W[N,K], Y[M,N]), kind, write_hi, chunk_kb, timing
CASES = [
    # name, M, N, K, kind, write_hi, chunk_kb, timing, pair
    ("fwd_small", 256, 256, 64, "fwd", 0, 8, 0, 1),
    ("fwd_small_1cta", 256, 256, 64, "fwd", 0, 8, 0, 0),
    ("fwd_bias_tanh", 384, 512, 160, "fwd_act", 0, 8, 0, 1),
    ("fwd_ragged", 300, 132, 100, "fwd", 0, 8, 0, 1),
    ("fwd_n128", 512, 128, 256, "fwd", 0, 8, 0, 1),
    ("fwd_chunks", 512, 256, 1024, "fwd", 0, 4, 0, 1),
    ("dx_small", 256, 256, 128, "dx", 0, 8, 0, 1),
    ("dx_prologue", 512, 256, 128, "dx_pro", 0, 8, 0, 1),
    ("dw_small", 256, 256, 256, "dw", 0, 8, 0, 1),
    ("dw_prologue_acc", 512, 256, 384, "dw_pro_acc", 0, 8, 0, 1),
    ("dw_ragged", 300, 260, 420, "dw_pro", 0, 8, 0, 1),
    ("fwd_4096_pair", 8192, 4096, 4096, "fwd_act", 0, 8, 1, 1),
    ("fwd_4096_1cta", 8192, 4096, 4096, "fwd_act", 0, 8, 1, 0),
    ("fwd_c2_l1_pair", 65536, 4096, 1024, "fwd_act", 0, 8, 1, 1),
    ("fwd_c2_l2_pair", 65536, 4096, 4096, "fwd_act", 0, 8, 1, 1),
    ("fwd_c2_l2_1cta", 65536, 4096, 4096, "fwd_act", 0, 8, 1, 0),
    ("dx_c2_l2_pair", 65536, 4096, 4096, "dx_pro", 0, 8, 1, 1),
    ("dw_c2_l2_pair", 65536, 4096, 4096, "dw_pro", 0, 8, 1, 1),
    ("dw_c2_l2_1cta", 65536, 4096, 4096, "dw_pro", 0, 8, 1, 0),
    ("dw_c2_l1_pair", 65536, 4096, 1024, "dw_pro", 0, 8, 1, 1),
]


def run_case(name, M, N, K, kind, write_hi, chunk_kb, timing, pair):
    import ctypes as C
    import numpy as np
    from microtorch_b200 import _capi
    from capi_util import Dev, timed
    lib = _capi.lib()
    _capi.require_device(0)
    lib.mt_gemm_tc_set_write_hi.argtypes = [C.c_int]
    lib.mt_gemm_tc_set_chunk_kb.argtypes = [C.c_int]
    lib.mt_gemm_tc_set_write_hi(write_hi)
    lib.mt_gemm_tc_set_chunk_kb(chunk_kb)
    lib.mt_gemm_tc_set_pair.argtypes = [C.c_int]
    lib.mt_gemm_tc_set_pair(pair)
    rng = np.random.default_rng(1)
    X = rng.standard_normal((M, K), dtype=np.float32)
    W = (rng.standard_normal((N, K), dtype=np.float32) * 0.05).astype(np.float32)
    b = rng.standard_normal(N, dtype=np.float32)
    dY = rng.standard_normal((M, N), dtype=np.float32)
    Yact = np.tanh(rng.standard_normal((M, N), dtype=np.float32)).astype(np.float32)
    S = 64   # slice width for the float64 check of big problems
    big = M * N * K > 2 ** 31

    def ref64():
        """(reference, slicer) - full for small problems, a slice for big ones."""
        if kind.startswith("fwd"):
            rows = slice(0, S) if big else slice(None)
            z = X[rows].astype(np.float64) @ W.astype(np.float64).T
            if kind == "fwd_act":
                z = np.tanh(z + b)
            return z, (lambda o: o[rows])
        dz = dY.astype(np.float64)
        if "pro" in kind:
            dz = dz * (1 - Yact.astype(np.float64) ** 2)
        if kind.startswith("dx"):
            rows = slice(0, S) if big else slice(None)
            return dz[rows] @ W.astype(np.float64), (lambda o: o[rows])
        rows = slice(0, S) if big else slice(None)
        r = dz[:, rows].T @ X.astype(np.float64)
        if kind.endswith("acc"):
            r = r + 1.0
        return r, (lambda o: o[rows])

    Xd, Wd, bd, dYd, Yd = Dev(X), Dev(W), Dev(b), Dev(dY), Dev(Yact)

    def call(engine, want_dx=True):
        lib.mt_gemm_set_engine(engine)
        if kind.startswith("fwd"):
            out = Dev(shape=(M, N))
            rc = lib.mt_linear_fwd_f32(out.ptr, Xd.ptr, Wd.ptr, bd.ptr if kind == "fwd_act" else None, M, N, K,
                                       1 if kind == "fwd_act" else 0)
            _capi.check(rc, "mt_linear_fwd_f32")
            return out
        yact = Yd.ptr if "pro" in kind else None
        dW = Dev(np.ones((N, K), np.float32)) if kind.endswith("acc") else Dev(shape=(N, K))
        dX = Dev(shape=(M, K)) if (kind.startswith("dx") and want_dx) else None
        rc = lib.mt_linear_bwd_f32(dX.ptr if dX else None, dW.ptr, dYd.ptr, yact, Xd.ptr, Wd.ptr, M, N, K,
                                   1 if kind.endswith("acc") else 0)
        _capi.check(rc, "mt_linear_bwd_f32")
        return dX if (kind.startswith("dx") and want_dx) else dW

    res = {"case": name, "M": M, "N": N, "K": K, "kind": kind, "write_hi": write_hi, "chunk_kb": chunk_kb, "pair": pair}
    out_tc = call(1)
    _capi.check(lib.mt_sync(), "sync")
    eng = C.c_int(-2)
    lib.mt_gemm_last_engine(C.byref(eng))
    res["engine"] = eng.value
    tc = out_tc.get()
    res["finite"] = bool(np.isfinite(tc).all())
    r, sl = ref64()
    s = float(np.abs(r).max()) + 1e-30
    res["tc_vs_f64"] = float(np.abs(sl(tc) - r).max() / s)
    if not big or not timing:
        out_simt = call(0)
        _capi.check(lib.mt_sync(), "sync")
        simt = out_simt.get()
        res["simt_vs_f64"] = float(np.abs(sl(simt) - r).max() / s)
        res["tc_vs_simt"] = float(np.abs(tc.astype(np.float64) - simt).max() / (np.abs(simt).max() + 1e-30))
        del out_simt
    if timing:
        lib.mt_gemm_set_engine(1)
        flops = 2.0 * M * N * K
        outs = []

        def f(want_dx=True):
            o = call(1, want_dx)
            outs.append(o)
            if len(outs) > 2:
                outs.pop(0)
        ms = timed(f, iters=5, warmup=2, flush=False)
        if kind.startswith("dx"):                       # bwd call = dW GEMM + dX GEMM: subtract the dW-only time
            ms_dw = timed(lambda: f(False), iters=5, warmup=2, flush=False)
            res["ms_dw_plus_dx"], res["ms_dw_only"] = ms, ms_dw
            ms = ms - ms_dw
        res["ms"] = ms
        res["tflops_logical"] = flops / ms / 1e9
    print("RESULT " + json.dumps(res), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--case":
        c = [x for x in CASES if x[0] == sys.argv[2]][0]
        try:
            run_case(*c)
        except Exception as e:           # after a device trap: report which barrier wait timed out
            import ctypes as C
            from microtorch_b200 import _capi
            lib = _capi.lib()
            lib.mt_gemm_tc_debug_word.restype = C.POINTER(C.c_int)
            w = lib.mt_gemm_tc_debug_word()
            words = [hex(w[i]) for i in range(4)] if w else None
            print("RESULT " + json.dumps({"case": c[0], "error": str(e)[-300:], "debug_word": words}), flush=True)
            sys.exit(1)
        sys.exit(0)
    only = sys.argv[1:] or None
    for c in CASES:
        if only and c[0] not in only:
            continue
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", c[0]], capture_output=True,
                               text=True, timeout=240)
            lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            if lines:
                print(lines[-1][7:], flush=True)
            else:
                print(json.dumps({"case": c[0], "rc": r.returncode, "stdout": r.stdout[-600:], "stderr": r.stderr[-1500:]}),
                      flush=True)
        except subprocess.TimeoutExpired:
            print(json.dumps({"case": c[0], "error": "timeout"}), flush=True)
