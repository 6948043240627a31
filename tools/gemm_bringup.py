#!/usr/bin/env python
"""Hardware bring-up of the tcgen05 3xTF32 GEMM through the C ABI (run under gpurun).

Each case runs in its own subprocess with a timeout: a device trap (bounded mbarrier wait)
or a wrong descriptor must not take the remaining cases down.  Prints one JSON line per case:
error of the tensor-core result against float64 NumPy, against the SIMT engine, and TFLOP/s.
"""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

CASES = [
    # name, M, N, K, kind, write_hi, timing
    ("fwd_small", 256, 256, 64, "fwd", 1, 0),
    ("fwd_small_trunc", 256, 256, 64, "fwd", 0, 0),
    ("fwd_bias_tanh", 384, 512, 160, "fwd_act", 1, 0),
    ("fwd_ragged", 200, 132, 100, "fwd", 1, 0),
    ("fwd_n128", 512, 128, 256, "fwd", 1, 0),
    ("dx_small", 256, 256, 128, "dx", 1, 0),
    ("dx_prologue", 256, 256, 128, "dx_pro", 1, 0),
    ("dw_small", 256, 256, 256, "dw", 1, 0),
    ("dw_prologue_acc", 512, 256, 384, "dw_pro_acc", 1, 0),
    ("fwd_4096", 8192, 4096, 4096, "fwd_act", 1, 1),
    ("fwd_4096_trunc", 8192, 4096, 4096, "fwd_act", 0, 1),
    ("dx_4096", 8192, 4096, 4096, "dx_pro", 1, 1),
    ("dw_4096", 8192, 4096, 4096, "dw_pro", 1, 1),
    ("fwd_c2_l1", 65536, 4096, 1024, "fwd_act", 1, 1),
    ("fwd_c2_l2", 65536, 4096, 4096, "fwd_act", 1, 1),
    ("dw_c2_l2", 65536, 4096, 4096, "dw_pro", 1, 1),
]


def run_case(name, M, N, K, kind, write_hi, timing):
    import ctypes as C
    import numpy as np
    from microtorch_b200 import _capi
    from capi_util import Dev, timed
    lib = _capi.lib()
    _capi.require_device(0)
    lib.mt_gemm_tc_set_write_hi.argtypes = [C.c_int]
    lib.mt_gemm_tc_set_write_hi(write_hi)
    rng = np.random.default_rng(1)
    # Linear convention: X [M,K] rows, W [N,K], Y/dY [M,N]
    X = rng.standard_normal((M, K), dtype=np.float32)
    W = (rng.standard_normal((N, K), dtype=np.float32) * 0.05).astype(np.float32)
    b = rng.standard_normal(N, dtype=np.float32)
    dY = rng.standard_normal((M, N), dtype=np.float32)
    Yact = np.tanh(rng.standard_normal((M, N), dtype=np.float32)).astype(np.float32)
    dX_, dW_ = None, None
    big = M * N * K > 2 ** 31

    def ref64():
        if big:   # sample rows/cols only
            return None
        if kind.startswith("fwd"):
            z = X.astype(np.float64) @ W.astype(np.float64).T
            if kind == "fwd_act":
                z = np.tanh(z + b)
            return z
        dz = dY.astype(np.float64)
        if "pro" in kind:
            dz = dz * (1 - Yact.astype(np.float64) ** 2)
        if kind.startswith("dx"):
            return dz @ W.astype(np.float64)
        r = dz.T @ X.astype(np.float64)
        if kind.endswith("acc"):
            r = r + 1.0
        return r

    dX_d, dW_d = None, None
    Xd, Wd, bd, dYd, Yd = Dev(X), Dev(W), Dev(b), Dev(dY), Dev(Yact)

    def call(engine):
        lib.mt_gemm_set_engine(engine)
        if kind.startswith("fwd"):
            out = Dev(shape=(M, N))
            rc = lib.mt_linear_fwd_f32(out.ptr, Xd.ptr, Wd.ptr, bd.ptr if kind == "fwd_act" else None, M, N, K,
                                       1 if kind == "fwd_act" else 0)
            _capi.check(rc, "mt_linear_fwd_f32")
            return out
        yact = Yd.ptr if "pro" in kind else None
        dW = Dev(np.ones((N, K), np.float32)) if kind.endswith("acc") else Dev(shape=(N, K))
        dX = Dev(shape=(M, K)) if kind.startswith("dx") else None
        rc = lib.mt_linear_bwd_f32(dX.ptr if dX else None, dW.ptr, dYd.ptr, yact, Xd.ptr, Wd.ptr, M, N, K,
                                   1 if kind.endswith("acc") else 0)
        _capi.check(rc, "mt_linear_bwd_f32")
        return dX if kind.startswith("dx") else dW

    res = {"case": name, "M": M, "N": N, "K": K, "kind": kind, "write_hi": write_hi}
    out_tc = call(1)
    _capi.check(lib.mt_sync(), "sync")
    eng = C.c_int(-2)
    lib.mt_gemm_last_engine(C.byref(eng))
    res["engine"] = eng.value
    tc = out_tc.get()
    out_simt = call(0)
    _capi.check(lib.mt_sync(), "sync")
    simt = out_simt.get()
    scale = float(np.abs(simt).max()) + 1e-30
    res["tc_vs_simt_maxabs_over_max"] = float(np.abs(tc.astype(np.float64) - simt).max() / scale)
    r = ref64()
    if r is not None:
        s = float(np.abs(r).max()) + 1e-30
        res["tc_vs_f64"] = float(np.abs(tc - r).max() / s)
        res["simt_vs_f64"] = float(np.abs(simt - r).max() / s)
    res["finite"] = bool(np.isfinite(tc).all())
    if timing:
        del out_simt
        lib.mt_gemm_set_engine(1)
        flops = 2.0 * M * N * K
        outs = []

        def f():
            o = call(1)
            outs.append(o)
            if len(outs) > 2:
                outs.pop(0)
        ms = timed(f, iters=5, warmup=2, flush=False)
        res["ms"] = ms
        res["tflops_logical"] = flops / ms / 1e9
    print("RESULT " + json.dumps(res), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--case":
        c = [x for x in CASES if x[0] == sys.argv[2]][0]
        run_case(*c)
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
                print(json.dumps({"case": c[0], "rc": r.returncode, "stdout": r.stdout[-600:], "stderr": r.stderr[-1200:]}),
                      flush=True)
        except subprocess.TimeoutExpired:
            print(json.dumps({"case": c[0], "error": "timeout"}), flush=True)
