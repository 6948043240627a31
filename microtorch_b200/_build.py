"""Build recipe for libmicrotorch_b200.so (nvcc, sm_100a only, in-tree).

    python -m microtorch_b200._build [--force] [--verbose]

The shared library is written next to the package (microtorch_b200/lib/) so that it
travels with a repo snapshot; it is git-ignored.  nvcc cross-compiles without a GPU.
"""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from typing import List

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
BUILDDIR = os.path.join(PKG, "build")
LIBNAME = "libmicrotorch_b200.so"

SOURCES = [
    "mt_runtime.cu",
    "mt_elementwise.cu",
    "mt_reduce.cu",
    "mt_loss_sgd.cu",
    "mt_gemm_simt.cu",
    "mt_gemm_skinny.cu",
    "mt_gemm_tc.cu",
    "mt_gemm_tc_mid.cu",
    "mt_linear.cu",
    "mt_mlp_small.cu",
    "mt_comm.cu",
    "mt_p2p.cu",
    "mt_rng.cu",
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libmicrotorch_b200.so cannot be built")
    return exe


def lib_path() -> str:
    return os.path.join(LIBDIR, LIBNAME)


def _digest() -> str:
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)):
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(name.encode())
            h.update(f.read())
    with open(os.path.join(os.path.dirname(PKG), "include", "microtorch_b200.h"), "rb") as f:
        h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    stamp = os.path.join(LIBDIR, "build.sha256")
    return (os.path.exists(lib_path()) and os.path.exists(stamp)
            and open(stamp).read().strip() == _digest())


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu for sm_100a and link the shared library.  Returns its path."""
    if not force and is_current():
        return lib_path()
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(BUILDDIR, exist_ok=True)
    cc = nvcc()
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(src: str) -> str:
        obj = os.path.join(BUILDDIR, src.replace(".cu", ".o"))
        cmd = [cc, *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs: List[str] = list(ex.map(compile_one, SOURCES))
    link = [cc, "-shared", "-o", lib_path(), *objs, "-gencode", "arch=compute_100a,code=sm_100a",
            "-cudart", "static", "-ldl", "-lpthread"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(os.path.join(LIBDIR, "build.sha256"), "w") as f:
        f.write(_digest())
    return lib_path()


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
