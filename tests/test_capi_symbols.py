"""CPU-only: the C-ABI library builds, loads without a GPU, and exports every symbol that
include/microtorch_b200.h declares (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

from microtorch_b200 import _build, _capi

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(REPO, "include", "microtorch_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mt_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    _build.build()
    return _capi.load_library()


def test_header_and_binding_agree():
    declared = set(header_symbols())
    bound = set(_capi.PROTOTYPES) | set(_capi.STRING_FUNCS)
    assert declared == bound, (declared - bound, bound - declared)


def test_library_exports_every_declared_symbol(lib):
    for name in header_symbols():
        assert hasattr(lib, name), name


def test_no_gpu_means_loud_failure_not_fallback(lib):
    if _capi.device_count() > 0:
        pytest.skip("a GPU is visible")
    assert lib.mt_is_initialized() == 0
    out = C.c_void_p()
    assert lib.mt_malloc(C.byref(out), 16) == 3          # MT_ERR_NOT_INIT
    assert b"mt_init" in lib.mt_last_error()
    assert lib.mt_init(0) != 0
    with pytest.raises(_capi.MicrotorchB200Error):
        _capi.require_device(0)


def test_header_is_plain_c_and_the_desc_struct_matches_ctypes(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 (no C++, no torch types), and the one struct it passes
    by pointer must have the layout the ctypes mirror assumes."""
    import subprocess
    hdr = os.path.join(REPO, "include", "microtorch_b200.h")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr], check=True)
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "microtorch_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(mt_mlp_desc), offsetof(mt_mlp_desc, dims),'
                   ' offsetof(mt_mlp_desc, act), offsetof(mt_mlp_desc, weight), offsetof(mt_mlp_desc, bias),'
                   ' offsetof(mt_mlp_desc, grad_weight)); return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(REPO, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    D = _capi.MlpDesc
    assert got == [C.sizeof(D), D.dims.offset, D.act.offset, D.weight.offset, D.bias.offset, D.grad_weight.offset]
