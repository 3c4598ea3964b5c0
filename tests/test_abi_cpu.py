"""The C-ABI shared library: loads without a GPU, exports every symbol include/paml_b200.h declares, and
refuses loudly to compute when no GPU is present (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from paml_b200 import _lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_and_symbol_list_agree():
    hdr = open(os.path.join(REPO, "include", "paml_b200.h")).read()
    declared = set(re.findall(r"\b(pb200_[A-Za-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS)


def test_library_exports_every_symbol():
    L = _lib.lib()
    for s in _lib.SYMBOLS:
        assert hasattr(L, s), s
    assert b"sm_100a" in L.pb200_version()


def test_no_cpu_fallback(have_gpu):
    if have_gpu:
        pytest.skip("GPU present")
    ctx = C.c_void_p()
    rc = _lib.lib().pb200_create(C.byref(ctx), C.c_int(0))
    assert rc != 0 and not ctx
    assert b"no CPU fallback" in _lib.lib().pb200_last_error() or b"CUDA" in _lib.lib().pb200_last_error()


def test_product_does_not_touch_the_oracle():
    """Nothing under paml_b200/ or include/ may import, include or link oracle/."""
    bad = []
    for root in ("paml_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(REPO, root)):
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".c", ".inc")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"(from|import)\s+oracle|#include\s+[\"<].*oracle|liboracle", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
