"""The C-ABI library loads and exports every symbol include/rfa_b200.h declares (no compute:
this runs on the CPU-only build box), and the ctypes table stays in step with the header."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rfa_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rfa_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_functions():
    names = declared_functions()
    assert "rfa_spectrum_process" in names and "rfa_ctx_create" in names and len(names) >= 25


def test_library_exports_every_declared_symbol():
    from rfanalyzer_b200 import _lib
    lib = _lib.load()
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in rfa_b200.h but not exported by librfa_b200.so"


def test_ctypes_table_matches_header():
    from rfanalyzer_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_functions()


def test_version_and_error_reporting_without_gpu():
    from rfanalyzer_b200 import _lib
    lib = _lib.load()
    assert lib.rfa_version() == 100
    # host-only entry points work anywhere
    import numpy as np
    w = np.empty(16, np.float32)
    assert lib.rfa_make_window(_lib.WIN_BLACKMAN_REF, 16, w.ctypes.data) == 0
    assert abs(w[0]) < 1e-6 and abs(w[15]) < 1e-6
    assert lib.rfa_make_window(99, 16, w.ctypes.data) == _lib.ERR_INVALID
    assert b"unknown window" in lib.rfa_last_error()


def test_no_cpu_fallback():
    """Without a CUDA device context creation must fail loudly (never compute on the host)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import rfanalyzer_b200 as rfa
    with pytest.raises(rfa.RfaError) as e:
        rfa.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_touch_the_oracle():
    """Nothing under rfanalyzer_b200/ may import, include or link oracle/."""
    pkg = os.path.join(ROOT, "rfanalyzer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower(), f"{os.path.join(dirpath, f)} mentions the oracle"
