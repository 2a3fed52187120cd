"""The C-ABI library loads without a GPU and exports exactly what include/kws_b200.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "kws_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kws_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(kws_lib):
    from multilingual_kws_b200 import _lib
    names = header_functions()
    assert len(names) >= 10
    for n in names:
        assert hasattr(kws_lib, n), f"{n} declared in include/kws_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "python binding list and header differ"
    assert kws_lib.kws_abi_version() >= 1


def test_no_cpu_fallback(kws_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = kws_lib.kws_frontend_create(ctypes.byref(h), 16000, 30, 20, 40, 125.0, 7500.0, 10, .025, .06, .05, 1, .95, 80.0,
                                     21, 1, 6)
    assert rc == -2 and not h.value                      # KWS_ERR_CUDA: fails loudly without a device
    assert b"no CPU fallback" in kws_lib.kws_last_error()


def test_bad_config_rejected(kws_lib):
    h = ctypes.c_void_p()
    rc = kws_lib.kws_frontend_create(ctypes.byref(h), 16000, 10, 20, 40, 125.0, 7500.0, 10, .025, .06, .05, 1, .95, 80.0,
                                     21, 1, 6)
    assert rc == -3 and b"fft_size 512" in kws_lib.kws_last_error()
