import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def kws_lib():
    """Builds (if stale) and loads libkws_b200.so."""
    from multilingual_kws_b200 import build as b
    b.build()
    from multilingual_kws_b200 import _lib
    return _lib.lib()


@pytest.fixture(scope="session")
def emu_lib():
    """TEST-ONLY host replay of the CUDA frontend's lane choreography (tests/emu)."""
    import ctypes
    import subprocess
    here = os.path.join(ROOT, "tests", "emu")
    so = os.path.join(here, "libkws_emu.so")
    srcs = [os.path.join(here, "frontend_emulator.cpp"),
            os.path.join(ROOT, "multilingual_kws_b200", "csrc", "frontend_tables.cpp")]
    deps = srcs + [os.path.join(ROOT, "multilingual_kws_b200", "csrc", "frontend_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so] + srcs + ["-lm"])
    return ctypes.CDLL(so)
