import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def product_lib():
    """The built C-ABI library.  Built on demand (nvcc cross-compiles without a GPU)."""
    import cedarx_h264_encoder_b200 as cx
    if not os.path.exists(cx.library_path()):
        cx.build_library()
    return cx.load_library()


@pytest.fixture(scope="session")
def harness():
    import ctypes as C
    so = os.path.join(ROOT, "tests", "libhost_harness.so")
    src = os.path.join(ROOT, "tests", "host_harness.cpp")
    deps = [src] + [os.path.join(ROOT, "cedarx_h264_encoder_b200", "csrc", f) for f in ("entropy.cuh", "h264_core.cuh", "h264_tables.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-g", "-Wall", "-Wno-unknown-pragmas", "-fPIC", "-shared", "-std=c++17", "-o", so, src])
    H = C.CDLL(so)
    H.hh_cavlc_slice.restype = C.c_long
    H.hh_cabac_slice.restype = C.c_long
    H.hh_epb.restype = C.c_long
    H.hh_cavlc_slice.argtypes = [C.c_void_p] * 3 + [C.c_int] * 5 + [C.c_uint64, C.c_int, C.c_void_p, C.c_long, C.c_void_p]
    H.hh_cabac_slice.argtypes = [C.c_void_p] * 3 + [C.c_int] * 6 + [C.c_uint64, C.c_int, C.c_void_p, C.c_long, C.c_int,
                                 C.c_void_p]
    H.hh_epb.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_long]
    return H
