"""pytest configuration: registers the `gpu` marker and puts the product package
(`local-hyperdb_b200/hyperdb_b200`) and the test-only `oracle/` on sys.path."""
import ctypes as C
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "local-hyperdb_b200"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def emul(tmp_path_factory):
    """The product's csrc/canonical.cuh + csrc/certificate.cuh compiled for the HOST (tests/emul/canonical_host.cpp)."""
    inc = "/usr/local/cuda/include"
    if shutil.which("g++") is None or not os.path.exists(os.path.join(inc, "cuda_fp16.h")):
        pytest.skip("needs g++ and the CUDA headers")
    so = str(tmp_path_factory.mktemp("emul") / "libcanon_host.so")
    cmd = ["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-I" + inc, "-Wno-attributes", "-o", so,
           os.path.join(ROOT, "tests", "emul", "canonical_host.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    lib = C.CDLL(so)
    lib.emul_pairwise_sum.restype = C.c_double
    lib.emul_pairwise_sum.argtypes = [C.c_int, C.c_void_p, C.c_int]
    lib.emul_pairwise_sum_plan.restype = C.c_double
    lib.emul_pairwise_sum_plan.argtypes = [C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    lib.emul_norm.restype = C.c_double
    lib.emul_norm.argtypes = [C.c_int, C.c_void_p, C.c_int64]
    lib.emul_mean_std.restype = None
    lib.emul_mean_std.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.emul_similarity.restype = C.c_double
    lib.emul_similarity.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_double,
                                    C.c_void_p, C.c_void_p, C.c_int]
    lib.emul_outsider_bound.restype = C.c_double
    lib.emul_outsider_bound.argtypes = ([C.c_double, C.c_int, C.c_int, C.c_int, C.c_int64] + [C.c_double] * 5 +
                                        [C.c_int, C.c_double, C.c_int, C.c_double, C.c_double])
    lib.emul_outsider_bound_tc.restype = C.c_double
    lib.emul_outsider_bound_tc.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, C.c_int64] + [C.c_double] * 8
    return lib


