import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.build()
    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def hostlib():
    """g++ build of the product's host tables + the lane-array emulator of the SCL kernel (tests/scl_emulator.cc)."""
    import ctypes
    from modem_b200 import build as B
    so = B.HOSTTEST
    src = [os.path.join(ROOT, "tests", "scl_emulator.cc"), os.path.join(B.CSRC, "host_tables.cc"), os.path.join(B.CSRC, "host_tables.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        import subprocess
        os.makedirs(B.OBJ, exist_ok=True)
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", src[0], src[1], "-o", so], check=True)
    return ctypes.CDLL(so)


@pytest.fixture(scope="session", autouse=True)
def _built_product(request):
    """GPU sessions need libofdmrx.so and the `decode` host driver: (re)build them when sources are newer (nvcc, in-tree)."""
    if "gpu" in (request.config.getoption("-m") or "") and "not gpu" not in (request.config.getoption("-m") or ""):
        from modem_b200 import build as B
        B.build(helpers=False)


@pytest.fixture(scope="session")
def rx():
    import modem_b200 as M
    r = M.Receiver(max_frames=2048, keep_taps=True)
    yield r
    r.close()
