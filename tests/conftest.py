import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")
    # make sure the checker and the product library exist (both build on a CPU-only box)
    import __graft_entry__ as ge
    ge.build(quiet=True, only_missing=True)


@pytest.fixture(scope="session")
def oracle():
    from oracle import loa
    return loa.port()


@pytest.fixture(scope="session")
def ref_oracle():
    from oracle import loa
    r = loa.ref()
    if r is None:
        pytest.skip("oracle/_ref not built (needs /root/reference; container only)")
    return r


@pytest.fixture(scope="session", params=[2, 1, 0], ids=["sweep", "lut", "bitboard"])
def gpu(request):
    """The engine, once per device formulation of the rules (both must be bit-exact against the oracle)."""
    import gameai_b200 as g
    eng = g.GpuRollout(0)      # raises loudly without a device / library: no fallback
    eng.set_rules_impl(request.param)
    eng.impl = request.param
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def hostlib():
    """TEST HELPER library: the product's __host__ __device__ rule code (Pan rules, LOA line LUT, the static sweep)
    compiled for the CPU, so that its logic is checked against the oracle without a GPU."""
    import ctypes
    import glob
    import subprocess
    so = os.path.join(ROOT, "tests", "helpers", "libhost_rules.so")
    srcs = [os.path.join(ROOT, "tests", "helpers", "host_rules.cu")] + glob.glob(os.path.join(ROOT, "game-ai_b200", "csrc", "*.cuh")) \
        + [os.path.join(ROOT, "include", "gameai", "philox.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-shared",
                               "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "game-ai_b200", "csrc"), "-o", so, srcs[0], "-cudart", "static"])
    H = ctypes.CDLL(so)
    H.host_pan_movegen.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    H.host_pan_apply.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p]
    H.host_pan_terminal.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return H
