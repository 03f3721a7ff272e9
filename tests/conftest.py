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
