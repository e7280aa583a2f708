import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def port():
    import oracle_lib
    return oracle_lib.port()


@pytest.fixture(scope="session")
def ref():
    import oracle_lib
    r = oracle_lib.ref()
    if r is None:
        pytest.skip("oracle/_ref/libnutsref.so not built (reference sources absent)")
    return r


@pytest.fixture(scope="session")
def gpu_ctx():
    """A Context on cuda:0 through the product library.  No fallback: if the
    library or the device is missing the test errors out."""
    from nuts333_b200 import build, api
    build.build()
    ctx = api.Context(0)
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def sim_lib():
    """libnutsb200_sim.so: the product's .cu sources compiled against the SIMT
    emulator (tests/cpusim).  Test infrastructure only."""
    import ctypes
    from cpusim.build_sim import build_sim
    from nuts333_b200 import api
    return api.bind(ctypes.CDLL(str(build_sim())))
