import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def oracle():
    from tests.oracle_lib import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def engine():
    import modelx_b200
    eng = modelx_b200.Engine(devices=[0])
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def mock_lib():
    """Path of the CPU test double of libmodelxdigest.so (host sources + a synchronous CUDA stand-in + oracle
    hashing, see tests/mock/include/cuda_runtime.h).  Test infrastructure: never loaded by the product."""
    from tests import mock_build
    # MXD_MOCK_SANITIZE=address|thread|address,undefined selects an instrumented build (run pytest with the matching
    # runtime preloaded, see tools/asan_host_suite.sh)
    return mock_build.build(os.environ.get("MXD_MOCK_SANITIZE", ""))


@pytest.fixture(params=[pytest.param("mock"), pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request, mock_lib):
    """Host-logic tests run twice: against the test double here (`-m "not gpu"`) and against the CUDA build on the
    B200 (`-m gpu`).  Yields the lib_path argument for modelx_b200.Engine (None = the product library)."""
    return mock_lib if request.param == "mock" else None


@pytest.fixture
def any_engine(backend):
    import modelx_b200
    eng = modelx_b200.Engine(devices=[0], lib_path=backend)
    yield eng
    eng.close()
