import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle_mod():
    """The CPU oracle (test infrastructure).  Built on demand with oracle/Makefile."""
    from oracle import oracle as orc

    orc.build()
    return orc


@pytest.fixture(scope="session")
def handler():
    """One CUDA handler for the whole GPU session (fails loudly if libsdfmesh.so or the device is missing)."""
    import bsdmg_b200

    h = bsdmg_b200.CudaHandler(0)
    yield h
    h.close()
