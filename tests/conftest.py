import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.dirname(os.path.abspath(__file__)) not in sys.path:
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))  # tests/independent.py


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def native_artifacts():
    """Compile the CUDA library (nvcc cross-compiles without a GPU) and the oracle if they are not
    there yet, so that the suite also runs on a fresh checkout."""
    import __graft_entry__ as entry
    if not os.path.exists(entry.LIB):
        entry.build()
    from oracle import oracle as orc
    orc.build()


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rfsurfhmc_b200._lib import Context
    return Context(0)
