import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as O

    O.build()
    O.load()
    return O


@pytest.fixture(scope="session")
def nk():
    """The product package with its CUDA library built (nvcc cross-compiles without a GPU)."""
    import newtonkrylov_jl_b200 as nk_

    if not os.path.exists(nk_.LIB_PATH):
        nk_._build.build()
    return nk_


@pytest.fixture(scope="session")
def ctx(nk):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return nk.get_context(0)
