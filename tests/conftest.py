import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests skip (instead of failing inside the library) on a box without a CUDA device."""
    try:
        import torch

        have_gpu = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


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
