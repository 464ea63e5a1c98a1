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
    from oracle import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def nds():
    """The product package (ndarray-stats_b200/), imported under the name ndarray_stats_b200."""
    import _ndstats_import
    return _ndstats_import.load()


@pytest.fixture(scope="session")
def gpu_ctx(nds):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    ctx = nds.Context(device=0)
    yield ctx
    ctx.close()
