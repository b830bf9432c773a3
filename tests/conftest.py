import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `-m gpu` on the GPU box)")


@pytest.fixture(scope="session")
def dab():
    import darray_b200

    return darray_b200


@pytest.fixture()
def rt8(dab):
    """8 workers on the one visible GPU: multi-chunk layouts (grids, fibres, halos) on a single-GPU box."""
    rt = dab.init(workers_per_rank=8, use_dist=False)
    yield rt
    dab.d_closeall()


@pytest.fixture()
def rt1(dab):
    rt = dab.init(workers_per_rank=1, use_dist=False)
    yield rt
    dab.d_closeall()


@pytest.fixture()
def rt2(dab):
    rt = dab.init(workers_per_rank=2, use_dist=False)
    yield rt
    dab.d_closeall()
