import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


def have_reference():
    return os.path.isfile("/root/reference/src/kernels/sim_kernels.cl")


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def gpu_lib():
    """The CUDA library through its C ABI; fails (never skips, never falls back) when it is missing on a GPU box."""
    from ionsolver_b200 import capi
    lib = capi.load()
    assert capi.device_count() >= 1, "no sm_100 device visible"
    return lib
