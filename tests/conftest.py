import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pk():
    import petsc_openacc_b200 as pk
    return pk


@pytest.fixture(scope="session")
def cuda(pk):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    pk.init(0)
    return torch
