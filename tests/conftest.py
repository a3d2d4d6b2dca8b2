import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 (B200) device; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def cuda_dev():
    """GPU tests call through the C ABI; a missing library or device is a failure, never a skip."""
    import torch
    import hybrid_rag_colbertv2_b200 as hrc
    assert torch.cuda.is_available(), "gpu-marked test started without a CUDA device"
    hrc._lib.load()
    return torch.device("cuda:0")
