import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 (B200) device; run with -m gpu on the GPU box")


def pytest_sessionstart(session):
    """The in-tree libraries are build artefacts (git-ignored): compile them if a fresh checkout has none yet
    (nvcc cross-compiles sm_100a without a GPU; same commands as __graft_entry__.build())."""
    import subprocess
    lib = os.path.join(ROOT, "hybrid-rag-colbertv2_b200", "libhrc.so")
    if not os.path.exists(lib):
        subprocess.run(["make", "-C", os.path.join(ROOT, "hybrid-rag-colbertv2_b200", "csrc"), "-j", "8"], check=True,
                       stdout=subprocess.DEVNULL)
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)


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
