import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# the library's on-disk cubin cache defaults to ~/.cache/swec; the test suite keeps its files inside the repo
os.environ.setdefault("SWEC_CACHE_DIR", os.path.join(ROOT, ".pytest_cache", "swec_cubins"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def kat():
    with open(os.path.join(ROOT, "tests", "golden", "kat.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (checker only).  Builds liboracle.so on first use."""
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def swec():
    """The product library; built in-tree if missing (nvcc cross-compiles without a GPU)."""
    from seaweedfs_b200 import _native, build
    if not os.path.exists(_native.library_path()):
        build.build()
    import seaweedfs_b200
    seaweedfs_b200.lib()
    return seaweedfs_b200


@pytest.fixture(scope="session")
def cuda(swec):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.init()
    return torch
