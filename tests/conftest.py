import os
import sys

import pytest

# Tests that put several "virtual ranks" on one GPU (tests/test_gpu_group.py) keep one kernel per rank spinning on its
# peers' flags; every stream needs its own hardware queue so a queued successor never blocks a peer's launch.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    config.addinivalue_line("markers", "heavy: large (BASELINE-size) case")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def gpu_ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import minarrow_b200 as mnr
    return mnr.Context(0)
