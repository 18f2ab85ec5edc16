import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def sb_lib():
    """Built C-ABI library (compiles it if needed; nvcc cross-compiles without a GPU)."""
    from sionna_b200.csrc import build as b
    b.build()
    from sionna_b200 import _lib
    return _lib.lib()


@pytest.fixture(scope="session")
def cuda_device(sb_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


@pytest.fixture(autouse=True)
def _seed_per_test(request):
    """Every test starts from its own fixed seed (CRC of the test id), so random inputs do not depend on which tests ran
    before it; tests that need a particular stream still set ``config.seed`` themselves."""
    import zlib
    from sionna_b200.phy import config
    config.seed = zlib.crc32(request.node.nodeid.encode()) & 0x7FFFFFFF
    yield
