import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `pytest -m gpu` under gpurun)")


@pytest.fixture(scope="session")
def lib_path():
    """builds the CUDA library if it is missing (nvcc cross-compiles without a GPU)"""
    from eqxvision_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        from eqxvision_b200.csrc import build

        build.build()
    return _lib.LIB_PATH


@pytest.fixture(scope="session")
def device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from eqxvision_b200 import _lib

    torch.cuda.set_device(0)
    _lib.init(0)
    return torch.device("cuda", 0)


@pytest.fixture()
def save_checkpoint(tmp_path):
    import torch

    def _save(sd, name="ckpt.pth"):
        p = tmp_path / name
        torch.save(sd, p)
        return str(p)

    return _save
