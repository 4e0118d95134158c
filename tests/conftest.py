import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 via gpurun)")


@pytest.fixture(scope="session")
def pkg():
    import _pkg
    return _pkg.load()


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/libqr_oracle.so).  Test infrastructure only."""
    import oracle as O
    O.build()
    O.lib()
    return O


@pytest.fixture(scope="session")
def emul():
    """Host emulation of the device code (tests/emul).  Test infrastructure only."""
    import emul_binding
    return emul_binding.load()


@pytest.fixture(scope="session")
def gpu(pkg):
    """Initialised product library on cuda:0; fails loudly if the CUDA extension is unusable."""
    import torch
    assert torch.cuda.is_available(), "GPU test selected but no CUDA device"
    from quadruped_robot_b200 import capi
    capi.init(0)
    return capi
