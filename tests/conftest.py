import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "host_emul")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def emul():
    """Block-emulator device: the real kernel bodies + the real engine, on the CPU (tests only)."""
    from emul_device import EmulDevice
    return EmulDevice()


@pytest.fixture(scope="session")
def cuda_dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from audio_suite_b200.engine import CudaDevice
    return CudaDevice(0)


def rel_err(a, b):
    import numpy as np
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))) /
                 max(1e-300, float(np.max(np.abs(b)))))
