import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding

    binding.lib()
    return binding


@pytest.fixture(scope="session")
def noise64(oracle):
    """64^3 rgba16f noise volume pair from the oracle's xor.wgsl restatement (time = 0)."""
    return oracle.generate_xor(64)


@pytest.fixture(scope="session")
def xor_cam(oracle):
    """examples/xor/main.rs:273-279 camera at 16:9."""
    return oracle.camera_uniform(3.0, -0.5, 1.0, (0.0, 0.0, 0.0), 16 / 9)
