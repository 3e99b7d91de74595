import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Make sure libdotsocp.so and the oracle's C restatement exist (built in-tree; they travel with gpurun)."""
    import subprocess
    lib = os.path.join(ROOT, "dotsocp_b200", "libdotsocp.so")
    olib = os.path.join(ROOT, "oracle", "liboracle_kernels.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "dotsocp_b200", "csrc"), "-j4"])
    if not os.path.exists(olib):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle_kernels.so"])
    return True


@pytest.fixture(scope="session")
def gpu(built):
    from dotsocp_b200 import _lib
    n = _lib.lib().dotsocp_device_count()
    if n <= 0:
        pytest.fail("GPU test selected but no CUDA device is usable: " + _lib.lib().dotsocp_last_error().decode())
    return n
