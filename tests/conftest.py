import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE = "/root/reference"  # only present in the build container, never on the GPU box


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure).  Built on demand with gcc."""
    from oracle import oracle as o
    o.build()
    o.lib()
    return o


@pytest.fixture(scope="session")
def exact_math_host():
    """Host build of turbo_metrics_b200/csrc/exact_math.cuh (test-only)."""
    import ctypes
    d = os.path.join(ROOT, "tests", "native")
    subprocess.check_call(["make", "-C", d, "libexact_math_host.so"], stdout=subprocess.DEVNULL)
    return ctypes.CDLL(os.path.join(d, "libexact_math_host.so"))


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
