import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
sys.dont_write_bytecode = True

GOLDEN = os.path.join(REPO, "tests", "golden")
REFERENCE = os.environ.get("MICROTORCH_REFERENCE", "/root/reference")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def have_reference() -> bool:
    return os.path.isdir(os.path.join(REFERENCE, "src", "tensor"))


needs_reference = pytest.mark.skipif(not have_reference(),
                                     reason="/root/reference only exists in the build container")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
