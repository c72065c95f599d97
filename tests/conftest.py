import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

REF_INPUTS = "/root/reference/EN234_FEA/input_files"
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    import __graft_entry__ as g
    g.build()


def input_file(name):
    """Reference input files are not available on the GPU box: tests use the copies of the two hex8 UEL inputs
    committed under tests/golden/ (fixtures, not source code)."""
    p = os.path.join(GOLDEN, name)
    if os.path.exists(p):
        return p
    return os.path.join(REF_INPUTS, name)
