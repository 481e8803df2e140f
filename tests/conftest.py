import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
# parity tests compare against the fp32 CPU oracle at summation-order tolerances: strict-fp32 tiles unless a test
# asks for the tensor-core path explicitly (precision="tf32")
os.environ.setdefault("TD3_PRECISION", "fp32")

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def has_reference():
    return os.path.isfile(os.path.join(REFERENCE, "TD3_featured.py"))
