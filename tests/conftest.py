import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def kats():
    with open(os.path.join(GOLDEN, "reference_kats.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def gold64():
    with open(os.path.join(GOLDEN, "goldilocks_vectors.json")) as f:
        return json.load(f)


def pt(v):
    """golden JSON point → 4-byte wire format"""
    return bytes([0xFF] * 4) if v == "inf" else bytes(v)
