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


def _has_b200() -> bool:
    try:
        import torch
        return torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] >= 10
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a box without a B200 skips the gpu-marked tests instead of failing them;
    `-m gpu` on such a box still runs (and fails) them, so a GPU tier can never pass vacuously."""
    if _has_b200() or "gpu" in (config.getoption("-m") or ""):
        return
    skip = pytest.mark.skip(reason="needs a B200 (sm_100): run with -m gpu under gpurun")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


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
