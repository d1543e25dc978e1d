import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def asia():
    d = np.load(os.path.join(GOLDEN, "asia.npz"))
    return d["codes"], d["card"]


@pytest.fixture(scope="session")
def sachs():
    d = np.load(os.path.join(GOLDEN, "sachs.npz"))
    return d["codes"], d["card"]


@pytest.fixture(scope="session")
def known_answer():
    with open(os.path.join(GOLDEN, "asia_known_answer.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
