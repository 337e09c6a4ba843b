"""Shared fixtures. GPU tests are marked ``@pytest.mark.gpu`` and call the CUDA path through
the C-ABI; everything else runs on CPU (oracle vs golden vectors, host logic, ABI exports)."""
import base64
import json
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
GOLDEN = Path(__file__).resolve().parent / "golden"
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def load_golden(name):
    with open(GOLDEN / f"{name}.json") as f:
        g = json.load(f)
    for key in ("idx", "dat"):
        if key in g and isinstance(g[key], str):
            g[key] = base64.b64decode(g[key])
    return g


@pytest.fixture(scope="session")
def golden():
    return load_golden
