import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(params=["warp", "thread"])
def norm_path(request, monkeypatch):
    """The scan kernels (normalisers, BytesToChars, CharsToBytes, UTF8Validate) have a warp-per-string and a
    thread-per-string form chosen by the average string length; tests pin each in turn (B200TOK_NORM_PATH)."""
    monkeypatch.setenv("B200TOK_NORM_PATH", request.param)
    return request.param
