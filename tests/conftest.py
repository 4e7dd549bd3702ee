import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib

    oracle_lib.load()
    return oracle_lib


@pytest.fixture(scope="session")
def product_lib():
    from rs_pbrt_b200 import _abi, _build

    if not _abi.LIB_PATH.exists():
        _build.build()
    return _abi.load()
