import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def _cpu_only_run(config):
    return (config.getoption("markexpr", "") or "").strip() == "not gpu"


@pytest.hookimpl(tryfirst=True)
def pytest_cmdline_main(config):
    """The CPU suite (-m "not gpu") spends most of its time in the kernel-emulation tests, which are independent of each other: spread
    it over a few pytest-xdist workers when the plugin is there and the caller did not choose (-n ...; RS_PBRT_TEST_WORKERS=0 turns it
    off).  GPU runs (-m gpu) are never parallelised: one device, one context."""
    import os

    if not _cpu_only_run(config) or os.environ.get("PYTEST_XDIST_WORKER") or not config.pluginmanager.hasplugin("xdist"):
        return
    if getattr(config.option, "numprocesses", None) is not None:
        return
    n = os.environ.get("RS_PBRT_TEST_WORKERS")
    n = int(n) if n is not None else max(1, min(6, (os.cpu_count() or 2) // 2))  # (the emulation is one host thread per process since the fiber engine)
    if n > 1:
        config.option.numprocesses = n
        config.option.dist = "worksteal"  # the few long tests (the -m gpu files through the emulation) sit in one file: idle workers take them over


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    # shared libraries are built once, here in the controlling process, before any worker could race to build them
    import os
    import shutil

    if not os.environ.get("PYTEST_XDIST_WORKER") and _cpu_only_run(config):
        import oracle_lib
        from rs_pbrt_b200 import _build

        oracle_lib.build()
        if not os.environ.get("RS_PBRT_B200_LIB"):  # (a developer override names a library that is built elsewhere)
            _build.build()
        if shutil.which("g++") is not None:
            sys.path.insert(0, str(ROOT / "tests" / "emu"))
            import build_emu

            build_emu.build()


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
