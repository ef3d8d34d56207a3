"""Shared test plumbing.  `-m "not gpu"` runs on a CPU-only box; `-m gpu` needs a B200."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "portable-multigrid_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _has_gpu():
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
def oracle():
    import pyoracle

    pyoracle.lib()  # builds oracle/_build/liborc.so on first use
    return pyoracle


@pytest.fixture(scope="session")
def emu():
    """Host emulator of the CUDA tile programs (tests/emu), built on demand with g++."""
    import ctypes

    path = os.path.join(ROOT, "tests", "emu", "_build", "libemu.so")
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "emu")], check=True, stdout=subprocess.DEVNULL)
    return ctypes.CDLL(path)


@pytest.fixture(scope="session")
def pmg():
    import pmg_b200

    return pmg_b200


@pytest.fixture(scope="session")
def ctx(pmg):
    c = pmg.Context(0)
    yield c
    c.sync()
