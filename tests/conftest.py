"""Shared fixtures.  CPU tests (-m "not gpu") cover the oracle pair, the host
logic and the ABI; GPU tests (-m gpu) are the parity tests proper and call the
sweep through the C-ABI."""
import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools", "amgen")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # build in-tree artefacts once (no-op when up to date)
    build = importlib.import_module("active-monitor_b200.build")
    build.build_all()


@pytest.fixture(scope="session")
def am():
    """the product package (directory name is not an identifier)"""
    return importlib.import_module("active-monitor_b200")


@pytest.fixture(scope="session")
def lib(am):
    return am.load()


@pytest.fixture(scope="session")
def orc():
    import oracle_c
    oracle_c.load()
    return oracle_c


@pytest.fixture(scope="session")
def opy():
    import oracle_py
    return oracle_py


@pytest.fixture(scope="session")
def gen():
    import amgen
    amgen.load()
    return amgen


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
