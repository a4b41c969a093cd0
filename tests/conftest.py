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
    # Time-zone ids are handed out in order of first appearance, per implementation: introduce the
    # generator's zones first and in its order everywhere, so that flags columns (and the golden
    # digests over them) do not depend on which test happened to parse which "CRON_TZ=" first.
    import ctypes as C
    import amgen
    import oracle_c
    import oracle_py
    am = importlib.import_module("active-monitor_b200")
    for z in amgen.ZONES:
        i = C.c_int32()
        assert am.tz_lookup(z) == oracle_py.tz_lookup(z)
        assert oracle_c.load().orc_tz_lookup(z.encode(), len(z), C.byref(i)) == 0 and i.value == oracle_py.tz_lookup(z)


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
    import amgen
    import oracle_py
    for z in amgen.ZONES:  # zone ids in the generator's order, as in the other implementations
        oracle_py.tz_lookup(z)
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
