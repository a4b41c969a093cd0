"""Compiles and runs tests/cpp/test_reconciler.cpp — the C++ host-side mirror of the
reference's HealthCheckReconciler surface (include/amsweep_reconciler.hpp) — against
libamsweep.  The compile step also runs on CPU; running needs a GPU."""
import os
import subprocess

import pytest

from conftest import ROOT

SRC = os.path.join(ROOT, "tests", "cpp", "test_reconciler.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_reconciler.bin")
LIBDIR = os.path.join(ROOT, "active-monitor_b200", "lib")


def _build():
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC,
                    "-L", LIBDIR, "-lamsweep", f"-Wl,-rpath,{LIBDIR}", "-o", EXE], check=True)


def test_reconciler_mirror_compiles():
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_reconciler_mirror_behaves_like_the_reference_tests():
    _build()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout + out.stderr
