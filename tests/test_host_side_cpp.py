"""Compiles and runs tests/cpp/test_host_side.cpp: the host-only entry points of the C-ABI
(bulk ingest, hand-off queue) called from plain C++ threads, as a compiled controller would.
No GPU needed."""
import os
import subprocess

from conftest import ROOT

SRC = os.path.join(ROOT, "tests", "cpp", "test_host_side.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_host_side.bin")
LIBDIR = os.path.join(ROOT, "active-monitor_b200", "lib")


def test_host_side_entry_points_from_cpp_threads():
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-pthread", "-I", os.path.join(ROOT, "include"),
                    SRC, "-L", LIBDIR, "-lamsweep", f"-Wl,-rpath,{LIBDIR}", "-o", EXE], check=True)
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout + out.stderr
