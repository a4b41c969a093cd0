"""The WHOLE product library on the CPU emulator, in-process: csrc/*.cu + kernels compiled against
tests/emu/cuda_emu.h (execution model) and tests/emu/cuda_rt_emu.h (runtime calls) into
tests/emu/libamsweep_emu.so, and `EmuSweep` = `am.Sweep` bound to that library instead of the CUDA
build — same C-ABI, same host runtime, same kernel source.  Test infrastructure only.

(Round 1 kept a second, kernel-only harness here that restated the launch sequence of sweep.cu;
it is gone: the real launch sequence runs on the emulator.)"""
import ctypes as C
import importlib
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "active-monitor_b200", "csrc")
LIB = os.path.join(HERE, "libamsweep_emu.so")
_lib = None


def sources():
    b = importlib.import_module("active-monitor_b200.build")
    return [os.path.join(CSRC, f) for f in b.SOURCES], [os.path.join(CSRC, f) for f in b.HEADERS]


def build():
    srcs, hdrs = sources()
    deps = srcs + hdrs + [os.path.join(HERE, "cuda_emu.h"), os.path.join(HERE, "cuda_rt_emu.h"),
                          os.path.join(ROOT, "include", "amsweep.h")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wno-unknown-pragmas", "-pthread", "-fPIC", "-shared",
                        # same-named symbols exist in libamsweep.so (loaded RTLD_GLOBAL by the tests): bind locally
                        "-Wl,-Bsymbolic",
                        "-DAMSWEEP_EMULATE", "-include", os.path.join(HERE, "cuda_emu.h"),
                        "-include", os.path.join(HERE, "cuda_rt_emu.h"), "-x", "c++"] + srcs + ["-o", LIB],
                       check=True)
    return LIB


def load():
    global _lib
    if _lib is None:
        abi = importlib.import_module("active-monitor_b200._lib")
        lib = C.CDLL(build(), mode=os.RTLD_LOCAL)
        for name, (res, args) in abi.SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        # this second copy of the library has its own time-zone registry: introduce the generator's
        # zones in the generator's order, so that records classified by the CUDA build of the
        # library (the tests do that) carry ids this copy understands
        sys.path.insert(0, os.path.join(ROOT, "tools", "amgen"))
        import amgen
        for z in amgen.ZONES:
            i = C.c_int32()
            assert lib.am_tz_lookup(z.encode(), len(z), C.byref(i)) == 0
        _lib = lib
    return _lib


def _sweep_class():
    am = importlib.import_module("active-monitor_b200")

    class EmuSweep(am.Sweep):
        def __init__(self, capacity: int, shard_base: int = 0):
            super().__init__(capacity, device=0, shard_base=shard_base, lib=load())

        def tick(self, unix_sec, mode=0, cap=None, buffers=None):
            """like am.Sweep.tick, but a short buffer returns the partial list (the kernel-level
            harness of round 1 did) instead of raising"""
            try:
                return super().tick(unix_sec, mode, cap, buffers)
            except am.AmError as e:
                if e.code != am.AM_E_NOSPACE:
                    raise
                return e.partial

    return EmuSweep


def EmuSweep(capacity: int, shard_base: int = 0):
    return _sweep_class()(capacity, shard_base)
