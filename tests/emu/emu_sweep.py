"""ctypes face of tests/emu/libemu_sweep.so: the sweep kernels of csrc/sweep_kernels.cuh run on
the CPU emulator (cuda_emu.h).  Mirrors the part of `am.Sweep` the parity tests use, so the
emulator tests read like tests/test_sweep_gpu.py.  Test infrastructure only.

The staging of upsert / remove / post_result into 8-byte events is the one piece of host logic
of csrc/sweep.cu restated here (am_sweep_upsert / _remove / _post_result build the same array)."""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "libemu_sweep.so")
_lib = None


def build():
    src = os.path.join(HERE, "emu_sweep.cpp")
    deps = [src, os.path.join(HERE, "cuda_emu.h"), os.path.join(ROOT, "include", "amsweep.h"),
            os.path.join(ROOT, "active-monitor_b200", "csrc", "sweep_kernels.cuh"),
            os.path.join(ROOT, "active-monitor_b200", "csrc", "civil.h")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wno-unknown-pragmas", "-pthread", "-fPIC", "-shared",
                        # same-named kernel symbols exist in libamsweep.so (loaded RTLD_GLOBAL): bind locally
                        "-fvisibility=hidden", "-Wl,-Bsymbolic",
                        "-I", os.path.join(ROOT, "include"), src, "-o", LIB], check=True)
    return LIB


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        lib.emu_sweep_create.restype = C.c_void_p
        lib.emu_sweep_create.argtypes = [C.c_uint64, C.c_uint64]
        lib.emu_sweep_destroy.argtypes = [C.c_void_p]
        lib.emu_sweep_load.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
        lib.emu_sweep_read.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
        lib.emu_sweep_tick.argtypes = [C.c_void_p, C.c_int64, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p,
                                       C.c_uint64, C.c_void_p, C.c_void_p]
        lib.emu_sweep_apply_ops.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_int]
        lib.emu_sweep_repeat_after_sec.argtypes = [C.c_void_p, C.c_int64, C.c_uint32, C.c_uint32, C.c_void_p]
        lib.emu_sweep_gather.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        for f in ("emu_op_upsert", "emu_op_remove", "emu_op_result"):
            getattr(lib, f).restype = C.c_uint32
        _lib = lib
    return _lib


class EmuSweep:
    def __init__(self, capacity: int, shard_base: int = 0):
        self.am = importlib.import_module("active-monitor_b200")
        self.lib = load()
        self.capacity, self.shard_base, self.seed = capacity, shard_base, 0
        self.h = C.c_void_p(self.lib.emu_sweep_create(capacity, shard_base))
        self.size = 0
        self._ops, self._recs, self._n_state, self._n_result = [], [], 0, 0
        self.OP_UPSERT, self.OP_REMOVE, self.OP_RESULT = (self.lib.emu_op_upsert(), self.lib.emu_op_remove(),
                                                          self.lib.emu_op_result())

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.lib.emu_sweep_destroy(self.h)
        self.h = None

    def _ptrs(self, cols):
        arr = (C.c_void_p * 16)()
        for k, (name, dt) in enumerate(self.am.abi.COLUMNS):
            a = cols.get(name)
            if a is not None:
                assert a.dtype == np.dtype(dt) and a.flags["C_CONTIGUOUS"], name
                arr[k] = a.ctypes.data
        return arr

    def set_seed(self, seed):
        self.seed = seed

    def load_range(self, first, cols):
        n = len(cols["flags"])
        assert self.lib.emu_sweep_load(self.h, first, n, self._ptrs(cols)) == 0
        self.size = max(self.size, first + n)

    def read_range(self, first, n):
        cols = {name: np.zeros(n, dtype=dt) for name, dt in self.am.abi.COLUMNS}
        self._drain()
        assert self.lib.emu_sweep_read(self.h, first, n, self._ptrs(cols)) == 0
        return cols

    def read(self, idx):
        self._drain()
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        out = np.zeros(len(idx), dtype=self.am.RECORD_DTYPE)
        self.lib.emu_sweep_gather(self.h, idx.ctypes.data, out.ctypes.data, len(idx))
        return self.am.records_to_columns(out)

    # -- staged events, in call order (am_sweep_upsert / _remove / _post_result)
    def upsert(self, idx, recs):
        for i, r in zip(np.asarray(idx, dtype=np.uint64).tolist(), recs):
            r = np.array(r, dtype=self.am.RECORD_DTYPE).reshape(())
            r["flags"] &= ~np.uint32(self.am.F_TOMBSTONE)
            self._ops.append((i, self.OP_UPSERT | len(self._recs)))
            self._recs.append(r)
            self._n_state += 1

    def remove(self, idx):
        for i in np.asarray(idx, dtype=np.uint64).tolist():
            self._ops.append((i, self.OP_REMOVE))
            self._n_state += 1

    def post_result(self, idx, phase, remedy_phase=None):
        A = self.am
        pb = {0: 0, 1: A.F_PENDING_OK, 2: A.F_PENDING_FAIL}
        rb = {0: 0, 1: A.F_REMEDY_PENDING | A.F_REMEDY_OUTCOME_OK, 2: A.F_REMEDY_PENDING}
        idx = np.asarray(idx, dtype=np.uint64).tolist()
        rp = [0] * len(idx) if remedy_phase is None else list(np.asarray(remedy_phase).tolist())
        for i, p, r in zip(idx, np.asarray(phase).tolist(), rp):
            self._ops.append((i, self.OP_RESULT | pb[int(p)] | rb[int(r)]))
            self._n_result += 1

    def _drain(self):
        if not self._ops:
            return
        ops = np.array(self._ops, dtype=np.uint32).reshape(-1, 2)
        recs = (np.array(self._recs, dtype=self.am.RECORD_DTYPE) if self._recs
                else np.zeros(1, dtype=self.am.RECORD_DTYPE))
        rc = self.lib.emu_sweep_apply_ops(self.h, ops.ctypes.data, len(ops), recs.ctypes.data,
                                          self._n_state, self._n_result)
        assert rc == 0, rc
        ups = [i for i, a in self._ops if (a >> 30) == 0]
        if ups:
            self.size = max(self.size, max(ups) + 1)
        self._ops, self._recs, self._n_state, self._n_result = [], [], 0, 0

    def tick(self, T, mode=0, cap=None):
        """(global idx u64[n], action u32[n], stats dict) like am.Sweep.tick"""
        self._drain()
        cap = self.capacity if cap is None else cap
        idx = np.zeros(max(cap, 1), dtype=np.uint32)
        act = np.zeros(max(cap, 1), dtype=np.uint8)
        cnt = C.c_uint32(0)
        st = np.zeros(1, dtype=self.am.abi.STATS_DTYPE)
        assert self.lib.emu_sweep_tick(self.h, T, mode, self.seed, idx.ctypes.data, act.ctypes.data, cap,
                                       C.byref(cnt), st.ctypes.data) == 0
        n = min(cnt.value, cap)
        stats = {f: int(st[f][0]) for f in self.am.abi.STAT_FIELDS}
        return idx[:n].astype(np.uint64) + np.uint64(self.shard_base), act[:n].astype(np.uint32), stats

    def repeat_after_sec(self, T, first, n):
        self._drain()
        out = np.zeros(n, dtype=np.int64)
        self.lib.emu_sweep_repeat_after_sec(self.h, T, first, n, out.ctypes.data)
        return out
