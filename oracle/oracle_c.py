"""ctypes binding of the C oracle (oracle/_build/libamsweep_oracle.so).

TEST INFRASTRUCTURE ONLY — see oracle/amsweep_oracle.h.  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libamsweep_oracle.so")

u64, i64, u32, i32 = C.c_uint64, C.c_int64, C.c_uint32, C.c_int32


class OrcCron(C.Structure):
    _fields_ = [("minute", u64), ("hour", u64), ("dom", u64), ("month", u64), ("dow", u64),
                ("delay_sec", i64), ("kind", i32), ("tz_id", i32)]


COLUMNS = [("minute", np.uint64), ("hour", np.uint64), ("dom", np.uint64), ("month", np.uint64),
           ("dow", np.uint64), ("ras", np.int32), ("flags", np.uint32), ("finished_at", np.int64),
           ("runs_limit", np.int32), ("reset_interval", np.int32), ("success", np.int32),
           ("failed", np.int32), ("remedy_success", np.int32), ("remedy_failed", np.int32),
           ("remedy_total", np.int32), ("remedy_finished_at", np.int64)]


class OrcCols(C.Structure):
    _fields_ = [(n, C.c_void_p) for n, _ in COLUMNS]


STAT_FIELDS = ["n_records", "n_emitted", "n_submit_hc", "n_run_remedy", "n_stopped",
               "n_parse_error", "n_remedy_skip", "n_reset_on_pass", "n_reset_on_interval",
               "n_anomaly", "n_result_ok", "n_result_fail", "n_remedy_ok", "n_remedy_fail",
               "idx_xor", "idx_sum"]


class OrcStats(C.Structure):
    _fields_ = [(n, u64) for n in STAT_FIELDS]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n in STAT_FIELDS}


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} missing: run `make -C oracle`")
        lib = C.CDLL(LIB_PATH)
        lib.orc_tz_lookup.restype = C.c_int
        lib.orc_tz_lookup.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(i32)]
        lib.orc_tz_offset.restype = C.c_int
        lib.orc_tz_offset.argtypes = [i32, i64, C.POINTER(i32)]
        lib.orc_cron_parse.restype = C.c_int
        lib.orc_cron_parse.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(OrcCron), C.c_char_p, C.c_size_t]
        lib.orc_cron_matches.restype = C.c_int
        lib.orc_cron_matches.argtypes = [C.POINTER(OrcCron), i64]
        lib.orc_cron_next.restype = i64
        lib.orc_cron_next.argtypes = [C.POINTER(OrcCron), i64]
        lib.orc_cron_repeat_after_sec.restype = i64
        lib.orc_cron_repeat_after_sec.argtypes = [C.POINTER(OrcCron), i64]
        lib.orc_parse_duration.restype = C.c_int
        lib.orc_parse_duration.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(i64)]
        lib.orc_classify.restype = C.c_int
        lib.orc_classify.argtypes = [C.c_void_p, C.c_void_p]
        lib.orc_remedy_is_empty.restype = C.c_int
        lib.orc_remedy_is_empty.argtypes = [C.c_size_t, C.c_int, i64, C.c_int]
        lib.orc_sweep.restype = C.c_int
        lib.orc_sweep.argtypes = [C.POINTER(OrcCols), u64, u64, i64, u32, u64, C.c_void_p,
                                  C.c_void_p, u64, C.POINTER(u64), C.POINTER(OrcStats)]
        lib.orc_sweep_mt.restype = C.c_int
        lib.orc_sweep_mt.argtypes = lib.orc_sweep.argtypes + [C.c_int]
        lib.orc_faithful_eval.restype = u64
        lib.orc_faithful_eval.argtypes = [C.c_void_p, u64, i64]
        lib.orc_key.restype = u64
        lib.orc_key.argtypes = [u64, u64, u64]
        lib.orc_civil_from_unix.restype = None
        lib.orc_civil_from_unix.argtypes = [i64, C.POINTER(i32 * 6)]
        _lib = lib
    return _lib


def cron_parse(spec):
    """-> (rc, OrcCron, message)"""
    raw = spec if isinstance(spec, (bytes, bytearray)) else spec.encode("utf-8", "surrogateescape")
    out = OrcCron()
    err = C.create_string_buffer(256)
    rc = load().orc_cron_parse(bytes(raw), len(raw), C.byref(out), err, len(err))
    return rc, out, err.value.decode("utf-8", "replace")


def cols_struct(cols: dict) -> OrcCols:
    s = OrcCols()
    for name, dt in COLUMNS:
        a = cols[name]
        assert a.dtype == np.dtype(dt) and a.flags["C_CONTIGUOUS"], name
        setattr(s, name, a.ctypes.data)
    return s


def sweep(cols: dict, T: int, mode: int = 0, seed: int = 0, shard_base: int = 0, threads: int = 1,
          buffers=None):
    """Run the oracle over numpy SoA columns IN PLACE.
    Returns (global idx u64[n], action u32[n], stats dict).  With `buffers=(idx u64[cap],
    act u32[cap])` the lists are written there and VIEWS are returned (no allocation, no copy:
    the form the timed CPU baseline uses)."""
    n = len(cols["flags"])
    if buffers is None:
        idx = np.empty(n, dtype=np.uint64)
        act = np.empty(n, dtype=np.uint32)
    else:
        idx, act = buffers
        assert idx.dtype == np.uint64 and act.dtype == np.uint32 and len(idx) == len(act)
    cap = len(idx)
    cnt = u64(0)
    st = OrcStats()
    cs = cols_struct(cols)
    lib = load()
    if threads <= 1:
        rc = lib.orc_sweep(C.byref(cs), n, shard_base, T, mode, seed, idx.ctypes.data,
                           act.ctypes.data, cap, C.byref(cnt), C.byref(st))
    else:
        rc = lib.orc_sweep_mt(C.byref(cs), n, shard_base, T, mode, seed, idx.ctypes.data,
                              act.ctypes.data, cap, C.byref(cnt), C.byref(st), threads)
    assert rc == 0, rc
    if buffers is not None:
        return idx[:cnt.value], act[:cnt.value], st.as_dict()
    return idx[:cnt.value].copy(), act[:cnt.value].copy(), st.as_dict()
