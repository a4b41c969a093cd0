"""ctypes binding of tools/amgen/libamgen.so — deterministic synthetic
HealthCheck populations.  Neutral tooling: the classify
function (product's or oracle's) is passed in as a C function pointer."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libamgen.so")

# named time zones of the generated populations, in the order amgen_fill introduces them to the
# implementation behind the classify function (zone ids are handed out in order of first appearance)
ZONES = ("America/New_York", "Europe/Paris", "Asia/Kolkata", "Asia/Kathmandu", "Australia/Lord_Howe",
         "America/St_Johns")
T0_MON_0915 = 1789982100   # 2026-09-21 09:15:00 UTC, Monday  
T0_OCT_1 = 1790812800      # 2026-10-01 00:00:00 UTC, Thursday (hour/day/month boundary)
T0_DAY_START = 1789948800  # 2026-09-21 00:00:00 UTC           (config 5)

COLUMNS = [("minute", np.uint64), ("hour", np.uint64), ("dom", np.uint64), ("month", np.uint64),
           ("dow", np.uint64), ("ras", np.int32), ("flags", np.uint32), ("finished_at", np.int64),
           ("runs_limit", np.int32), ("reset_interval", np.int32), ("success", np.int32),
           ("failed", np.int32), ("remedy_success", np.int32), ("remedy_failed", np.int32),
           ("remedy_total", np.int32), ("remedy_finished_at", np.int64)]


class _Cols(C.Structure):
    _fields_ = [(n, C.c_void_p) for n, _ in COLUMNS]


class HealthCheckC(C.Structure):  # == am_healthcheck_t
    _fields_ = [("repeat_after_sec", C.c_int64), ("cron", C.c_void_p), ("cron_len", C.c_size_t),
                ("has_resource", C.c_int32), ("has_remedy", C.c_int32),
                ("remedy_runs_limit", C.c_int64), ("remedy_reset_interval", C.c_int64),
                ("finished_at", C.c_int64), ("remedy_finished_at", C.c_int64),
                ("finished_at_set", C.c_int32), ("remedy_finished_at_set", C.c_int32),
                ("success_count", C.c_int64), ("failed_count", C.c_int64),
                ("remedy_success_count", C.c_int64), ("remedy_failed_count", C.c_int64),
                ("remedy_total_runs", C.c_int64), ("fail_p8", C.c_uint32), ("timer_armed", C.c_uint32)]


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} missing: run python active-monitor_b200/build.py")
        lib = C.CDLL(LIB_PATH)
        lib.amgen_fill.restype = C.c_int64
        lib.amgen_fill.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int64,
                                   C.c_void_p, C.POINTER(_Cols), C.c_int]
        lib.amgen_healthchecks.restype = None
        lib.amgen_healthchecks.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int64,
                                           C.c_void_p, C.c_void_p, C.c_void_p]
        lib.amgen_str_stride.restype = C.c_int
        lib.amgen_select_submitted.restype = C.c_uint64
        lib.amgen_select_submitted.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
        lib.amgen_e2e_closed_loop.restype = C.c_int
        lib.amgen_e2e_closed_loop.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_uint32, C.c_uint64,
                                              C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int,
                                              C.POINTER(C.c_double), C.POINTER(C.c_double * 4), C.POINTER(C.c_uint64),
                                              C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        lib.amgen_select_submitted_view.restype = C.c_uint64
        lib.amgen_select_submitted_view.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        lib.amgen_key.restype = C.c_uint64
        lib.amgen_key.argtypes = [C.c_uint64] * 3
        _lib = lib
    return _lib


def fill(config: int, seed: int, first: int, n: int, T0: int, classify_fn, threads: int = 0) -> dict:
    """SoA columns (numpy) for records [first, first+n).  `classify_fn` is a
    ctypes function object: lib.am_healthcheck_classify or lib.orc_classify."""
    cols = {name: np.zeros(n, dtype=dt) for name, dt in COLUMNS}
    cs = _Cols()
    for name, _ in COLUMNS:
        setattr(cs, name, cols[name].ctypes.data)
    if threads <= 0:
        threads = min(os.cpu_count() or 1, 64)
    fnp = C.cast(classify_fn, C.c_void_p)
    bad = load().amgen_fill(config, seed, first, n, T0, fnp, C.byref(cs), threads)
    if bad:
        raise RuntimeError(f"amgen_fill: {bad} unexpected classify failures")
    return cols


def healthchecks(config: int, seed: int, first: int, n: int, T0: int):
    """The HealthChecks themselves: (array of HealthCheckC, list[str] cron, post_flags u32[n]).
    The returned ctypes array keeps its string pool alive via ._pool."""
    lib = load()
    stride = lib.amgen_str_stride()
    hcs = (HealthCheckC * n)()
    pool = C.create_string_buffer(n * stride)
    post = np.zeros(n, dtype=np.uint32)
    lib.amgen_healthchecks(config, seed, first, n, T0, C.byref(hcs), pool, post.ctypes.data)
    hcs._pool = pool
    raw = pool.raw
    crons = []
    for k in range(n):
        ln = hcs[k].cron_len
        crons.append(raw[k * stride:k * stride + ln].decode("utf-8"))
    return hcs, crons, post


def select_submitted(idx: np.ndarray, act: np.ndarray, base: int, out: np.ndarray) -> np.ndarray:
    """local slots of the entries whose action has AM_ACT_SUBMIT_HC (compiled loop: stands
    in for the Go shim walking the tick's result); `out` must hold len(idx) u64."""
    m = load().amgen_select_submitted(idx.ctypes.data, act.ctypes.data, len(idx), base, out.ctypes.data)
    return out[:m]


def select_submitted_view(idx_local: np.ndarray, act: np.ndarray, out: np.ndarray) -> np.ndarray:
    """local slots (u64) of the entries of an am_tick_view_t whose action carries SUBMIT_HC"""
    if len(idx_local) == 0:
        return out[:0]
    m = load().amgen_select_submitted_view(idx_local.ctypes.data, act.ctypes.data, len(idx_local), out.ctypes.data)
    return out[:m]


def e2e_closed_loop(product_lib, sweep_handle, T_first: int, mode: int, warm: int, steps: int, capacity: int,
                    workers: int = 1, gather_handle=None) -> dict:
    """bench.py's e2e loop in compiled code (amgen_e2e_closed_loop): am_sweep_tick_view, then `workers`
    consumer threads walking the list in pieces and posting through am_sweep_post_result of
    `product_lib` (a ctypes CDLL of the C-ABI) on `sweep_handle`."""
    slots = np.empty(capacity, dtype=np.uint64)
    ok = np.full(capacity, 1, dtype=np.uint8)  # AM_PHASE_SUCCEEDED
    sec, split = C.c_double(), (C.c_double * 4)()
    h2d, d2h, ne, ns = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
    # one GPU: am_sweep_tick_view on the sweep; a multi-GPU shard: am_gather_tick_view on the (bound) exchange
    tick_fn = product_lib.am_gather_tick_view if gather_handle is not None else product_lib.am_sweep_tick_view
    rc = load().amgen_e2e_closed_loop(C.cast(product_lib.am_sweep_post_result, C.c_void_p),
                                      C.cast(tick_fn, C.c_void_p), sweep_handle, gather_handle, T_first, mode,
                                      warm, steps, slots.ctypes.data, capacity, ok.ctypes.data, workers, C.byref(sec),
                                      C.byref(split), C.byref(h2d), C.byref(d2h), C.byref(ne), C.byref(ns))
    if rc <= 0:
        raise RuntimeError(f"amgen_e2e_closed_loop: C-ABI call failed with {rc}")
    return {"seconds": sec.value, "post_s": split[0], "tick_s": split[1], "walk_s": split[2], "consumer_s": split[3],
            "workers": rc, "h2d_bytes": h2d.value, "d2h_bytes": d2h.value, "last_emitted": ne.value,
            "last_submitted": ns.value}
