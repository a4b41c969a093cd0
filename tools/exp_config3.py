"""Why does the config-3 sweep take ~260 us by CUDA events in bench.py and ~210 us under ncu?
Times sweep_tick_kernel (events inside the library) on the 10 M config-3 population under different
pre-conditions of L2 / DRAM state.  Prints one JSON object."""
import importlib
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools", "amgen")):
    sys.path.insert(0, p)
import amgen  # noqa: E402

am = importlib.import_module("active-monitor_b200")
n, T0 = 10_000_000, amgen.T0_MON_0915
dev = torch.device("cuda", 0)
cols = amgen.fill(3, 3, 0, n, T0, am.load().am_healthcheck_classify)
s = am.Sweep(capacity=n)
s.load_range(0, cols)
d_idx = torch.empty(n, dtype=torch.int32, device=dev)
d_act = torch.empty(n, dtype=torch.uint8, device=dev)
d_cnt = torch.zeros(1, dtype=torch.int32, device=dev)
d_st = torch.zeros(16, dtype=torch.int64, device=dev)
MUT = ["flags", "finished_at", "success", "failed", "remedy_success", "remedy_failed", "remedy_total", "remedy_finished_at"]


def view(name):
    dt = torch.int64 if name in ("finished_at", "remedy_finished_at") else torch.int32
    class A:
        pass
    a = A()
    a.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8" if dt == torch.int64 else "<i4",
                                  "data": (s.column_ptr(name), False), "version": 3, "strides": (8 if dt == torch.int64 else 4,)}
    return torch.as_tensor(a, device=dev)


live = {m: view(m) for m in MUT}
torch.cuda.synchronize()
snap = {m: v.clone() for m, v in live.items()}
flush = torch.zeros(64 << 20, dtype=torch.int32, device=dev)
sink = torch.zeros((), dtype=torch.int64, device=dev)
s.set_profiling(True)


def tick():
    s.tick_device(T0, 0, d_idx.data_ptr(), d_act.data_ptr(), n, d_cnt.data_ptr(), d_st.data_ptr(), 0)
    return s.last_profile()


def run(pre, reps=12):
    a, b = [], []
    for _ in range(reps):
        pre()
        x, y = tick()
        a.append(x * 1e3); b.append(y * 1e3)
    return {"sweep_us_median": round(statistics.median(a[2:]), 1), "sweep_us_min": round(min(a[2:]), 1),
            "rest_us_median": round(statistics.median(b[2:]), 1)}


def restore():
    for m, v in snap.items():
        live[m].copy_(v)


def pre_restore():
    restore()


def pre_restore_sync():
    restore(); torch.cuda.synchronize()


def pre_restore_readflush():
    restore(); sink.copy_(flush.sum())


def pre_restore_writeflush():
    restore(); flush.zero_()


def pre_reload():
    s.load_range(0, cols)


import threading
import time


def pre_restore_idle(ms):
    def f():
        restore(); torch.cuda.synchronize(); time.sleep(ms * 1e-3)
    return f


def clocks_while(fn, seconds=1.0):
    """SM clock / power / throttle reasons polled through NVML (~1 kHz) while `fn` runs in a loop"""
    import pynvml
    pynvml.nvmlInit()
    hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
    smp, stop = [], threading.Event()

    def poll():
        while not stop.is_set():
            smp.append((pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(hnd) / 1000.0,
                        pynvml.nvmlDeviceGetCurrentClocksEventReasons(hnd)))
            time.sleep(0.001)
    th = threading.Thread(target=poll)
    th.start()
    t0, times = time.time(), []
    while time.time() - t0 < seconds:
        times.append(fn())
    torch.cuda.synchronize()
    stop.set(); th.join()
    mhz = sorted(x[0] for x in smp)
    reasons = 0
    for x in smp:
        reasons |= x[2]
    return {"samples": len(smp), "sm_mhz_min": mhz[0], "sm_mhz_median": mhz[len(mhz) // 2], "sm_mhz_max": mhz[-1],
            "power_w_max": round(max(x[1] for x in smp), 1), "reasons_or": hex(reasons),
            "sweep_us_median": round(statistics.median(times[3:]), 1) if len(times) > 3 else None}


def loop_dense():
    restore()
    return tick()[0] * 1e3


def loop_plain():
    return tick()[0] * 1e3


out = {"what": "sweep_tick_kernel<0,1> on 10 M config-3 records, CUDA events, by what ran before it",
       "restore_sync_idle_2ms": run(pre_restore_idle(2)),
       "restore_sync_idle_20ms": run(pre_restore_idle(20)),
       "restore_sync_idle_200ms": run(pre_restore_idle(200), 8),
       "clocks_back_to_back_dense": clocks_while(loop_dense),
       "reload_columns_from_host": run(pre_reload, 6),
       "restore_d2d": run(pre_restore),
       "restore_d2d_then_sync": run(pre_restore_sync),
       "restore_then_read_256MB": run(pre_restore_readflush),
       "restore_then_write_256MB": run(pre_restore_writeflush)}
# the same tick once the results are applied (nothing pending): the sparse shape
out["no_pending_results"] = run(lambda: None)
out["clocks_back_to_back_plain"] = clocks_while(loop_plain)
print(json.dumps(out))
