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


out = {"what": "sweep_tick_kernel<0,1> on 10 M config-3 records, CUDA events, by what ran before it",
       "reload_columns_from_host": run(pre_reload, 6),
       "restore_d2d": run(pre_restore),
       "restore_d2d_then_sync": run(pre_restore_sync),
       "restore_then_read_256MB": run(pre_restore_readflush),
       "restore_then_write_256MB": run(pre_restore_writeflush)}
# the same tick once the results are applied (nothing pending): the sparse shape
out["no_pending_results"] = run(lambda: None)
print(json.dumps(out))
