"""Where the host-buffer (e2e) step spends its time: post_result, tick, glue."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools", "amgen")):
    sys.path.insert(0, p)
import numpy as np
import amgen
am = importlib.import_module("active-monitor_b200")
n, T0 = 10_000_000, amgen.T0_MON_0915
cols = amgen.fill(2, 2, 0, n, T0, am.load().am_healthcheck_classify)
s = am.Sweep(capacity=n)
s.load_range(0, cols)
idx_h, act_h = np.empty(n, np.uint64), np.empty(n, np.uint32)
ok = np.full(n, am.PHASE_SUCCEEDED, np.uint8)
sel = np.empty(n, np.uint64)
prev = None
acc = {"post": 0.0, "tick": 0.0, "glue": 0.0, "kernel_ms": 0.0}
for k in range(60):
    t0 = time.perf_counter()
    if prev is not None:
        s.post_result(prev, ok[: len(prev)])
    t1 = time.perf_counter()
    gi, ga, st = s.tick(T0 + k, mode=am.SWEEP_FULL_SCAN, buffers=(idx_h, act_h))
    t2 = time.perf_counter()
    prev = amgen.select_submitted(gi, ga, 0, sel)
    t3 = time.perf_counter()
    if k >= 10:
        acc["post"] += t1 - t0; acc["tick"] += t2 - t1; acc["glue"] += t3 - t2; acc["kernel_ms"] += s.last_kernel_ms
print({k: round(v / 50 * 1e3, 3) for k, v in acc.items()}, "ms per step;", len(prev), "submitted,", len(gi), "emitted")
