"""Minimal driver for ncu captures of the sweep kernel: BASELINE configs[1]
(10 M records, config-2 mix) or configs[2] (--config 3), a handful of ticks."""
import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools", "amgen")):
    sys.path.insert(0, p)
import amgen  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--n", type=int, default=10_000_000)
ap.add_argument("--ticks", type=int, default=6)
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--dt", type=int, default=0, help="seconds between ticks (0 = same tick T0)")
a = ap.parse_args()
am = importlib.import_module("active-monitor_b200")
T0 = amgen.T0_MON_0915
cols = amgen.fill(a.config, a.config, 0, a.n, T0, am.load().am_healthcheck_classify)
with am.Sweep(capacity=a.n) as s:
    s.load_range(0, cols)
    s.set_profiling(True)
    for k in range(a.ticks):
        if a.config == 3 and k:  # re-arm the pending results so every tick does the same work
            s.load_range(0, cols)
        idx, act, st = s.tick(T0 + k * a.dt, mode=a.mode)
        ka, kb = s.last_profile()
        print(k, st["n_emitted"], st["n_submit_hc"], f"pair {s.last_kernel_ms * 1e3:.1f} us  sweep {ka * 1e3:.1f} us  compact {kb * 1e3:.1f} us")
