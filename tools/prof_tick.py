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
ap.add_argument("--e2e", action="store_true", help="the e2e shape: consecutive seconds, FULL_SCAN, every check "
                "the previous tick submitted is posted as Succeeded before the tick (host list via tick_view)")
a = ap.parse_args()
am = importlib.import_module("active-monitor_b200")
T0 = amgen.T0_MON_0915
cols = amgen.fill(a.config, a.config, 0, a.n, T0, am.load().am_healthcheck_classify)
import torch  # noqa: E402
dev = torch.device("cuda", 0)
d_idx = torch.empty(a.n, dtype=torch.int32, device=dev)
d_act = torch.empty(a.n, dtype=torch.uint8, device=dev)
d_cnt = torch.zeros(1, dtype=torch.int32, device=dev)
d_st = torch.zeros(16, dtype=torch.int64, device=dev)
with am.Sweep(capacity=a.n) as s:
    s.load_range(0, cols)
    s.set_profiling(True)
    if a.e2e:
        import numpy as np
        ok = np.full(a.n, am.PHASE_SUCCEEDED, dtype=np.uint8)
        sel = np.empty(a.n, dtype=np.uint64)
        prev = None
        for k in range(a.ticks):
            if prev is not None and len(prev):
                s.post_result(prev, ok[: len(prev)])
            vi, va, st = s.tick_view(T0 + k, mode=am.SWEEP_FULL_SCAN)
            prev = amgen.select_submitted_view(vi, va, sel)
            ka, kb = s.last_profile()
            print(k, st["n_emitted"], len(prev), f"sweep {ka * 1e3:.1f} us  scan+expand+publish {kb * 1e3:.1f} us")
        sys.exit(0)
    for k in range(a.ticks):
        if a.config == 3 and k:  # re-arm the pending results so every tick does the same work
            s.load_range(0, cols)
        # device-resident tick (the list stays in HBM), as in bench.py's timed loop
        s.tick_device(T0 + k * a.dt, a.mode, d_idx.data_ptr(), d_act.data_ptr(), a.n, d_cnt.data_ptr(), d_st.data_ptr(), 0)
        ka, kb = s.last_profile()
        st = dict(zip(am.abi.STAT_FIELDS, d_st.cpu().tolist()))
        print(k, st["n_emitted"], st["n_submit_hc"], f"sweep {ka * 1e3:.1f} us  scan+expand+publish {kb * 1e3:.1f} us")
