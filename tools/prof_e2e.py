"""Where the host-buffer (e2e) step spends its time: post_result, tick, glue — the same loop as
bench.py's e2e block at N=1.  Prints one JSON object (kept under profiles/)."""
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools", "amgen")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import amgen  # noqa: E402

am = importlib.import_module("active-monitor_b200")
n, T0 = 10_000_000, amgen.T0_MON_0915
cols = amgen.fill(2, 2, 0, n, T0, am.load().am_healthcheck_classify)
s = am.Sweep(capacity=n)
s.load_range(0, cols)
ok = np.full(n, am.PHASE_SUCCEEDED, np.uint8)
sel = np.empty(n, np.uint64)
out = {}
tick_no = 0
for api in ("view", "wide"):
    prev = None
    acc = {"post": 0.0, "tick": 0.0, "glue": 0.0, "device_ms": 0.0}
    idx_h, act_h = np.empty(n, np.uint64), np.empty(n, np.uint32)
    reps = 100
    for k in range(reps + 20):
        t0 = time.perf_counter()
        if prev is not None and len(prev):
            s.post_result(prev, ok[: len(prev)])
        t1 = time.perf_counter()
        if api == "view":
            vi, va, st = s.tick_view(T0 + tick_no, mode=am.SWEEP_FULL_SCAN)
            t2 = time.perf_counter()
            prev = amgen.select_submitted_view(vi, va, sel)
        else:
            gi, ga, st = s.tick(T0 + tick_no, mode=am.SWEEP_FULL_SCAN, buffers=(idx_h, act_h))
            t2 = time.perf_counter()
            prev = amgen.select_submitted(gi, ga, 0, sel)
        t3 = time.perf_counter()
        tick_no += 1
        if k >= 20:
            acc["post"] += t1 - t0; acc["tick"] += t2 - t1; acc["glue"] += t3 - t2; acc["device_ms"] += s.last_kernel_ms
    r = {k: round(v / reps * 1e3, 4) for k, v in acc.items()}
    r["device_ms"] = round(acc["device_ms"] / reps, 4)
    r["step_ms"] = round(r["post"] + r["tick"] + r["glue"], 4)
    r["submitted_per_tick"], r["emitted_per_tick"] = int(len(prev)), int(st["n_emitted"])
    out[api] = r
# where the device time of a host tick goes: the same loop with per-kernel events (sweep | scan+expand+publish),
# the list written to mapped host memory (tick_view) vs left in HBM (tick_device)
import torch  # noqa: E402
dev = torch.device("cuda", 0)
d_idx = torch.empty(n, dtype=torch.int32, device=dev)
d_act = torch.empty(n, dtype=torch.uint8, device=dev)
d_cnt = torch.zeros(1, dtype=torch.int32, device=dev)
s.set_profiling(True)
for where in ("host_mapped", "hbm"):
    a_ms = b_ms = 0.0
    for k in range(60):
        if prev is not None and len(prev):
            s.post_result(prev, ok[: len(prev)])
        if where == "host_mapped":
            vi, va, st = s.tick_view(T0 + tick_no, mode=am.SWEEP_FULL_SCAN)
            prev = amgen.select_submitted_view(vi, va, sel).copy()
        else:
            s.tick_device(T0 + tick_no, am.SWEEP_FULL_SCAN, d_idx.data_ptr(), d_act.data_ptr(), n, d_cnt.data_ptr(), 0, 0)
            torch.cuda.synchronize()
            c = int(d_cnt.item())
            prev = amgen.select_submitted_view(d_idx[:c].cpu().numpy().view(np.uint32), d_act[:c].cpu().numpy(), sel).copy()
        tick_no += 1
        a, b = s.last_profile()
        if k >= 10:
            a_ms += a; b_ms += b
    out["kernels_" + where] = {"sweep_ms": round(a_ms / 50, 4), "scan_expand_publish_ms": round(b_ms / 50, 4)}
s.set_profiling(False)
print(json.dumps({"what": "ms per e2e step, 10 M records config 2, consecutive seconds, host-closed loop, FULL_SCAN",
                  "post": "am_sweep_post_result (u64 slots + u8 phases -> pinned staging, chunked H2D)",
                  "tick": "am_sweep_tick_view | am_sweep_tick (drain tail + kernels + one sync [+ widening])",
                  "glue": "harness walk over the list -> next tick's slots (amgen.select_submitted*)",
                  "device_ms": "CUDA events around the tick's kernels (inside `tick`)", **out}))
