#!/usr/bin/env python
"""bench.py — HealthCheck schedule evaluations/sec of the B200 sweep.

One "step" = one tick of the hot path (schedule ladder + remedy state machine,
SURVEY.md Appendix B.3) over the whole resident record array.  Workload at N=1:
BASELINE.json configs[1] — 10 M HealthChecks, mixed 5-field cron + repeatAfterSec,
seed 2 (tools/amgen).  At N>1 every rank owns an index-range shard of the same
size (weak scaling) and the step ends with the all-gather of the due lists.

  value     device-resident: state and outputs stay in HBM, CUDA events on the
            launching stream, max over ranks.
  e2e       through the host C-ABI call a cgo shim makes (am_sweep_post_result +
            am_sweep_tick): results posted from host memory, due list copied
            back to host memory, every step.
  roofline  algorithmic bytes of the sweep kernel / its measured duration,
            against MEASURED_PEAKS.json's HBM copy bandwidth.
  cpu_baseline / --impl reference
            the CPU oracle (our restatement of hcc.go + robfig/cron v3.0.1; the
            Go reference cannot be executed here) on the box's host cores.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools", "amgen")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "healthcheck_schedule_evals_per_sec"
UNIT = "evals/s"
N_PER_GPU = 10_000_000
CONFIG, SEED = 2, 2
B_READ = 56          # algorithmic bytes read per evaluation, schedule-only
B_EMIT = 5           # u32 index + u8 action per emitted record (this layout)
B_STOP = 12          # flags + finishedAt written when "Stopped" is first reported
# dram__bytes_read.sum + dram__bytes_write.sum of ONE sweep_tick_kernel launch on this exact
# workload (10 M records, config 2, tick T0), from the `ncu --set full` capture summarised in
# profiles/r01_sweep_config2_ncu_full.csv (560.36 MB read + 3.86 MB written; part of the
# 16.7 MB of segment writes is still dirty in L2 when the kernel ends)
NCU_TRAFFIC_BYTES_10M_CONFIG2 = 564_213_760


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons while the GPU is under the bench load."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def n_since(self, t0):
        return sum(1 for t, _ in self.rows if t >= t0)

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for t, line in self.rows:
            if t < t0 or t > t1:
                continue
            p = [x.strip() for x in line.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_sweep_rate(cols, T0, seconds, threads, oracle_c):
    """The same step as the GPU arm — the single tick T0 over the whole
    population — repeated for about `seconds`; returns (evals/s, ticks run)."""
    import numpy as np
    n = len(cols["flags"])
    work = {k: v.copy() for k, v in cols.items()}
    # output lists preallocated once, as the GPU e2e arm's host buffers are
    bufs = (np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint32))
    oracle_c.sweep(work, T0, threads=threads, buffers=bufs)  # warm caches / page in / first "Stopped" reports
    t0 = time.perf_counter()
    k = 0
    while True:
        k += 1
        oracle_c.sweep(work, T0, threads=threads, buffers=bufs)
        if time.perf_counter() - t0 >= seconds or k >= 2000:
            break
    dt = time.perf_counter() - t0
    return n * k / dt, k


def run_reference(args, rank):
    """--impl reference: the CPU path on this box's host cores.  The reference is
    Go (robfig/cron un-vendored, no Go toolchain): this times oracle/ — our C
    restatement of the same decisions — with every host thread."""
    if rank != 0:
        return
    import amgen
    import oracle_c
    threads = os.cpu_count() or 1
    n = min(N_PER_GPU, args.n)
    cols = amgen.fill(CONFIG, SEED, 0, n, amgen.T0_MON_0915, oracle_c.load().orc_classify)
    import numpy as np
    work = {k: v.copy() for k, v in cols.items()}
    # output lists preallocated once (host buffers of the caller, as in the GPU e2e arm)
    bufs = (np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint32))
    for w in range(args.warmup):
        oracle_c.sweep(work, amgen.T0_MON_0915, threads=threads, buffers=bufs)
    t0 = time.perf_counter()
    for k in range(args.steps):
        oracle_c.sweep(work, amgen.T0_MON_0915, threads=threads, buffers=bufs)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = f"{n} records x {args.steps} ticks (whole config-2 population each step)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": f"BASELINE configs[1]: {n} HealthChecks, mixed 5-field cron + "
                               "repeatAfterSec, seed 2, one tick per step", "records": n,
                   "note": "CPU oracle port of hcc.go + robfig/cron v3.0.1; Go reference not runnable here"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "chunk per thread, two passes (evaluate + count, then write at the prefix), "
                                 "parked worker threads, preallocated output lists"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=N_PER_GPU, help="records per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--wire", default="auto", choices=["auto", "plain", "c3", "bm"],
                    help="peer gather wire format: plain = u32 index + u8 action (5 B/entry), "
                         "c3 = u16 group offset + u8 action + per-group counts (3 B/entry), expanded on the "
                         "receiver; auto = c3 from 4 GPUs up (NVLink-volume-bound), plain below "
                         "(profiles/r01_scaling.md); bm = EXPERIMENTAL bitmap format (1 bit per record + "
                         "non-default actions), never chosen by auto")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="N>1: NVLink peer-write kernel (csrc/gather.cu) or the padded NCCL all-gather")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.wire == "auto":
        args.wire = "c3" if world >= 4 else "plain"
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import amgen
    am = importlib.import_module("active-monitor_b200")
    gather = importlib.import_module("active-monitor_b200.gather")
    lib = am.load()  # no fallback: raises if the CUDA library is missing

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the sweep has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    T0 = amgen.T0_MON_0915
    n = args.n
    base = rank * n
    cols = amgen.fill(CONFIG, SEED, base, n, T0, lib.am_healthcheck_classify)
    sweep = am.Sweep(capacity=n, device=local_rank, shard_base=base)
    sweep.load_range(0, cols)

    nbuf = 2 if world > 1 else 1
    d_idx = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(nbuf)]
    d_act = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(nbuf)]
    d_cnt = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(nbuf)]
    d_st = torch.zeros(16, dtype=torch.int64, device=dev)
    stream = torch.cuda.Stream(device=dev)  # explicit stream: kernels, events and NCCL all on it
    torch.cuda.set_stream(stream)
    peer = gstream = None
    if world > 1 and args.gather == "peer":
        # global indices fit u32 here (records_total < 2^32): 5 B per entry on the wire.
        # All ranks must take the same path: agree on whether CUDA-IPC peer mapping worked.
        ok = 1
        try:
            peer = gather.PeerGather(local_rank, cap_total=n * world,
                                     idx_bytes=4 if n * world < (1 << 32) else 8,
                                     shard=(base, n) if args.wire in ("c3", "bm") else None,
                                     wire="bm" if args.wire == "bm" else "c3")
        except Exception as e:  # noqa: BLE001
            print(f"[bench] rank {rank}: peer-write gather unavailable ({e}); using NCCL", file=sys.stderr)
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            if peer is not None:
                peer.close()
            peer = None
    if peer is not None:
        gstream = torch.cuda.Stream(device=dev, priority=-1)  # exchange CTAs are placed ahead of pending sweep CTAs
        ev_sweep = [torch.cuda.Event() for _ in range(2)]
        ev_gather = [torch.cuda.Event() for _ in range(2)]
    step_no = [0]

    def step(_k):
        # BASELINE configs[1] is ONE tick: every step is that tick (T0, on the minute,
        # open loop => idempotent), over inputs 4.4x larger than L2
        b = step_no[0] % nbuf
        first_use = step_no[0] < nbuf
        step_no[0] += 1
        if peer is not None:
            # tick k's exchange (NVLink) overlaps tick k+1's sweep (HBM): two streams, two
            # output buffers; a buffer is rewritten only after the exchange that read it retired
            if not first_use:
                stream.wait_event(ev_gather[b])
            sweep.tick_device(T0, 0, d_idx[b].data_ptr(), d_act[b].data_ptr(), n, d_cnt[b].data_ptr(),
                              d_st.data_ptr(), stream.cuda_stream)
            ev_sweep[b].record(stream)
            gstream.wait_event(ev_sweep[b])
            # one kernel: counts + lists written into every peer over NVLink
            peer.push(d_idx[b].data_ptr(), d_act[b].data_ptr(), d_cnt[b].data_ptr(), base, gstream.cuda_stream)
            ev_gather[b].record(gstream)
            return None
        sweep.tick_device(T0, 0, d_idx[b].data_ptr(), d_act[b].data_ptr(), n, d_cnt[b].data_ptr(),
                          d_st.data_ptr(), stream.cuda_stream)
        if world > 1:
            return gather.allgather_due(d_idx[b], d_act[b], d_cnt[b], base)
        return None

    def drain():
        """make `stream` wait for every exchange still in flight on the gather stream"""
        if peer is not None:
            for e in ev_gather[: min(step_no[0], 2)]:
                stream.wait_event(e)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    # Every rank must run the SAME number of steps (each step ends in a cross-GPU
    # exchange at N>1), so warm-up and the trailing clock-sampling load are counted,
    # never timed: at least W steps and ~0.4 s of load so clocks settle.
    w = max(args.warmup, 3000)
    for i in range(w):
        step(-1 - i)
        if i % 256 == 255:
            torch.cuda.synchronize()
    drain()
    barrier()
    launches0 = sweep.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_load0 = time.perf_counter()
    ev0.record(stream)
    for k in range(args.steps):
        step(k)
    drain()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = sweep.launch_count - launches0 + (
        0 if peer is None else args.steps * (2 if peer.compressed else 1))
    stats = dict(zip(am.abi.STAT_FIELDS, [int(v) for v in d_st.cpu().tolist()]))
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # keep the identical load running ~0.8 s more (same count on every rank) so the
    # 100 ms nvidia-smi sampler sees the GPU under exactly this load
    extra = min(20000, max(64, int(0.8 / max(ms / args.steps * 1e-3, 1e-6))))
    for k in range(extra):
        step(k)
        if k % 256 == 255:
            torch.cuda.synchronize()
    barrier()
    t_load1 = time.perf_counter()
    clocks = sampler.stop(t_load0, t_load1) if sampler else None

    # ---- per-kernel durations for the roofline (CUDA events on the launching stream
    #      around each of the tick's two kernels, inside the library) ----
    sweep.set_profiling(True)
    ks, kc = [], []
    kreps = max(20, min(args.steps, 200))
    for k in range(kreps):
        sweep.tick_device(T0, 0, d_idx[0].data_ptr(), d_act[0].data_ptr(), n, d_cnt[0].data_ptr(),
                          d_st.data_ptr(), stream.cuda_stream)
        a_ms, b_ms = sweep.last_profile()
        ks.append(a_ms); kc.append(b_ms)
    sweep.set_profiling(False)
    k_ms = statistics.mean(ks)
    c_ms = statistics.mean(kc)
    # sweep_tick_kernel: reads the 56 B/record of schedule columns, writes 5 B per emitted
    # record into its tile segment (+12 B when "Stopped" is first reported)
    alg_bytes = n * B_READ + stats["n_emitted"] * B_EMIT + stats["n_stopped"] * B_STOP
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9

    # ---- e2e: the call a cgo shim makes, host buffers both ways, real cadence ----
    # consecutive one-second ticks; the host closes the loop: every check the previous
    # tick submitted is posted back as Succeeded (H2D), then the tick runs and its due
    # list is copied to host memory (D2H).  AM_SWEEP_FULL_SCAN keeps the device work per
    # step identical to the `value` step (all 56 B/record read on every tick).
    e2e = None
    if True:
        idx_h = np.empty(n, dtype=np.uint64)
        act_h = np.empty(n, dtype=np.uint32)
        ok_phase = np.full(n, am.PHASE_SUCCEEDED, dtype=np.uint8)
        sel = np.empty(n, dtype=np.uint64)
        prev = None
        h2d = d2h = 0
        reps = max(10, min(args.steps, 120))
        tick_no = 0
        for phase_name in ("warm", "timed"):
            if phase_name == "timed":
                barrier()
                t0 = time.perf_counter()
                h2d = d2h = 0
            for k in range(5 if phase_name == "warm" else reps):
                if prev is not None and len(prev):
                    sweep.post_result(prev, ok_phase[: len(prev)])
                    h2d += len(prev) * 8
                gi, ga, st = sweep.tick(T0 + tick_no, mode=am.SWEEP_FULL_SCAN, buffers=(idx_h, act_h))
                tick_no += 1
                # the checks just submitted, as local slots (compiled loop standing in for the
                # Go shim's walk over the tick's result, hcc.go:269-288)
                prev = amgen.select_submitted(gi, ga, base, sel)
                d2h += len(gi) * 5 + 128
            if phase_name == "timed":
                barrier()
                dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": n * world * reps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d // reps,
               "d2h_bytes_per_step": d2h // reps, "ms_per_step": dt / reps * 1e3, "steps": reps,
               "api": "am_sweep_post_result + am_sweep_tick (host buffers), consecutive 1 s ticks, "
                      "host-closed loop, AM_SWEEP_FULL_SCAN"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import oracle_c
        threads = os.cpu_count() or 1
        host_cols = amgen.fill(CONFIG, SEED, 0, n, T0, oracle_c.load().orc_classify)
        v_mt, k_mt = cpu_sweep_rate(host_cols, T0, args.cpu_seconds, threads, oracle_c)
        v_1t, k_1t = cpu_sweep_rate(host_cols, T0, min(4.0, args.cpu_seconds), 1, oracle_c)
        # B1 "faithful shape": re-parse + Next() per evaluation, 100k-record subsample
        hcs, _, _ = amgen.healthchecks(CONFIG, SEED, 0, 100_000, T0)
        import ctypes as C
        t0 = time.perf_counter()
        reps_b1 = 0
        while time.perf_counter() - t0 < 2.0:
            oracle_c.load().orc_faithful_eval(C.byref(hcs), 100_000, T0)
            reps_b1 += 1
        v_b1 = 100_000 * reps_b1 / (time.perf_counter() - t0)
        cpu = {"value": v_mt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{k_mt} ticks x {n} records (whole config-2 population), oracle sweep on {threads} threads",
               "single_thread_value": v_1t,
               "faithful_shape_value": v_b1,
               "faithful_shape_note": "1 thread, re-parse cron + Next() per evaluation as hcc.go:253-262 does, 100k-record subsample"}

    if rank == 0:
        value = n * world * args.steps / (ms * 1e-3)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": w, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[1]: {n} HealthChecks per GPU, mixed 5-field cron + "
                                   "repeatAfterSec (tools/amgen config 2, seed 2); step = the single tick T0=2026-09-21T09:15:00Z",
                       "records_per_gpu": n, "records_total": n * world,
                       "l2": "inputs (560 MB/GPU) larger than L2 (126 MB); no flush needed",
                       "parallelism": f"index-range shards x{world}" + (
                           "" if world == 1 else (f", due lists concatenated by the NVLink peer-write kernel, wire format {args.wire} (overlapped with the next sweep)"
                                                  if peer is not None else ", padded NCCL all-gather of due lists")),
                       "due_per_tick": stats["n_submit_hc"], "emitted_per_tick": stats["n_emitted"]},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC_BYTES_10M_CONFIG2 if n == N_PER_GPU else None,
                         "traffic_source": "profiles/r01_sweep_config2_ncu_full.csv (ncu --set full, one launch)",
                         "peak_source": peak_src,
                         "kernel": "sweep_tick_kernel<false>", "kernel_ms": k_ms,
                         "algorithmic_bytes": alg_bytes,
                         "bytes_model": f"N*{B_READ} + emitted*{B_EMIT} + stopped*{B_STOP}",
                         "second_kernel": {"name": "compact_kernel", "kernel_ms": c_ms,
                                           "bytes": stats["n_emitted"] * B_EMIT * 2},
                         "kernel_share_of_step": k_ms / (ms / args.steps)},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        if peer is not None:
            peer.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
