"""Time the NVLink tick exchange alone (run under torchrun, >= 2 GPUs) and count the NVLink bytes.
Each rank sweeps its shard once (am_sweep_tick_shard), then exchanges the SAME tick K times back to
back on one stream; CUDA events around the exchanges; `nvidia-smi nvlink -gt d` data counters before
and after (rank 0's GPU) give the bytes that really crossed the links.

env: N (records per rank, 10 M), K (exchanges, 50), CONFIG (2)."""
import importlib
import json
import os
import re
import subprocess
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools", "amgen")):
    sys.path.insert(0, p)
import amgen  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
n = int(os.environ.get("N", 10_000_000))
K = int(os.environ.get("K", 50))
config = int(os.environ.get("CONFIG", 2))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
am = importlib.import_module("active-monitor_b200")
gather = importlib.import_module("active-monitor_b200.gather")
T0 = amgen.T0_MON_0915


def nvlink_kib(gpu):
    """sum of the per-link data counters of one GPU: (tx KiB, rx KiB), or None"""
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(gpu)], capture_output=True, text=True, timeout=20).stdout
        tx = sum(int(v) for v in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out))
        rx = sum(int(v) for v in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out))
        return (tx, rx) if (tx or rx) else None
    except Exception:
        return None


cols = amgen.fill(config, config, rank * n, n, T0, am.load().am_healthcheck_classify, threads=8)
s = am.Sweep(capacity=n, device=lr, shard_base=rank * n)
s.load_range(0, cols)
d_st = torch.zeros(16, dtype=torch.int64, device=dev)
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
pg = gather.PeerGather(lr, cap_total=n * world, idx_bytes=4 if n * world < (1 << 32) else 8, shard=(rank * n, n))
s.tick_shard(T0, 0, st.cuda_stream)   # "Stopped" reports are one-shot: tick twice, exchange the second
s.tick_shard(T0, 0, st.cuda_stream)
for _ in range(5):
    pg.exchange(s, d_st.data_ptr(), st.cuda_stream)
st.synchronize()
dist.barrier()
c0 = nvlink_kib(lr) if rank == 0 else None
dist.barrier()  # the counter read takes a while on rank 0: do not let the peers run ahead into the timed loop
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(K):
    pg.exchange(s, d_st.data_ptr(), st.cuda_stream)
e1.record(st)
st.synchronize()
dist.barrier()
c1 = nvlink_kib(lr) if rank == 0 else None
us = e0.elapsed_time(e1) * 1e3 / K
pg.set_profiling(True)
prof = []
for _ in range(10):
    dist.barrier()
    pg.exchange(s, d_st.data_ptr(), st.cuda_stream)
    prof.append(pg.last_profile())
pg.set_profiling(False)
prof_us = [round(sum(p[k] for p in prof[2:]) / len(prof[2:]) * 1e3, 1) for k in range(3)]
stats = dict(zip(am.abi.STAT_FIELDS, d_st.cpu().tolist()))
idx, act, counts = pg.result()
n_exc = int((act[sum(counts[:rank]): sum(counts[: rank + 1])] > 1).sum().item())
payload = (((n + 8191) // 8192) * 1024 + ((n + 8191) // 8192 + 1) * 4 + ((n + 1023) // 1024) * 4 + n_exc * 4)
t = torch.tensor([us], dtype=torch.float64, device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    out = {"world": world, "records_per_rank": n, "config": config, "exchanges": K,
           "us_per_exchange_max_over_ranks": float(t.item()),
           "what": "push (bitmap + prefixes + exceptions to every peer) + counts + list rebuild of the global list + publish",
           "rank0_us_push_rebuild_publish_after_a_barrier": prof_us,
           "emitted_per_rank": int(counts[0]), "global_entries": int(sum(counts)), "exceptions_rank0": n_exc,
           "payload_bytes_per_peer": payload, "payload_bytes_out_per_exchange": payload * (world - 1),
           "rebuilt_list_bytes_per_gpu": int(sum(counts)) * (pg.idx_bytes + 1)}
    if c0 and c1:
        tx, rx = (c1[0] - c0[0]) * 1024 / K, (c1[1] - c0[1]) * 1024 / K
        out["nvlink_counters_rank0"] = {"tx_bytes_per_exchange": tx, "rx_bytes_per_exchange": rx,
                                        "tx_over_payload": tx / max(1, payload * (world - 1)),
                                        "tx_GBps_while_exchanging": tx / (float(t.item()) * 1e-6) / 1e9}
    else:
        out["nvlink_counters_rank0"] = None
    print(json.dumps(out), flush=True)
dist.barrier()
pg.close()
dist.destroy_process_group()
