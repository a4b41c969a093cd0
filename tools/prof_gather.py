"""Time the NVLink peer-write exchange alone (run under torchrun, >= 2 GPUs).
Each rank sweeps its 10 M-record shard once, then pushes the same due list K
times back to back on one stream; CUDA events around the pushes."""
import importlib
import os
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
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
am = importlib.import_module("active-monitor_b200")
gather = importlib.import_module("active-monitor_b200.gather")
T0 = amgen.T0_MON_0915
cols = amgen.fill(2, 2, rank * n, n, T0, am.load().am_healthcheck_classify)
s = am.Sweep(capacity=n, device=lr, shard_base=rank * n)
s.load_range(0, cols)
d_idx = torch.empty(n, dtype=torch.int32, device=dev)
d_act = torch.empty(n, dtype=torch.uint8, device=dev)
d_cnt = torch.zeros(1, dtype=torch.int32, device=dev)
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
s.tick_device(T0, 0, d_idx.data_ptr(), d_act.data_ptr(), n, d_cnt.data_ptr(), 0, st.cuda_stream)
st.synchronize()
CASES = [(4, "plain"), (4, "c3"), (8, "plain")] + ([(4, "bm")] if os.environ.get("AMSWEEP_TEST_EXPERIMENTAL_WIRES") else [])
for ib, wire in CASES:
    pg = gather.PeerGather(lr, cap_total=n * world, idx_bytes=ib, shard=(rank * n, n) if wire in ("c3", "bm") else None,
                           wire="bm" if wire == "bm" else "c3")
    for _ in range(5):
        pg.push(d_idx.data_ptr(), d_act.data_ptr(), d_cnt.data_ptr(), rank * n, st.cuda_stream)
    st.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(K):
        pg.push(d_idx.data_ptr(), d_act.data_ptr(), d_cnt.data_ptr(), rank * n, st.cuda_stream)
    e1.record(st)
    st.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / K
    cnt = int(d_cnt.item())
    wire_b = {"c3": 3, "bm": 0.4}.get(wire, ib + 1)  # bytes per entry on the wire (bm: bitmap + non-default actions)
    print(f"rank {rank} idx_bytes {ib} wire {wire}: {us:.1f} us/exchange, {cnt} entries, "
          f"{cnt * wire_b * (world - 1) / us / 1e3:.1f} GB/s out over NVLink", flush=True)
    dist.barrier()
    pg.close()
dist.barrier()
dist.destroy_process_group()
