"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): every rank sweeps its
index-range shard on its own B200, then the global due list is assembled on
every GPU twice — NCCL padded all-gather (baseline) and the NVLink peer-write
kernel (csrc/gather.cu) — and both must equal the UNSHARDED oracle run."""
import importlib
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
T0 = 1789982100


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, port, n_total, ticks, wire, q):
    import torch
    import torch.distributed as dist
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools", "amgen")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import amgen
        am = importlib.import_module("active-monitor_b200")
        gather = importlib.import_module("active-monitor_b200.gather")
        first, cnt = gather.shard_range(n_total, rank, world)
        cols = amgen.fill(3, 4, first, cnt, T0, am.load().am_healthcheck_classify, threads=8)
        out = []
        with am.Sweep(capacity=cnt, device=rank, shard_base=first) as s:
            s.load_range(0, cols)
            # u32 global indices on worlds divisible by 4, u64 otherwise; wire = plain | c3 (compressed)
            pg = gather.PeerGather(rank, cap_total=n_total, idx_bytes=4 if world % 4 == 0 else 8,
                                   shard=(first, cnt) if wire in ("c3", "bm") else None,
                                   wire="bm" if wire == "bm" else "c3")
            d_idx = torch.empty(cnt, dtype=torch.int32, device=dev)
            d_act = torch.empty(cnt, dtype=torch.uint8, device=dev)
            d_cnt = torch.zeros(1, dtype=torch.int32, device=dev)
            stream = torch.cuda.Stream(device=dev)
            torch.cuda.set_stream(stream)
            for k in range(ticks):
                T = T0 + 60 * k
                s.tick_device(T, 0, d_idx.data_ptr(), d_act.data_ptr(), cnt, d_cnt.data_ptr(), 0,
                              stream.cuda_stream)
                pg.push(d_idx.data_ptr(), d_act.data_ptr(), d_cnt.data_ptr(), first, stream.cuda_stream)
                stream.synchronize()
                pi, pa, pc = pg.result()
                ni, na, nc = gather.allgather_due(d_idx, d_act, d_cnt, first)
                out.append((pi.cpu().numpy().copy(), pa.cpu().numpy().copy(), pc,
                            ni.cpu().numpy().copy(), na.cpu().numpy().copy(), nc))
            dist.barrier()
            pg.close()
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


# "bm" (experimental bitmap wire format, not yet run on hardware) only on request
WIRES = ["plain", "c3"] + (["bm"] if os.environ.get("AMSWEEP_TEST_EXPERIMENTAL_WIRES") else [])


@pytest.mark.parametrize("wire", WIRES)
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_sweep_and_both_gathers_equal_unsharded_oracle(world, wire):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    import amgen
    import oracle_c
    n_total, ticks = 400_003, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 1000
    port += {"plain": 0, "c3": 7, "bm": 14}[wire]
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, ticks, wire, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    whole = amgen.fill(3, 4, 0, n_total, T0, oracle_c.load().orc_classify)
    for k in range(ticks):
        wi, wa, _ = oracle_c.sweep(whole, T0 + 60 * k)
        for rank in range(world):
            pi, pa, pc, ni, na, nc = results[rank][k]
            assert sum(pc) == len(wi) and pc == nc
            pi = pi.astype(np.int64) & (0xFFFFFFFF if pi.dtype == np.int32 else -1)  # u32 viewed as i32
            np.testing.assert_array_equal(pi.astype(np.uint64), wi, err_msg=f"peer gather rank {rank} tick {k}")
            np.testing.assert_array_equal(pa.astype(np.uint32), wa)
            np.testing.assert_array_equal(ni.astype(np.uint64), wi, err_msg=f"nccl gather rank {rank} tick {k}")
            np.testing.assert_array_equal(na.astype(np.uint32), wa)
