"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): every rank sweeps its index-range shard
on its own B200, then the global due list is assembled on every GPU three ways — the tick
exchange (am_sweep_tick_shard + am_gather_exchange: bitmap + exceptions over NVLink, list
rebuilt locally), the round-1 plain list push, and the NCCL padded all-gather (baseline) — and
all three must equal the UNSHARDED oracle run, every tick, every rank; the shard statistics the
exchange publishes must equal the oracle's for that shard.

Run on hardware with the final SASS: see profiles/r02_multigpu_parity.txt."""
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


def _worker(rank, world, port, n_total, ticks, config, overlap, q):
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
        cols = amgen.fill(config, 4, first, cnt, T0, am.load().am_healthcheck_classify, threads=8)
        out = []
        # two sweeps over the same shard: one feeds the exchange, one the list-based paths
        with am.Sweep(capacity=cnt, device=rank, shard_base=first) as s, \
                am.Sweep(capacity=cnt, device=rank, shard_base=first) as s2:
            s.load_range(0, cols)
            s2.load_range(0, cols)
            # u32 global indices on worlds divisible by 4, u64 otherwise
            ib = 4 if world % 4 == 0 else 8
            pg = gather.PeerGather(rank, cap_total=n_total, idx_bytes=ib, shard=(first, cnt))
            pg2 = gather.PeerGather(rank, cap_total=n_total, idx_bytes=ib)
            d_idx = torch.empty(cnt, dtype=torch.int32, device=dev)
            d_act = torch.empty(cnt, dtype=torch.uint8, device=dev)
            d_cnt = torch.zeros(1, dtype=torch.int32, device=dev)
            d_st = torch.zeros(16, dtype=torch.int64, device=dev)
            sa = torch.cuda.Stream(device=dev)
            sb = torch.cuda.Stream(device=dev) if overlap else sa
            torch.cuda.set_stream(sa)
            for k in range(ticks):
                T = T0 + 60 * k - (k % 2)
                own = None
                if overlap and k % 2 == 1:
                    # the whole step in one call: tick_shard + exchange + this rank's own part of the list as local
                    # slots in pinned host memory (am_gather_bind / am_gather_tick_view)
                    if k == 1:
                        pg.bind(s, sa.cuda_stream, sb.cuda_stream)
                    vi, va, vst = pg.tick_view(T)
                    own = (vi.astype(np.uint64) + np.uint64(first), va.astype(np.uint32))
                    st = [vst[f] for f in am.abi.STAT_FIELDS]
                else:
                    s.tick_shard(T, 0, sa.cuda_stream)
                    pg.exchange(s, d_st.data_ptr(), sb.cuda_stream)  # the library orders sb after sa
                    sb.synchronize()
                    st = [int(v) & ((1 << 64) - 1) for v in d_st.cpu().tolist()]
                xi, xa, xc = pg.result()
                xi, xa = xi.cpu().numpy().copy(), xa.cpu().numpy().copy()
                s2.tick_device(T, 0, d_idx.data_ptr(), d_act.data_ptr(), cnt, d_cnt.data_ptr(), 0, sa.cuda_stream)
                pg2.push(d_idx.data_ptr(), d_act.data_ptr(), d_cnt.data_ptr(), first, sa.cuda_stream)
                sa.synchronize()
                pi, pa, pc = pg2.result()
                ni, na, nc = gather.allgather_due(d_idx, d_act, d_cnt, first)
                out.append((xi, xa, xc, st, pi.cpu().numpy().copy(), pa.cpu().numpy().copy(), pc,
                            ni.cpu().numpy().copy(), na.cpu().numpy().copy(), nc, own))
            final = s.read_range(0, cnt)
            dist.barrier()
            pg.close()
            pg2.close()
        q.put((rank, out, {k: v for k, v in final.items()}))
    finally:
        dist.destroy_process_group()


def _u(a):
    return (a.astype(np.int64) & (0xFFFFFFFF if a.dtype == np.int32 else -1)).astype(np.uint64)  # u32 viewed as i32


@pytest.mark.parametrize("overlap", [False, True], ids=["one_stream", "two_streams"])
@pytest.mark.parametrize("config", [3, 2])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_sweep_and_all_gathers_equal_unsharded_oracle(world, config, overlap):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    import amgen
    import oracle_c
    gather = importlib.import_module("active-monitor_b200.gather")
    am = importlib.import_module("active-monitor_b200")
    n_total, ticks = 400_003, 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 1000 + 3 * config + (1 if overlap else 0)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, ticks, config, overlap, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(world)]
    results = {r: o for r, o, _ in got}
    finals = {r: f for r, _, f in got}
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    whole = amgen.fill(config, 4, 0, n_total, T0, oracle_c.load().orc_classify)
    shard_cols = {}
    for rank in range(world):
        first, cnt = gather.shard_range(n_total, rank, world)
        shard_cols[rank] = (first, {k: v[first:first + cnt].copy() for k, v in whole.items()})
    for k in range(ticks):
        T = T0 + 60 * k - (k % 2)
        wi, wa, _ = oracle_c.sweep(whole, T)
        for rank in range(world):
            xi, xa, xc, st, pi, pa, pc, ni, na, nc, own = results[rank][k]
            first, sc = shard_cols[rank]
            if own is not None:
                mine = (wi >= first) & (wi < first + len(sc["flags"]))
                np.testing.assert_array_equal(own[0], wi[mine], err_msg=f"tick_view rank {rank} tick {k}")
                np.testing.assert_array_equal(own[1], wa[mine])
            _, _, ws = oracle_c.sweep(sc, T, shard_base=first)
            assert dict(zip(am.abi.STAT_FIELDS, st)) == ws, f"shard stats rank {rank} tick {k}"
            assert sum(xc) == len(wi) and xc == pc == nc
            np.testing.assert_array_equal(_u(xi), wi, err_msg=f"tick exchange rank {rank} tick {k}")
            np.testing.assert_array_equal(xa.astype(np.uint32), wa)
            np.testing.assert_array_equal(_u(pi), wi, err_msg=f"plain push rank {rank} tick {k}")
            np.testing.assert_array_equal(pa.astype(np.uint32), wa)
            np.testing.assert_array_equal(ni.astype(np.uint64), wi, err_msg=f"nccl gather rank {rank} tick {k}")
            np.testing.assert_array_equal(na.astype(np.uint32), wa)
    for rank in range(world):
        for name in am.COLUMN_NAMES:
            np.testing.assert_array_equal(finals[rank][name], shard_cols[rank][1][name], err_msg=f"rank {rank} column {name}")
