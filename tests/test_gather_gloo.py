"""N>1 host logic on CPU: gloo ranks shard a population by index range, each
produces its local due list (here from the CPU oracle: this test is about the
sharding/concatenation plumbing, not the kernel) and allgather_due must hand
every rank the global ascending list of the unsharded run."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

T0 = 1789982100


def _worker(rank, world, port, n_total, q):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools", "amgen")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import amgen
        import oracle_c
        gather = importlib.import_module("active-monitor_b200.gather")
        first, cnt = gather.shard_range(n_total, rank, world)
        cols = amgen.fill(2, 4, first, cnt, T0, oracle_c.load().orc_classify, threads=2)
        gidx, act, _ = oracle_c.sweep(cols, T0, shard_base=first)
        local = torch.from_numpy((gidx - np.uint64(first)).astype(np.int64)).to(torch.int32)
        # over-allocated buffers, as on the device: only `count` entries are valid
        pad = torch.full((len(local) + 7,), -1, dtype=torch.int32)
        pad[: len(local)] = local
        a = torch.zeros(len(local) + 7, dtype=torch.uint8)
        a[: len(local)] = torch.from_numpy(act.astype(np.uint8))
        idx_all, act_all, counts = gather.allgather_due(pad, a, len(local), first)
        q.put((rank, idx_all.numpy().copy(), act_all.numpy().copy(), counts))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 20_001), (3, 7_000)])
def test_allgather_due_equals_unsharded(world, n_total):
    import amgen
    import oracle_c
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole = amgen.fill(2, 4, 0, n_total, T0, oracle_c.load().orc_classify, threads=2)
    widx, wact, _ = oracle_c.sweep(whole, T0)
    for rank, idx_all, act_all, counts in results:
        np.testing.assert_array_equal(idx_all.astype(np.uint64), widx)
        np.testing.assert_array_equal(act_all.astype(np.uint32), wact)
        assert sum(counts) == len(widx) and len(counts) == world


def test_shard_range_partitions_exactly():
    gather = importlib.import_module("active-monitor_b200.gather")
    for n, w in [(100_000_000, 8), (10, 3), (7, 8), (0, 4), (1, 1)]:
        spans = [gather.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == n
        for (f0, c0), (f1, _) in zip(spans, spans[1:]):
            assert f0 + c0 == f1
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
