"""am_healthcheck_classify_batch: the threaded bulk form of the ladder must be
the per-record classifier, record for record, whatever the thread count.  Host only."""
import importlib
import time

import numpy as np
import pytest

am = importlib.import_module("active-monitor_b200")
ingest = importlib.import_module("active-monitor_b200.ingest")

CRONS = ["", "@every 5s", "@every 1m", "@every 1h30m", "*/5 * * * *", "0 9 * * mon-fri", "15 10 1,15 * ?",
         "0 0 1 jan *", "@hourly", "@daily", "NOT_A_VALID_CRON", "60 * * * *", "* * * * 7", "* * * * * *",
         "CRON_TZ=UTC 0 12 * * *", "CRON_TZ=Asia/Tokyo 0 12 * * *", "@every 500ms", "@every -5s", "1-5/2 * * * *"]


def _population(n, seed):
    rng = np.random.default_rng(seed)
    items = []
    for i in range(n):
        k = int(rng.integers(0, 10))
        it = dict(
            repeat_after_sec=int(rng.choice([0, 0, 0, -1, 5, 60, 3600, 1 << 40])) if k < 8 else 0,
            cron=CRONS[int(rng.integers(0, len(CRONS)))],
            has_resource=bool(rng.integers(0, 20)),
            has_remedy=bool(rng.integers(0, 2)),
            remedy_runs_limit=int(rng.choice([0, 1, 2, 5, -1, 1 << 35])),
            remedy_reset_interval=int(rng.choice([0, 60, 300])),
            finished_at=None if k == 0 else 1789982100 - int(rng.integers(0, 7200)),
            remedy_finished_at=None if k < 5 else 1789982100 - int(rng.integers(1, 600)),
            success_count=int(rng.integers(0, 1000)), failed_count=int(rng.integers(0, 1000)),
            remedy_success_count=int(rng.integers(0, 4)), remedy_failed_count=int(rng.integers(0, 4)),
        )
        it["remedy_total_runs"] = it["remedy_success_count"] + it["remedy_failed_count"]
        items.append(it)
    return items


@pytest.mark.parametrize("threads", [1, 2, 7, 0])
def test_batch_equals_single_calls(threads):
    items = _population(20_000, seed=11)
    rcs, recs = ingest.classify_batch(items, n_threads=threads)
    # every error class of the single call shows up in the population
    assert {am.AM_OK, am.AM_E_RANGE} <= set(int(x) for x in np.unique(rcs))
    for i in range(0, len(items), 7):  # a 1/7 sample through the one-record entry point
        rc, rec = am.classify(**items[i])
        assert rc == int(rcs[i]), (i, items[i])
        if rc != am.AM_E_RANGE:  # the single call leaves a rejected record zeroed or partial: only rc is contractual
            assert rec.tobytes() == recs[i:i + 1].tobytes(), (i, items[i])
    # thread count must not change a single byte
    rcs1, recs1 = ingest.classify_batch(items, n_threads=1)
    assert np.array_equal(rcs, rcs1) and recs.tobytes() == recs1.tobytes()


def test_batch_edge_cases():
    import ctypes as C
    lib = am.load()
    bad = C.c_uint64(7)
    assert lib.am_healthcheck_classify_batch(None, 0, None, None, 4, C.byref(bad)) == am.AM_OK and bad.value == 0
    assert lib.am_healthcheck_classify_batch(None, 5, None, None, 4, None) == am.AM_E_INVAL
    rcs, recs = ingest.classify_batch([dict(cron="@every 5s")], n_threads=64)  # more threads than records
    assert int(rcs[0]) == 0 and int(recs["ras"][0]) == 5 and int(recs["flags"][0]) & 7 == am.KIND_CRON_EVERY
    with pytest.raises(TypeError):
        ingest.classify_batch([dict(cron_expr="* * * * *")])


def test_batch_throughput_is_reported(capsys):
    """Not a performance gate: prints the host-side ingest rate for the record."""
    items = _population(5_000, seed=3) * 40  # 200 k records, the per-item Python cost is paid once below
    n = len(items)
    import ctypes as C
    arr = (am.AmHealthCheck * n)()
    keep = []
    for i, kw in enumerate(items):
        raw = kw["cron"].encode()
        keep.append(raw)
        arr[i].cron, arr[i].cron_len = raw, len(raw)
        arr[i].repeat_after_sec = 0
        arr[i].has_resource = 1
    recs = np.zeros(n, dtype=am.RECORD_DTYPE)
    lib = am.load()
    out = {}
    for nt in (1, 0):
        t0 = time.perf_counter()
        assert lib.am_healthcheck_classify_batch(C.cast(arr, C.c_void_p), n, recs.ctypes.data, None, nt, None) == 0
        out[nt] = n / (time.perf_counter() - t0)
    with capsys.disabled():
        print(f"\n[ingest] am_healthcheck_classify_batch: {out[1]/1e6:.2f} M records/s on 1 thread, "
              f"{out[0]/1e6:.2f} M records/s on all host threads")
    assert out[1] > 1e5
