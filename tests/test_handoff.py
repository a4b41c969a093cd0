"""Hand-off queue: host-only part of the C-ABI, runs without a GPU.

What the reference guarantees at this seam and the queue must keep: every due
HealthCheck is handed to exactly one worker (one AfterFunc fire per timer,
hcc.go:751), in the order the sweep emitted it; bounded concurrency
(MaxConcurrentReconciles = MaxParallel, hcc.go:298) becomes bounded depth.
"""
import importlib
import threading

import numpy as np
import pytest

am = importlib.import_module("active-monitor_b200")
handoff = importlib.import_module("active-monitor_b200.handoff")

SUBMIT, REMEDY, STOPPED = am.ACT_SUBMIT_HC, am.ACT_RUN_REMEDY, am.ACT_STOPPED


def test_fifo_order_and_mask_filter():
    with handoff.Handoff(16) as q:
        idx = np.array([3, 5, 8, 13, 21], dtype=np.uint64)
        act = np.array([SUBMIT, STOPPED, SUBMIT | am.ACT_RESET_ON_PASS, REMEDY | am.ACT_RESET_ON_INTERVAL, 0x08],
                       dtype=np.uint32)
        assert q.publish(100, idx, act) == 3  # STOPPED and PARSE_ERROR are the ticker's own business
        got = q.pop(10)
        assert got["idx"].tolist() == [3, 8, 13]
        assert got["action"].tolist() == [SUBMIT, SUBMIT, REMEDY]  # only the masked bits travel
        assert set(got["unix_sec"].tolist()) == {100}
        assert q.pop(10).size == 0
        assert q.stats() == {"pending": 0, "published": 3, "popped": 3, "rejected_batches": 0}


def test_all_or_nothing_when_full_and_wraparound():
    with handoff.Handoff(8) as q:
        a = np.full(6, SUBMIT, dtype=np.uint32)
        assert q.publish(1, np.arange(6), a) == 6
        with pytest.raises(am.AmError) as e:
            q.publish(2, np.arange(100, 103), a[:3])  # 3 do not fit in the 2 free slots
        assert e.value.code == am.AM_E_NOSPACE
        assert q.stats()["pending"] == 6 and q.stats()["rejected_batches"] == 1
        assert q.pop(4)["idx"].tolist() == [0, 1, 2, 3]
        assert q.publish(2, np.arange(100, 106), a) == 6  # wraps around the ring
        got = q.pop(100)
        assert got["idx"].tolist() == [4, 5, 100, 101, 102, 103, 104, 105]
        assert got["unix_sec"].tolist() == [1, 1, 2, 2, 2, 2, 2, 2]


def test_empty_publish_and_bad_arguments():
    lib = am.load()
    with handoff.Handoff(4) as q:
        assert q.publish(5, np.empty(0, np.uint64), np.empty(0, np.uint32)) == 0
        assert q.publish(5, np.arange(3), np.zeros(3, np.uint32)) == 0  # nothing matches the mask
    assert lib.am_handoff_create(None, 4) == am.AM_E_INVAL
    import ctypes as C
    h = C.c_void_p()
    assert lib.am_handoff_create(C.byref(h), 0) == am.AM_E_INVAL
    assert lib.am_handoff_pop(None, 1, None, None) == am.AM_E_INVAL
    lib.am_handoff_destroy(None)  # no-op


def test_every_item_reaches_exactly_one_worker_in_order():
    """One publisher (the ticker), eight poppers (MaxParallel workers)."""
    n_ticks, per_tick, workers = 200, 500, 8
    q = handoff.Handoff(4 * per_tick)
    seen = [[] for _ in range(workers)]
    done = threading.Event()

    def worker(k):
        while True:
            got = q.pop(64)
            if got.size:
                seen[k].append(got.copy())
            elif done.is_set() and q.stats()["pending"] == 0:
                return

    th = [threading.Thread(target=worker, args=(k,)) for k in range(workers)]
    for t in th:
        t.start()
    act = np.full(per_tick, SUBMIT, dtype=np.uint32)
    for tick in range(n_ticks):
        idx = np.arange(tick * per_tick, (tick + 1) * per_tick, dtype=np.uint64)
        while True:  # back-pressure: retry the whole tick, as the Go ticker would
            try:
                q.publish(1000 + tick, idx, act)
                break
            except am.AmError as e:
                assert e.code == am.AM_E_NOSPACE
    done.set()
    for t in th:
        t.join(timeout=60)
        assert not t.is_alive()
    per_worker = [np.concatenate(s) if s else np.empty(0, dtype=am.WORK_ITEM_DTYPE) for s in seen]
    for w in per_worker:  # each worker sees a subsequence of the global order
        assert np.all(np.diff(w["idx"].astype(np.int64)) > 0)
    allv = np.sort(np.concatenate(per_worker)["idx"])
    assert np.array_equal(allv, np.arange(n_ticks * per_tick, dtype=np.uint64))  # exactly once
    items = np.concatenate(per_worker)
    assert np.array_equal(items["unix_sec"], 1000 + (items["idx"] // per_tick).astype(np.int64))
    st = q.stats()
    assert st["published"] == st["popped"] == n_ticks * per_tick and st["pending"] == 0
    q.close()
