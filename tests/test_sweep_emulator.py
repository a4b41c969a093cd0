"""The whole product library (csrc/sweep.cu host runtime + csrc/sweep_kernels.cuh + csrc/sweep_block.cuh) on the
CPU emulation of the CUDA execution model (tests/emu), in-process through `EmuSweep` = `am.Sweep` bound to the
emulated build: the SAME kernel and host source the GPU runs — sweep_tick_kernel in its four variants,
scan_groups / expand / publish, the staged-op kernels, sweep_block_kernel, next_fire / next_due — against the
oracle, bit-exact: emitted lists, action bytes, statistics, every mutated column.

This is the CPU-side twin of tests/test_sweep_gpu.py at oracle-friendly sizes.  It checks logic (decisions, the
ordered list rebuild, statistics, staging and drain, the parity hand-off between ticks); it says nothing about
the GPU memory model or performance."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import emu_sweep  # noqa: E402

T0 = 1789982100      # 2026-09-21 09:15:00 UTC Monday
T_OCT1 = 1790812800  # 2026-10-01 00:00:00 UTC Thursday


def _gen_pair(gen, am, orc, config, seed, n, T, first=0):
    prod = gen.fill(config, seed, first, n, T, am.load().am_healthcheck_classify)
    orac = gen.fill(config, seed, first, n, T, orc.load().orc_classify)
    return prod, orac


def _assert_tick_equal(am, got, want, sweep, ocols, n, what=""):
    gi, ga, gs = got
    wi, wa, ws = want
    assert gs == ws, f"{what}: stats differ\n emu={gs}\n cpu={ws}"
    np.testing.assert_array_equal(gi, wi, err_msg=f"{what}: due indices")
    np.testing.assert_array_equal(ga, wa, err_msg=f"{what}: action bytes")
    dev = sweep.read_range(0, n)
    for name in am.COLUMN_NAMES:
        np.testing.assert_array_equal(dev[name], ocols[name], err_msg=f"{what}: column {name}")


@pytest.mark.parametrize("config", [1, 11])
def test_config1(am, orc, gen, config):
    n = 1000
    prod, orac = _gen_pair(gen, am, orc, config, 1, n, T0)
    with emu_sweep.EmuSweep(n) as s:
        s.load_range(0, prod)
        _assert_tick_equal(am, s.tick(T0), orc.sweep(orac, T0), s, orac, n, f"config {config}")


@pytest.mark.parametrize("T", [T0, T_OCT1, T0 + 1, T0 + 45, 1709164800])
@pytest.mark.parametrize("n", [1, 63, 1025, 20_003])
def test_config2_mixed(am, orc, gen, T, n):
    prod, orac = _gen_pair(gen, am, orc, 2, 2, n, T0)
    with emu_sweep.EmuSweep(n) as s:
        s.load_range(0, prod)
        _assert_tick_equal(am, s.tick(T), orc.sweep(orac, T), s, orac, n, f"config2 n={n} T={T}")
        # one second later: "Stopped" is reported once only; other parity of the group counters
        _assert_tick_equal(am, s.tick(T + 1), orc.sweep(orac, T + 1), s, orac, n, "second tick")


@pytest.mark.parametrize("T", [T0, T_OCT1, T0 + 1])
def test_config3_remedy_state_machine_dense_and_sparse(am, orc, gen, T):
    n = 20_000
    prod, orac = _gen_pair(gen, am, orc, 3, 3, n, T0)
    with emu_sweep.EmuSweep(n) as s:
        s.load_range(0, prod)
        got = s.tick(T)
        _assert_tick_equal(am, got, orc.sweep(orac, T), s, orac, n, "config3")
        if T == T0:
            for k in ("n_run_remedy", "n_remedy_skip", "n_reset_on_pass", "n_reset_on_interval",
                      "n_result_ok", "n_result_fail", "n_remedy_ok", "n_remedy_fail"):
                assert got[2][k] > 0, f"population does not exercise {k}"
        # nothing pending any more: the sparse path (few or no needy lanes per warp)
        _assert_tick_equal(am, s.tick(T + 60), orc.sweep(orac, T + 60), s, orac, n, "config3 next minute")


def test_full_scan_mode_and_shard_base(am, orc, gen):
    n, base = 9000, 12_500_000
    prod, orac = _gen_pair(gen, am, orc, 2, 7, n, T0, first=base)
    with emu_sweep.EmuSweep(n, shard_base=base) as s:
        s.load_range(0, prod)
        for T in (T0 + 7, T0 + 60):
            got = s.tick(T, mode=am.SWEEP_FULL_SCAN)
            want = orc.sweep(orac, T, shard_base=base)
            _assert_tick_equal(am, got, want, s, orac, n, f"full-scan T={T}")
            assert got[0].min() >= base


@pytest.mark.parametrize("base", [(1 << 32) - 5000, (1 << 32) - 8192 - 3, (3 << 32) - 20_001, 8192 * 1000 + 4097])
def test_index_checksums_where_a_group_straddles_2_to_the_32(am, orc, gen, base):
    """expand_kernel checksums a group's indices on their low words (the high word is shared) unless the
    group straddles a multiple of 2^32; statistics (idx_xor, idx_sum), list and short-buffer behaviour
    must equal the oracle's on both sides of that rule."""
    n = 20_000
    prod, orac = _gen_pair(gen, am, orc, 2, 7, n, T0, first=base)
    with emu_sweep.EmuSweep(n, shard_base=base) as s:
        s.load_range(0, prod)
        got = s.tick(T0)
        want = orc.sweep(orac, T0, shard_base=base)
        _assert_tick_equal(am, got, want, s, orac, n, f"base={base}")
        assert len(got[0]) > 5000
        # a short caller buffer: the statistics still describe the whole tick, the list is its prefix
        gi, ga, gs = s.tick(T0 + 60, cap=1001)
        wi, wa, ws = orc.sweep(orac, T0 + 60, shard_base=base)
        assert gs == ws
        np.testing.assert_array_equal(gi, wi[:1001])
        np.testing.assert_array_equal(ga, wa[:1001])


@pytest.mark.parametrize("config", [5, 55])
def test_closed_loop_many_ticks(am, orc, gen, config):
    n, seed = 6000, 5
    prod, orac = _gen_pair(gen, am, orc, config, seed, n, T0)
    with emu_sweep.EmuSweep(n) as s:
        s.load_range(0, prod)
        s.set_seed(seed)
        for k in range(75):
            T = T0 - 10 + k  # crosses a minute boundary: both MASKS variants of the closed-loop kernel
            got = s.tick(T, mode=am.SWEEP_CLOSED_LOOP)
            want = orc.sweep(orac, T, mode=1, seed=seed)
            assert got[2] == want[2], f"config {config} tick {k}"
            np.testing.assert_array_equal(got[0], want[0])
            np.testing.assert_array_equal(got[1], want[1])
        dev = s.read_range(0, n)
        for name in am.COLUMN_NAMES:
            np.testing.assert_array_equal(dev[name], orac[name], err_msg=f"config {config} {name}")


def test_short_output_buffer(am, gen):
    n = 4096
    prod = gen.fill(1, 1, 0, n, T0, am.load().am_healthcheck_classify)
    with emu_sweep.EmuSweep(n) as s:
        s.load_range(0, prod)
        part = s.tick(T0, cap=10)
        full = s.tick(T0)
        assert part[2]["n_emitted"] == full[2]["n_emitted"] == len(full[0]) > 10
        np.testing.assert_array_equal(part[0], full[0][:10])


def test_staged_events_take_effect_in_call_order(am):
    T = T0

    def rec(ras, fin):
        rc, r = am.classify(repeat_after_sec=ras, finished_at=fin)
        assert rc == 0
        return r

    A, B = rec(60, T - 1000), rec(7, T - 1)  # A is due at T, B is not
    with emu_sweep.EmuSweep(64) as s:
        s.upsert([0], A); s.remove([0])
        s.remove([1]); s.upsert([1], A)
        s.upsert([2], A); s.upsert([2], B)
        s.upsert([3], B); s.upsert([3], A)
        s.upsert([4], A)
        s.post_result([4], [am.PHASE_FAILED]); s.upsert([4], B)
        s.upsert([5], B); s.post_result([5], [am.PHASE_SUCCEEDED])
        s.upsert([6], B)
        s.post_result([6], [am.PHASE_SUCCEEDED]); s.post_result([6], [am.PHASE_FAILED])
        s.upsert(np.array([7, 7, 7]), np.concatenate([A, B, A]))
        idx, act, st = s.tick(T)
        got = dict(zip(idx.tolist(), act.tolist()))
        assert got == {1: am.ACT_SUBMIT_HC, 3: am.ACT_SUBMIT_HC, 7: am.ACT_SUBMIT_HC}, got
        cols = s.read_range(0, 8)
        assert cols["flags"][0] == am.F_TOMBSTONE
        assert cols["ras"].tolist()[1:8] == [60, 7, 60, 7, 7, 7, 60]
        assert cols["failed"][4] == 0 and cols["finished_at"][4] == T - 1
        assert cols["success"][5] == 1 and cols["finished_at"][5] == T
        assert cols["failed"][6] == 1 and cols["success"][6] == 0
        assert st["n_result_ok"] == 1 and st["n_result_fail"] == 1 and s.size == 8
        s.post_result([5], [am.PHASE_FAILED])
        _, _, st2 = s.tick(T + 1)
        assert st2["n_result_fail"] == 1 and s.read([5])["failed"][0] == 1


def test_upsert_remove_and_deferred_remedy_result(am, orc, gen):
    n = 3000
    prod, orac = _gen_pair(gen, am, orc, 3, 11, n, T0)
    recs = am.columns_to_records(prod)
    with emu_sweep.EmuSweep(n + 100) as s:
        perm = np.random.default_rng(0).permutation(n)
        s.upsert(perm[: n // 2], recs[perm[: n // 2]])
        s.upsert(perm[n // 2:], recs[perm[n // 2:]])
        gone = np.array([3, 77, 1024, n - 1], dtype=np.uint64)
        s.remove(gone)
        orac["flags"][gone.astype(np.int64)] = am.F_TOMBSTONE
        got = s.tick(T0)
        assert s.size == n
        want = orc.sweep(orac, T0)
        assert got[2] == want[2]
        np.testing.assert_array_equal(got[0], want[0])
        np.testing.assert_array_equal(got[1], want[1])
        live = np.flatnonzero((orac["flags"] & 7 == am.KIND_INTERVAL) & (orac["flags"] & am.F_TOMBSTONE == 0))[:50]
        s.post_result(live, np.zeros(len(live), np.uint8), np.full(len(live), am.PHASE_FAILED, np.uint8))
        orac["flags"][live] |= am.F_REMEDY_PENDING
        got = s.tick(T0 + 3)
        want = orc.sweep(orac, T0 + 3)
        assert got[2] == want[2] and got[2]["n_remedy_fail"] == len(live)
        dev = s.read(live)
        np.testing.assert_array_equal(dev["remedy_failed"], orac["remedy_failed"][live])
        np.testing.assert_array_equal(dev["remedy_finished_at"], np.full(len(live), T0 + 3))


def test_repeat_after_sec_kernel_equals_oracle(am, orc, gen):
    n = 4000
    prod, _ = _gen_pair(gen, am, orc, 2, 2, n, T0)
    lib = orc.load()
    with emu_sweep.EmuSweep(n) as s:
        s.load_range(0, prod)
        for T in (T0, T_OCT1 - 1, 1709164799):
            got = s.repeat_after_sec(T, 0, n)
            kind = prod["flags"] & 7
            want = np.zeros(n, dtype=np.int64)
            iv = (kind == am.KIND_INTERVAL) | (kind == am.KIND_CRON_EVERY)
            want[iv] = prod["ras"][iv]
            for i in np.flatnonzero(kind == am.KIND_CRON_SPEC):
                c = orc.OrcCron(int(prod["minute"][i]), int(prod["hour"][i]), int(prod["dom"][i]),
                                int(prod["month"][i]), int(prod["dow"][i]), 0, 1, int(prod["flags"][i] >> 24))
                want[i] = lib.orc_cron_repeat_after_sec(C.byref(c), T)
            np.testing.assert_array_equal(got, want, err_msg=f"T={T}")


@pytest.mark.parametrize("workers,piece", [(1, "300"), (3, "300"), (4, None)])
def test_compiled_e2e_loop_equals_the_oracle_loop(am, orc, gen, monkeypatch, workers, piece):
    """bench.py's e2e driver (tools/amgen amgen_e2e_closed_loop: tick_view, then walk the list in pieces
    and post every submitted check as Succeeded) on the emulated library = the same loop on the oracle:
    counts of the last tick and every column afterwards."""
    n, steps = 60_000, 6
    if piece:
        monkeypatch.setenv("AMGEN_E2E_PIECE", piece)  # several pieces per tick at this size
    else:
        monkeypatch.delenv("AMGEN_E2E_PIECE", raising=False)  # the list divided evenly among the workers
    prod, orac = _gen_pair(gen, am, orc, 2, 5, n, T0)
    with emu_sweep.EmuSweep(n) as s:
        s.load_range(0, prod)
        got = gen.e2e_closed_loop(emu_sweep.load(), s._h, T0, am.SWEEP_FULL_SCAN, 2, steps - 2, n, workers=workers)
        for k in range(steps):
            idx, act, st = orc.sweep(orac, T0 + k, am.SWEEP_FULL_SCAN)
            sub = idx[(act & am.ACT_SUBMIT_HC) != 0]
            orac["flags"][sub] |= am.F_PENDING_OK
        assert (got["last_emitted"], got["last_submitted"]) == (len(idx), len(sub))
        assert len(idx) > 900  # four pieces and more
        # (every worker owns one piece of scratch: at the default piece size this population has room for three)
        assert got["h2d_bytes"] > 0 and got["d2h_bytes"] > 0 and got["workers"] == (workers if piece else min(workers, n // 16384))
        dev = s.read_range(0, n)  # (drains the last piece's posts, as the oracle's flags hold them)
        for name in am.COLUMN_NAMES:
            np.testing.assert_array_equal(dev[name], orac[name], err_msg=f"column {name}")


def test_large_post_is_chunked_and_a_bad_entry_commits_nothing(am, orc, gen):
    """am_sweep_post_result stages a large batch one copy chunk (32768 ops) at a time, handing each to the
    copy stream; an out-of-range slot or a bad phase in a LATER chunk must leave nothing of the call
    behind (the chunks already handed over are rolled back), and the next valid post must land."""
    n = 90_000
    prod, orac = _gen_pair(gen, am, orc, 2, 9, n, T0)
    slots = np.arange(0, n, dtype=np.uint64)[:80_000]
    with emu_sweep.EmuSweep(n) as s:
        s.load_range(0, prod)
        bad = slots.copy()
        bad[70_000] = n + 5
        with pytest.raises(am.AmError) as e:
            s.post_result(bad, np.full(len(bad), am.PHASE_FAILED, np.uint8))
        assert e.value.code == am.AM_E_RANGE
        ph = np.full(len(slots), am.PHASE_SUCCEEDED, np.uint8)
        ph[40_000] = 7
        with pytest.raises(am.AmError) as e:
            s.post_result(slots, ph)
        assert e.value.code == am.AM_E_INVAL
        _assert_tick_equal(am, s.tick(T0), orc.sweep(orac, T0), s, orac, n, "after two rejected posts")
        ph = np.where(slots % 3 == 0, am.PHASE_FAILED, am.PHASE_SUCCEEDED).astype(np.uint8)
        s.post_result(slots, ph)
        live = slots.astype(np.int64)
        orac["flags"][live] |= np.where(ph == am.PHASE_FAILED, am.F_PENDING_FAIL, am.F_PENDING_OK).astype(np.uint32)
        _assert_tick_equal(am, s.tick(T0 + 1), orc.sweep(orac, T0 + 1), s, orac, n, "after a chunked post")


def test_concurrent_posts_from_many_threads(am, orc, gen):
    """Six threads post results for disjoint slot sets at once (ctypes releases the GIL: the calls
    really overlap, each staging its reserved range outside the handle's lock), two of them with a
    call that must be rejected as a whole, while the main thread ticks now and then.  Every posted
    result must be applied exactly once: the counters after the last tick equal the oracle's with
    all results applied."""
    import threading
    n, per, rounds = 60_000, 9_000, 5
    prod, orac = _gen_pair(gen, am, orc, 2, 4, n, T0)
    rng = np.random.default_rng(7)
    with emu_sweep.EmuSweep(n) as s:
        s.load_range(0, prod)
        errors = []

        def worker(w):
            try:
                mine = np.arange(w * per, (w + 1) * per, dtype=np.uint64)
                for r in range(rounds):
                    part = mine[r::rounds]
                    ph = np.where(part % 2 == 0, am.PHASE_SUCCEEDED, am.PHASE_FAILED).astype(np.uint8)
                    if w in (1, 4) and r == 2:  # a rejected call in the middle of the traffic
                        bad = part.copy()
                        bad[len(bad) // 2] = n + 1
                        try:
                            s.post_result(bad, ph)
                            errors.append("bad call accepted")
                        except am.AmError as e:
                            assert e.code == am.AM_E_RANGE
                    s.post_result(part, ph)
            except Exception as e:  # noqa: BLE001
                errors.append(repr(e))

        th = [threading.Thread(target=worker, args=(w,)) for w in range(6)]
        for t in th:
            t.start()
        s.tick(T0)  # a tick in the middle: drains whatever complete calls it finds
        for t in th:
            t.join()
        assert not errors, errors
        s.tick(T0 + 1)
        dev = s.read_range(0, n)
        # results are posted once per slot: whichever of the two ticks applied one, the counters add up
        slots = np.arange(0, 6 * per)
        live = ((0x3E >> (prod["flags"][slots] & am.KIND_MASK)) & 1).astype(bool) & ((prod["flags"][slots] & am.F_TOMBSTONE) == 0)
        ok = slots % 2 == 0
        np.testing.assert_array_equal(dev["success"][slots] - prod["success"][slots], (live & ok).astype(np.int32))
        np.testing.assert_array_equal(dev["failed"][slots] - prod["failed"][slots], (live & ~ok).astype(np.int32))
        assert not np.any(dev["flags"][slots][live] & (am.F_PENDING_OK | am.F_PENDING_FAIL))
        rest = np.arange(6 * per, n)
        np.testing.assert_array_equal(dev["success"][rest], prod["success"][rest])


@pytest.mark.parametrize("every,early", [(23, None), (3, None), (23, "0")])
def test_posted_results_sparse_and_dense_paths(am, orc, gen, monkeypatch, every, early):
    """Results posted between ticks are applied at the tick's drain by apply_results_now_kernel when they
    are sparse (ops <= 1/8 of the records: `every` = 23) — their action bits travel to the same tick's sweep in
    the flags' carry bits — and inside the sweep when they are dense (`every` = 3).  Both must equal the
    oracle: list, action bytes (RUN_REMEDY, REMEDY_SKIP, RESET_ON_PASS, RESET_ON_INTERVAL, ANOMALY all
    occur in the config-3 mix), statistics, every column; over three ticks, with workflow and remedy phases
    posted by separate calls."""
    if early is not None:  # "0": the sparse batch too goes through the flags and the sweep's own result path
        monkeypatch.setenv("AMSWEEP_EARLY_RESULTS", early)
    n = 40_000
    prod, orac = _gen_pair(gen, am, orc, 3, 12, n, T0)
    pend = am.F_PENDING_OK | am.F_PENDING_FAIL | am.F_REMEDY_PENDING | am.F_REMEDY_OUTCOME_OK
    for c in (prod, orac):  # the posted results come through the C-ABI here, not with the population
        c["flags"] &= ~np.uint32(pend)
    seen = 0
    with emu_sweep.EmuSweep(n) as s:
        s.load_range(0, prod)
        for k in range(3):
            T = T0 + 1800 * k
            slots = np.arange(k, n, every, dtype=np.uint64)
            ph = (1 + (slots * 7 + k) % 2).astype(np.uint8)        # Succeeded / Failed
            rp = ((slots // every + k) % 3).astype(np.uint8)         # none / Succeeded / Failed
            s.post_result(slots, ph)                                 # workflow phase ...
            with_r = rp != 0
            s.post_result(slots[with_r], np.zeros(int(with_r.sum()), np.uint8), rp[with_r])  # ... remedy phase, separately
            i = slots.astype(np.int64)
            orac["flags"][i] |= np.where(ph == 1, am.F_PENDING_OK, am.F_PENDING_FAIL).astype(np.uint32)
            orac["flags"][i[with_r]] |= np.where(rp[with_r] == 1, am.F_REMEDY_PENDING | am.F_REMEDY_OUTCOME_OK,
                                                 am.F_REMEDY_PENDING).astype(np.uint32)
            l0 = s.launch_count
            got, want = s.tick(T), orc.sweep(orac, T)
            # mark, apply_result_ops OR apply_results_now, clear_marks + tz_words (on-minute tick, zones registered), sweep, scan, expand, publish
            assert s.launch_count - l0 == 8, s.launch_count - l0
            _assert_tick_equal(am, got, want, s, orac, n, f"every={every} tick {k}")
            seen |= int(np.bitwise_or.reduce(want[1]))
        assert not np.any(s.read_range(0, n)["flags"] & am.F_CARRY_MASK)
    assert seen & 0xF2 == 0xF2, hex(seen)  # every result-driven action bit occurred
