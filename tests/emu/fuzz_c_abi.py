"""Child process of tests/test_c_abi_on_emulator.py: random call sequences against the emulated
library through the real C-ABI (am.Sweep), checked against a sequential model.

The model is what the reference's concurrency contract promises per CR: calls take
effect in call order — an upsert replaces the record, a remove forgets it, a posted result marks
it — and a tick is the oracle's sweep over the resulting columns.  The library resolves the order
on the device (mark / apply / clear kernels); here every batch mixes duplicate slots, results
before and after upserts, removes of absent slots, reads between the calls, short output buffers,
open- and closed-loop ticks.

usage: fuzz_c_abi.py <sequences> <steps per sequence> <capacity>"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
am = importlib.import_module("active-monitor_b200")
assert "emu" in os.path.basename(am.abi.LIB_PATH), "run with AMSWEEP_LIB pointing at libamsweep_emu.so"
import oracle_c  # noqa: E402

T0 = 1789982100
SPECS = [dict(repeat_after_sec=5), dict(repeat_after_sec=60), dict(repeat_after_sec=3600), dict(cron="@every 7s"),
         dict(cron="* * * * *"), dict(cron="*/2 * * * *"), dict(cron="16 9 * * mon"), dict(cron="NOT_A_VALID_CRON"),
         dict(), dict(repeat_after_sec=30, has_resource=False), dict(cron="0 0 30 2 *")]
M = am.F_PENDING_OK | am.F_PENDING_FAIL | am.F_REMEDY_PENDING | am.F_REMEDY_OUTCOME_OK
HC = am.F_PENDING_OK | am.F_PENDING_FAIL
RM = am.F_REMEDY_PENDING | am.F_REMEDY_OUTCOME_OK
PB = {0: 0, 1: am.F_PENDING_OK, 2: am.F_PENDING_FAIL}
RB = {0: 0, 1: am.F_REMEDY_PENDING | am.F_REMEDY_OUTCOME_OK, 2: am.F_REMEDY_PENDING}


def random_record(rng, T):
    kw = dict(SPECS[int(rng.integers(0, len(SPECS)))])
    kw.update(has_remedy=bool(rng.integers(0, 2)), remedy_runs_limit=int(rng.choice([0, 1, 2, 5])),
              remedy_reset_interval=int(rng.choice([0, 60, 300])), fail_p8=int(rng.integers(0, 256)),
              finished_at=None if rng.integers(0, 6) == 0 else T - int(rng.integers(0, 200)),
              timer_armed=bool(rng.integers(0, 4)),
              success_count=int(rng.integers(0, 50)), failed_count=int(rng.integers(0, 50)))
    rs, rf = int(rng.integers(0, 4)), int(rng.integers(0, 4))
    kw.update(remedy_success_count=rs, remedy_failed_count=rf, remedy_total_runs=rs + rf,
              remedy_finished_at=None if rs + rf == 0 else T - int(rng.integers(1, 400)))
    rc, rec = am.classify(**kw)
    assert rc == 0
    return rec[0]


def assert_columns_equal(got, want, what):
    """flags everywhere; the other columns only for live slots (a removed slot keeps nothing but its
    tombstone: whatever an overtaken upsert of the same batch carried is a don't-care)."""
    np.testing.assert_array_equal(got["flags"], want["flags"], err_msg=f"{what} flags")
    live = (want["flags"] & np.uint32(am.F_TOMBSTONE)) == 0
    for name in am.COLUMN_NAMES:
        np.testing.assert_array_equal(got[name][live], want[name][live], err_msg=f"{what} {name}")


def run_sequence(seed, steps, cap):
    rng = np.random.default_rng(seed)
    model = am.alloc_columns(cap)
    model["flags"][:] = am.F_TOMBSTONE
    hw = 0  # high-water mark
    T = T0 - 30 + int(rng.integers(0, 60))
    base = int(rng.choice([0, 12_500_000]))
    with am.Sweep(capacity=cap, shard_base=base) as s:
        s.set_seed(seed)
        for step in range(steps):
            for _ in range(int(rng.integers(0, 5))):  # a few staged calls between two ticks
                op = int(rng.integers(0, 10))
                k = int(rng.integers(1, 40))
                idx = rng.integers(0, cap if rng.integers(0, 4) == 0 else max(1, min(cap, hw + 20)), k)
                if rng.integers(0, 3) == 0:
                    idx = np.repeat(idx[: max(1, k // 3)], 3)[:k]  # duplicates inside one call
                if op < 4:
                    recs = np.array([random_record(rng, T) for _ in idx], dtype=am.RECORD_DTYPE)
                    s.upsert(idx, recs)
                    for i, r in zip(idx.tolist(), recs):
                        for name in am.COLUMN_NAMES:
                            model[name][i] = r[name]
                        model["flags"][i] &= ~np.uint32(am.F_TOMBSTONE)
                    hw = max(hw, int(idx.max()) + 1)
                elif op < 5:
                    s.remove(idx)
                    model["flags"][idx] = am.F_TOMBSTONE
                elif op < 9:
                    ph = rng.integers(0, 3, len(idx)).astype(np.uint8)
                    rp = rng.integers(0, 3, len(idx)).astype(np.uint8) if rng.integers(0, 2) else None
                    s.post_result(idx, ph, rp)
                    # the two phases are observed by separate watch loops: each call replaces only the
                    # bit group(s) it carries a phase for ("none" leaves a group alone)
                    for j, i in enumerate(idx.tolist()):
                        if int(ph[j]):
                            model["flags"][i] = (model["flags"][i] & ~np.uint32(HC)) | np.uint32(PB[int(ph[j])])
                        if rp is not None and int(rp[j]):
                            model["flags"][i] = (model["flags"][i] & ~np.uint32(RM)) | np.uint32(RB[int(rp[j])])
                else:  # a read between the calls observes everything staged before it
                    q = rng.integers(0, cap, 7)
                    assert_columns_equal(s.read(q), {k: v[q] for k, v in model.items()}, f"seed {seed} step {step} read")
            mode = int(rng.choice([0, 0, am.SWEEP_FULL_SCAN, am.SWEEP_CLOSED_LOOP]))
            T += int(rng.choice([0, 1, 1, 1, 7, 60]))
            view = {k: v[:hw] for k, v in model.items()}  # the oracle sweeps the same hw slots, in place
            wi, wa, ws = oracle_c.sweep(view, T, mode=mode & 1, seed=seed, shard_base=base) if hw else (
                np.zeros(0, np.uint64), np.zeros(0, np.uint32), None)
            if hw and len(wi) > 3 and rng.integers(0, 5) == 0:  # short buffers: AM_E_NOSPACE, prefix valid
                try:
                    s.tick(T, mode=mode, cap=3)
                    raise AssertionError("expected AM_E_NOSPACE")
                except am.AmError as e:
                    assert e.code == am.AM_E_NOSPACE and e.needed == len(wi)
                    np.testing.assert_array_equal(e.partial[0], wi[:3])
                    gs = e.partial[2]
                    gi, ga = wi, wa  # the tick happened: state advanced exactly once
            else:
                gi, ga, gs = s.tick(T, mode=mode)
            assert s.size == hw, (seed, step, s.size, hw)  # high-water mark, updated when the staged calls drain
            if hw:
                assert gs == ws, (seed, step, {f: (gs[f], ws[f]) for f in gs if gs[f] != ws[f]})
                np.testing.assert_array_equal(gi, wi, err_msg=f"seed {seed} step {step} idx")
                np.testing.assert_array_equal(ga, wa, err_msg=f"seed {seed} step {step} act")
            if rng.integers(0, 4) == 0 and hw:
                assert_columns_equal(s.read_range(0, hw), {k: v[:hw] for k, v in model.items()}, f"seed {seed} step {step}")
        assert_columns_equal(s.read_range(0, cap), model, f"seed {seed} final")


if __name__ == "__main__":
    nseq, steps, cap = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    for seed in range(nseq):
        run_sequence(seed, steps, cap)
    print(f"ok {nseq} sequences x {steps} steps, capacity {cap}")
