"""Pin the CPU oracle before trusting it.

1. Everything the REFERENCE's own tests hold for this path — only four facts:
     healthcheck_controller_unit_test.go:617-634  "NOT_A_VALID_CRON" is an error
     healthcheck_controller_unit_test.go:636-660  "@every 5s" => RepeatAfterSec > 0
     healthcheck_controller_test.go:119-156 + examples/bdd/inlineHelloTest.yaml:8  pause => Stopped
     api/v1alpha1/healthcheck_types_unit_test.go:24-38  RemedyWorkflow.IsEmpty truth table
   plus healthcheck_controller_edge_test.go:47-74 (nil Workflow.Resource => untouched).
2. SURVEY Appendix C known-answer vectors (hand-derivable from robfig v3.0.1's
   published semantics) and vectors RECALLED from robfig's own spec_test.go /
   parser_test.go (flagged: recalled, the module is not in /root/reference).
3. C oracle == independent Python oracle, on fixed and on random inputs.
4. Algebraic identities: matches(T) <=> next(T-1)==T; next(T) > T; matches(next(T)).

5-field cron parity against the Go binary remains UNPINNED (no Go toolchain).
"""
import ctypes as C
import datetime as dt
import random

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

T0 = 1789982100  # 2026-09-21 09:15:00 UTC Monday


def utc(y, mo, d, h=0, mi=0, s=0):
    return int(dt.datetime(y, mo, d, h, mi, s, tzinfo=dt.timezone.utc).timestamp())


def c_parse(orc, spec):
    rc, c, msg = orc.cron_parse(spec)
    return rc, (c.kind, c.minute, c.hour, c.dom, c.month, c.dow, c.delay_sec), c, msg


def py_parse(opy, spec):
    try:
        c = opy.cron_parse(spec)
        return 0, (c.kind, c.minute, c.hour, c.dom, c.month, c.dow, c.delay_sec), c
    except opy.CronUnsupported:
        return -7, None, None
    except opy.CronError:
        return -6, None, None


# ---------------------------------------------------------------- 1. reference-pinned
def test_reference_invalid_cron_is_error(orc, opy):
    """healthcheck_controller_unit_test.go:617-634"""
    assert c_parse(orc, "NOT_A_VALID_CRON")[0] == -6
    assert py_parse(opy, "NOT_A_VALID_CRON")[0] == -6


def test_reference_every_5s_sets_positive_repeat(orc, opy):
    """healthcheck_controller_unit_test.go:636-660: RepeatAfterSec > 0 (exactly 5 by hcc.go:259-262)"""
    rc, vals, c, _ = c_parse(orc, "@every 5s")
    assert rc == 0 and vals[0] == 2 and vals[6] == 5
    assert orc.load().orc_cron_repeat_after_sec(C.byref(c), T0) == 5 > 0
    assert opy.cron_repeat_after_sec(opy.cron_parse("@every 5s"), T0) == 5


def test_reference_fixture_schedules(orc):
    """cron strings that appear in the reference's fixtures/examples"""
    for spec, delay in [("@every 1m", 60), ("@every 3s", 3), ("@every 5s", 5)]:
        rc, vals, _, _ = c_parse(orc, spec)
        assert rc == 0 and vals[0] == 2 and vals[6] == delay, spec


def test_reference_pause_rule_and_nil_resource(orc, opy):
    """repeatAfterSec: 0, no cron => Stopped (healthcheck_controller_test.go:119-156);
    Workflow.Resource == nil => nothing happens (edge_test.go:47-74)."""
    for ras in (0, -1, -100):
        rc, r = opy.classify(opy.HealthCheck(repeat_after_sec=ras, cron=""))
        assert rc == 0 and r.kind == opy.KIND_STOPPED
        act = opy.tick_record(r, T0)
        assert act == opy.ACT_STOPPED and r.finished_at == T0
        assert opy.tick_record(r, T0 + 1) == 0  # reported once
    rc, r = opy.classify(opy.HealthCheck(repeat_after_sec=60, has_resource=False))
    assert r.kind == opy.KIND_NO_RESOURCE and opy.tick_record(r, T0) == 0


def test_reference_remedy_is_empty_truth_table(orc, opy):
    """api/v1alpha1/healthcheck_types_unit_test.go:24-38"""
    cases = [(("", True, 0, True), True), (("remedy-", True, 0, True), False),
             (("", False, 0, True), False), (("", True, 30, True), False),
             (("", True, 0, False), False)]  # non-nil empty rbacRules: DeepEqual says different
    for (name, res_nil, timeout, rbac_nil), want in cases:
        assert bool(orc.load().orc_remedy_is_empty(len(name), res_nil, timeout, rbac_nil)) is want
        assert opy.remedy_is_empty(name, res_nil, timeout, rbac_nil) is want


def test_ladder_order_cron_ignored_when_interval_positive(orc, opy):
    """hcc.go:251 vs :264 (examples/inlineHello_cluster_cron_repeat.yaml): unpinned by the reference"""
    rc, r = opy.classify(opy.HealthCheck(repeat_after_sec=60, cron="NOT_A_VALID_CRON"))
    assert r.kind == opy.KIND_INTERVAL and r.ras == 60
    rc, r = opy.classify(opy.HealthCheck(repeat_after_sec=0, cron="NOT_A_VALID_CRON"))
    assert r.kind == opy.KIND_PARSE_ERROR
    assert opy.tick_record(r, T0) == opy.ACT_PARSE_ERROR and opy.tick_record(r, T0 + 1) == opy.ACT_PARSE_ERROR


# ---------------------------------------------------------------- 2. known answers
ALL_MIN, ALL_HR = 0x8FFFFFFFFFFFFFFF, 0x8000000000FFFFFF
ALL_DOM, ALL_MON, ALL_DOW = 0x80000000FFFFFFFE, 0x8000000000001FFE, 0x800000000000007F

APPENDIX_C = [
    ("* * * * *", (ALL_MIN, ALL_HR, ALL_DOM, ALL_MON, ALL_DOW)),
    ("*/15 9-17 * * 1-5", (0x200040008001, 0x3FE00, ALL_DOM, ALL_MON, 0x3E)),
    ("0 0 1 1 *", (1, 1, 2, 2, ALL_DOW)),
    ("@yearly", (1, 1, 2, 2, ALL_DOW)),
    ("@annually", (1, 1, 2, 2, ALL_DOW)),
    ("@monthly", (1, 1, 2, ALL_MON, ALL_DOW)),
    ("@weekly", (1, 1, ALL_DOM, ALL_MON, 1)),
    ("@daily", (1, 1, ALL_DOM, ALL_MON, ALL_DOW)),
    ("@midnight", (1, 1, ALL_DOM, ALL_MON, ALL_DOW)),
    ("@hourly", (1, ALL_HR, ALL_DOM, ALL_MON, ALL_DOW)),
    ("0 0 */2 * 1", (1, 1, 0xAAAAAAAA, ALL_MON, 2)),
    ("0 0 29 2 *", (1, 1, 1 << 29, 1 << 2, ALL_DOW)),
    ("5/15 * * * *", (1 << 5 | 1 << 20 | 1 << 35 | 1 << 50, ALL_HR, ALL_DOM, ALL_MON, ALL_DOW)),
    ("*/1 * * * *", (ALL_MIN, ALL_HR, ALL_DOM, ALL_MON, ALL_DOW)),      # step 1 keeps the star
    ("*/2 * ? * *", (0x555555555555555, ALL_HR, ALL_DOM, ALL_MON, ALL_DOW)),  # step>1 drops it
    ("*-5 * * * *", (ALL_MIN, ALL_HR, ALL_DOM, ALL_MON, ALL_DOW)),       # quirk: "*-5" is "*"
    ("0 0 * JAN-mar Sun,sAt", (1, 1, ALL_DOM, 0xE, 0x41)),
    ("+5 * * * *", (1 << 5, ALL_HR, ALL_DOM, ALL_MON, ALL_DOW)),         # strconv.Atoi takes a sign
    ("1,,2 * * * *", (6, ALL_HR, ALL_DOM, ALL_MON, ALL_DOW)),            # FieldsFunc drops empties
    (", * * * *", (0, ALL_HR, ALL_DOM, ALL_MON, ALL_DOW)),               # empty mask, never fires
    ("  0   0\t1 1  *  ", (1, 1, 2, 2, ALL_DOW)),                        # strings.Fields
    ("TZ=UTC 0 0 1 1 *", (1, 1, 2, 2, ALL_DOW)),
    ("CRON_TZ=UTC   0 0 1 1 *", (1, 1, 2, 2, ALL_DOW)),
    ("TZ= 0 0 1 1 *", (1, 1, 2, 2, ALL_DOW)),                            # LoadLocation("") is UTC
]

REJECTED = ["", "* * * * 7", "60 * * * *", "* * * *", "* * * * * *", "NOT_A_VALID_CRON", "* 24 * * *",
            "* * 0 * *", "* * 32 * *", "* * * 0 *", "* * * 13 *", "*/0 * * * *", "5-1 * * * *",
            "1-2-3 * * * *", "1/2/3 * * * *", "a * * * *", "- * * * *", "1- * * * *", "/5 * * * *",
            "5/ * * * *", "1/-1 * * * *", "* * * janu *", "* * * * sunday", "@every", "@every ",
            "@every 5", "@every 5 s", "@every s", "@every 1d", "@fortnightly", "@Daily", "@daily ",
            " @daily", "TZ=UTC", "CRON_TZ=UTC", "5_0 * * * *", "٥ * * * *", "* * * * * x",
            "99999999999999999999 * * * *", "@every 9223372036854775808ns", "@every 1e3s"]

# named zones: accepted (time.LoadLocation finds them) and carried as a zone id; unknown ones are errors
UNSUPPORTED = []  # (round 1 reported every named zone as "unsupported"; only a 256th distinct zone is, now)
ZONED = ["CRON_TZ=America/New_York * * * * *", "TZ=Asia/Tokyo 0 0 * * *", "CRON_TZ=Europe/Paris @daily"]
REJECTED_ZONES = ["TZ=Nowhere/Land * * * * *", "CRON_TZ=../etc/passwd * * * * *", "TZ=/usr/share/zoneinfo/UTC * * * * *",
                  "CRON_TZ=America/New_York", "TZ=Europe/Paris 61 * * * *"]

EVERY = [("@every 5s", 5), ("@every 500ms", 1), ("@every 90s", 90), ("@every 1h30m", 5400),
         ("@every 1.5s", 1), ("@every 1m", 60), ("@every 0", 1), ("@every -5s", 1),
         ("@every 1h1m1s1ms1us1ns", 3661), ("@every 1.9999999999s", 1), ("@every .5m", 30),
         ("@every 1µs", 1), ("@every 1μs", 1), ("@every +2h", 7200), ("@every 2540400h", 9145440000),
         ("@every 100000000000000ns", 100000), ("@every 0.000000001h", 1)]


@pytest.mark.parametrize("spec,masks", APPENDIX_C)
def test_known_masks(orc, opy, spec, masks):
    rc, vals, _, msg = c_parse(orc, spec)
    assert rc == 0, msg
    assert vals[0] == 1 and vals[1:6] == masks, [hex(v) for v in vals[1:6]]
    prc, pvals, _ = py_parse(opy, spec)
    assert prc == 0 and pvals == vals


@pytest.mark.parametrize("spec", REJECTED)
def test_rejected_specs(orc, opy, spec):
    assert c_parse(orc, spec)[0] == -6
    assert py_parse(opy, spec)[0] == -6


@pytest.mark.parametrize("spec", ZONED)
def test_named_zones_are_resolved(orc, opy, spec):
    """time.LoadLocation succeeds: a SpecSchedule with a Location (tz_id != 0), same masks as without the prefix"""
    rc, vals, c, msg = c_parse(orc, spec)
    assert rc == 0 and c.tz_id > 0, msg
    bare = spec.split(" ", 1)[1]
    assert vals == c_parse(orc, bare)[1]
    prc, pvals, pc = py_parse(opy, spec)
    assert prc == 0 and pvals == vals and pc.tz_id > 0


@pytest.mark.parametrize("spec", REJECTED_ZONES)
def test_unknown_zones_and_bad_specs_behind_a_zone_are_errors(orc, opy, spec):
    assert c_parse(orc, spec)[0] == -6
    assert py_parse(opy, spec)[0] == -6


@pytest.mark.parametrize("spec,delay", EVERY)
def test_every_delays(orc, opy, spec, delay):
    rc, vals, _, _ = c_parse(orc, spec)
    assert rc == 0 and vals[0] == 2 and vals[6] == delay
    prc, pvals, _ = py_parse(opy, spec)
    assert prc == 0 and pvals == vals


ACTIVATION = [  # (spec, UTC time, fires?)  SURVEY Appendix C
    ("* * * * *", utc(2026, 9, 21, 9, 15), True),
    ("*/15 9-17 * * 1-5", utc(2026, 9, 21, 9, 15), True),
    ("*/15 9-17 * * 1-5", 1789971971, False),                 # 06:26:11: sec != 0
    ("0 0 1 1 *", utc(2026, 9, 21, 9, 15), False),
    ("0 0 1 1 *", 1767225600, True),                          # 2026-01-01 00:00:00
    ("@yearly", 1767225600, True),
    ("0 0 */2 * 1", utc(2026, 9, 21), True),                  # Mon and dom 21 odd
    ("0 0 */2 * 1", utc(2026, 9, 22), False),                 # Tue, dom 22 even
    ("0 0 */2 * 1", utc(2026, 9, 28), True),                  # Mon; dom 28 not in mask: OR rule
    ("0 0 29 2 *", 1709164800, True),                         # leap day 2024
    ("0 0 29 2 *", utc(2025, 3, 1), False),
    # recalled from robfig/cron v3 spec_test.go TestActivation (module not vendored; from memory)
    ("0/15 * * * *", utc(2012, 7, 9, 15, 0), True),
    ("0/15 * * * *", utc(2012, 7, 9, 15, 45), True),
    ("0/15 * * * *", utc(2012, 7, 9, 15, 40), False),
    ("5/15 * * * *", utc(2012, 7, 9, 15, 5), True),
    ("5/15 * * * *", utc(2012, 7, 9, 15, 20), True),
    ("5/15 * * * *", utc(2012, 7, 9, 15, 50), True),
    ("10-30/15 * * * *", utc(2012, 7, 9, 15, 10), True),
    ("10-30/15 * * * *", utc(2012, 7, 9, 15, 25), True),
    ("10-30/15 * * * *", utc(2012, 7, 9, 15, 40), False),
    ("* * 1,15 * Sun", utc(2012, 7, 15), True),               # both restricted: either matches
    ("* * 1,15 * Sun", utc(2012, 6, 15), True),
    ("* * 1,15 * Sun", utc(2012, 8, 1), True),
    ("* * */10 * Sun", utc(2012, 7, 15), True),               # verifies robfig #70
    ("* * * * Mon", utc(2012, 7, 15), False),                 # a star: both must match
    ("* * 1,15 * *", utc(2012, 7, 9), False),
    ("* * 1,15 * *", utc(2012, 7, 15), True),
    ("* * */2 * Sun", utc(2012, 7, 15), True),
]


@pytest.mark.parametrize("spec,T,want", ACTIVATION)
def test_activation(orc, opy, spec, T, want):
    rc, _, c, _ = c_parse(orc, spec)
    assert rc == 0
    lib = orc.load()
    assert bool(lib.orc_cron_matches(C.byref(c), T)) is want
    assert (lib.orc_cron_next(C.byref(c), T - 1) == T) is want   # robfig's own test idiom
    pc = opy.cron_parse(spec)
    assert opy.cron_matches(pc, T) is want
    assert (opy.cron_next(pc, T - 1) == T) is want


NEXT = [  # recalled from robfig/cron v3 spec_test.go TestNext, translated to 5 fields / UTC
    ("0/15 * * * *", utc(2012, 7, 9, 14, 45), utc(2012, 7, 9, 15, 0)),
    ("0/15 * * * *", utc(2012, 7, 9, 14, 59), utc(2012, 7, 9, 15, 0)),
    ("0/15 * * * *", utc(2012, 7, 9, 14, 59, 59), utc(2012, 7, 9, 15, 0)),
    ("20-35/15 * * * *", utc(2012, 7, 9, 15, 45), utc(2012, 7, 9, 16, 20)),  # wrap around hours
    ("0 0 * Feb Mon", utc(2012, 7, 9, 23, 35), utc(2013, 2, 4)),            # wrap around years
    ("0 0 * Feb Mon/2", utc(2012, 7, 9, 23, 35), utc(2013, 2, 1)),
    ("0 0 29 Feb ?", utc(2012, 7, 9, 23, 35), utc(2016, 2, 29)),            # leap year
    ("0 0 31 Apr ?", utc(2012, 7, 9, 23, 35), None),                        # unsatisfiable
    ("0 0 30 Feb ?", utc(2012, 7, 9, 23, 35), None),
    ("0 0 31 * *", utc(2012, 4, 30), utc(2012, 5, 31)),
    ("0 0 29 2 *", utc(2096, 3, 1), utc(2104, 2, 29) if False else None),   # 2100 is not leap: > 5 years
    ("@yearly", utc(2026, 9, 21, 9, 15), utc(2027, 1, 1)),
    ("59 23 31 12 *", utc(2026, 12, 31, 23, 58, 59), utc(2026, 12, 31, 23, 59)),
    ("59 23 31 12 *", utc(2026, 12, 31, 23, 59), utc(2027, 12, 31, 23, 59)),
]


@pytest.mark.parametrize("spec,T,want", NEXT)
def test_next(orc, opy, spec, T, want):
    rc, _, c, _ = c_parse(orc, spec)
    assert rc == 0
    got = orc.load().orc_cron_next(C.byref(c), T)
    assert (None if got == -(1 << 63) else got) == want
    assert opy.cron_next(opy.cron_parse(spec), T) == want


REMEDY_GATE = [  # (runsLimit, resetInterval, RT, d|None) -> action
    (2, 300, 1, 10, 0x02), (2, 300, 2, 300, 0x10), (2, 300, 2, 301, 0x42), (0, 300, 9, 10, 0x02),
    (2, 0, 9, 10, 0x02), (2, 300, 2, None, 0x80), (-1, 300, 0, None, 0x80), (-1, -5, 0, 7, 0x42),
    (1, 60, 1, 60, 0x10), (1, 60, 1, 61, 0x42), (5, 1, 5, -10, 0x10),
]


@pytest.mark.parametrize("lim,rst,rt,d,want", REMEDY_GATE)
def test_remedy_gate_vectors(orc, opy, lim, rst, rt, d, want):
    r = opy.Record(ras=3600, flags=opy.KIND_INTERVAL | opy.F_HAS_REMEDY | opy.F_PENDING_FAIL,
                   finished_at=T0 - 5, runs_limit=lim, reset_interval=rst, remedy_total=rt,
                   remedy_failed=rt, remedy_finished_at=0 if d is None else T0 - d)
    assert opy.tick_record(r, T0) == want
    # the same record through the C oracle
    cols = {n: np.zeros(1, dtype=t) for n, t in orc.COLUMNS}
    cols["ras"][0], cols["flags"][0], cols["finished_at"][0] = 3600, 2 | 8 | 32, T0 - 5
    cols["runs_limit"][0], cols["reset_interval"][0] = lim, rst
    cols["remedy_total"][0] = cols["remedy_failed"][0] = rt
    cols["remedy_finished_at"][0] = 0 if d is None else T0 - d
    idx, act, _ = orc.sweep(cols, T0)
    assert act.tolist() == [want]
    assert cols["failed"][0] == 1 and cols["finished_at"][0] == T0
    assert cols["remedy_total"][0] == r.remedy_total and cols["remedy_finished_at"][0] == r.remedy_finished_at


def test_success_reset_threshold(opy):
    """hcc.go:649: reset only when RemedyTotalRuns >= 1 and the remedy is non-empty"""
    for rt, has, want in [(1, True, 0x20), (0, True, 0), (3, False, 0)]:
        r = opy.Record(ras=3600, finished_at=T0 - 5, remedy_total=rt, remedy_success=rt,
                       remedy_finished_at=T0 - 9 if rt else 0,
                       flags=opy.KIND_INTERVAL | (opy.F_HAS_REMEDY if has else 0) | opy.F_PENDING_OK)
        assert opy.tick_record(r, T0) == want
        assert r.success == 1 and r.finished_at == T0
        assert r.remedy_total == (0 if want else rt)


# ---------------------------------------------------------------- 3. C oracle == Python oracle
TOKENS = ["*", "?", "*/2", "*/15", "*/0", "0", "1", "5", "7", "12", "23", "24", "31", "59", "60",
          "1-5", "5-1", "0-59", "1-31/7", "10/20", "jan", "DEC", "Feb-apr", "sun", "SAT", "mon-fri",
          "1,2,3", "1,,3", ",", "a", "-", "1-", "/2", "1/2/3", "1-2-3", "+3", "-3", "007", "*-3",
          "?/3", "1-5/2,10", "fri-mon", "13", "0-6", "6-7", "2147483648", "１"]


@settings(max_examples=600, deadline=None)
@given(st.lists(st.sampled_from(TOKENS), min_size=4, max_size=6),
       st.sampled_from([" ", "  ", "\t", " \t "]), st.sampled_from(["", "", "", "TZ=UTC ", "CRON_TZ=UTC  "]))
def test_c_oracle_equals_python_oracle_on_token_specs(orc, opy, toks, sep, prefix):
    spec = prefix + sep.join(toks)
    rc, vals, c, _ = c_parse(orc, spec)
    prc, pvals, pc = py_parse(opy, spec)
    assert rc == prc, spec
    if rc == 0:
        assert vals == pvals, spec
        for T in (T0, utc(2024, 2, 29), utc(2026, 12, 31, 23, 59), utc(2027, 1, 1), T0 + 17):
            assert bool(orc.load().orc_cron_matches(C.byref(c), T)) == opy.cron_matches(pc, T)


@settings(max_examples=300, deadline=None)
@given(st.text(alphabet="0123456789*/-,? \t@everyjanfmsuTZ=UTC+.hmsµ", min_size=0, max_size=24))
def test_c_oracle_equals_python_oracle_on_random_text(orc, opy, spec):
    rc, vals, _, _ = c_parse(orc, spec)
    prc, pvals, _ = py_parse(opy, spec)
    assert rc == prc, repr(spec)
    if rc == 0:
        assert vals == pvals, repr(spec)


@settings(max_examples=300, deadline=None)
@given(st.text(alphabet="0123456789.hmsnuµμ+- ", min_size=0, max_size=16))
def test_parse_duration_c_equals_python(orc, opy, s):
    raw = s.encode()
    out = C.c_int64()
    rc = orc.load().orc_parse_duration(raw, len(raw), C.byref(out))
    try:
        want = opy.parse_duration(s)
        assert rc == 0 and out.value == want, s
    except opy.CronError:
        assert rc != 0, s


# ---------------------------------------------------------------- 4. identities
def _random_field(rng, lo, hi):
    k = rng.random()
    if k < 0.35:
        return "*"
    if k < 0.5:
        return f"*/{rng.choice([2, 3, 5, 7, 10, 15])}"
    if k < 0.7:
        return str(rng.randint(lo, hi))
    if k < 0.85:
        a = rng.randint(lo, hi)
        return f"{a}-{rng.randint(a, hi)}"
    return ",".join(str(rng.randint(lo, hi)) for _ in range(rng.randint(2, 4)))


def test_next_and_matches_identities(orc, opy):
    rng = random.Random(20260921)
    lib = orc.load()
    for _ in range(400):
        spec = " ".join(_random_field(rng, lo, hi) for lo, hi in [(0, 59), (0, 23), (1, 31), (1, 12), (0, 6)])
        rc, _, c, _ = c_parse(orc, spec)
        assert rc == 0, spec
        pc = opy.cron_parse(spec)
        T = rng.randint(utc(2020, 1, 1), utc(2035, 1, 1))
        nx = lib.orc_cron_next(C.byref(c), T)
        pnx = opy.cron_next(pc, T)
        assert (None if nx == -(1 << 63) else nx) == pnx, (spec, T)   # Go field-walk == day scan
        if pnx is not None:
            assert nx > T and nx % 60 == 0
            assert lib.orc_cron_matches(C.byref(c), nx) == 1
            assert lib.orc_cron_next(C.byref(c), nx - 1) == nx
            # nothing fires strictly between T and next(T): probe a few whole minutes
            for probe in range(T - T % 60 + 60, min(nx, T + 3600), 60):
                assert lib.orc_cron_matches(C.byref(c), probe) == 0, (spec, T, probe)
        m = bool(lib.orc_cron_matches(C.byref(c), T - T % 60))
        assert m == (lib.orc_cron_next(C.byref(c), T - T % 60 - 1) == T - T % 60)


def test_civil_time_c_vs_python(orc, opy):
    rng = random.Random(7)
    out = (C.c_int32 * 6)()
    for T in [0, -1, 86399, 86400, 951782400, 1709164800, T0, 4102444800, -2208988800] + \
             [rng.randint(-2**40, 2**40) for _ in range(2000)]:
        try:
            want = opy.civil_from_unix(T)
        except OverflowError:
            continue
        orc.load().orc_civil_from_unix(T, C.byref(out))
        assert tuple(out) == want, T


def test_sweep_c_equals_python_on_generated_population(orc, opy, gen):
    """whole-tick agreement of the two oracles on the config-3 population (all rules)"""
    n = 3000
    cols = gen.fill(3, 3, 0, n, T0, orc.load().orc_classify)
    recs = []
    for i in range(n):
        recs.append(opy.Record(**{name: int(cols[name][i]) for name, _ in orc.COLUMNS}))
    for T in (T0, T0 + 1, T0 + 60):
        due_py, st_py = opy.sweep(recs, T)
        idx, act, st_c = orc.sweep(cols, T)
        assert [(int(a), int(b)) for a, b in zip(idx, act)] == due_py
        for k, v in st_c.items():
            assert getattr(st_py, k) == v, k
        for i in (0, 1, 17, n - 1):
            for name, _ in orc.COLUMNS:
                assert int(cols[name][i]) == getattr(recs[i], name)
    for i in range(n):
        for name, _ in orc.COLUMNS:
            assert int(cols[name][i]) == getattr(recs[i], name), (i, name)


def test_sweep_mt_equals_single_thread(orc, gen):
    n = 50_000
    a = gen.fill(3, 3, 0, n, T0, orc.load().orc_classify)
    b = {k: v.copy() for k, v in a.items()}
    ia, aa, sa = orc.sweep(a, T0, threads=1)
    ib, ab, sb = orc.sweep(b, T0, threads=7)
    assert sa == sb
    np.testing.assert_array_equal(ia, ib)
    np.testing.assert_array_equal(aa, ab)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])


@pytest.mark.parametrize("threads", [2, 3, 16, 200])
@pytest.mark.parametrize("n", [1, 5, 4097])
def test_sweep_mt_thread_counts_and_tiny_inputs(orc, gen, n, threads):
    """More threads than records, thread counts that grow and shrink between calls (the
    workers are parked and reused), closed loop, a short output buffer."""
    a = gen.fill(55, 5, 0, n, T0, orc.load().orc_classify)
    b = {k: v.copy() for k, v in a.items()}
    for k in range(3):
        ia, aa, sa = orc.sweep(a, T0 + k, mode=1, seed=5, threads=1)
        ib, ab, sb = orc.sweep(b, T0 + k, mode=1, seed=5, threads=threads)
        assert sa == sb
        np.testing.assert_array_equal(ia, ib)
        np.testing.assert_array_equal(aa, ab)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])
    # short caller buffers: E_NOSPACE, the first `cap` entries valid, the count is the need
    import ctypes as C
    c = gen.fill(1, 1, 0, 4097, T0, orc.load().orc_classify)
    want_idx, _, want_st = orc.sweep({k: v.copy() for k, v in c.items()}, T0, threads=1)
    cap = 10
    idx = np.zeros(cap, np.uint64)
    act = np.zeros(cap, np.uint32)
    cnt, st = C.c_uint64(0), orc.OrcStats()
    cs = orc.cols_struct(c)
    rc = orc.load().orc_sweep_mt(C.byref(cs), 4097, 0, T0, 0, 0, idx.ctypes.data, act.ctypes.data, cap,
                                 C.byref(cnt), C.byref(st), threads)
    assert rc == -3 and cnt.value == len(want_idx) > cap
    np.testing.assert_array_equal(idx, want_idx[:cap])
    assert st.as_dict() == want_st


def test_sweep_mt_concurrent_callers_are_serialised(orc, gen):
    import threading
    n = 20_000
    base = gen.fill(2, 2, 0, n, T0, orc.load().orc_classify)
    want = orc.sweep({k: v.copy() for k, v in base.items()}, T0, threads=1)
    out = [None] * 4

    def run(j):
        out[j] = orc.sweep({k: v.copy() for k, v in base.items()}, T0, threads=3 + j)

    th = [threading.Thread(target=run, args=(j,)) for j in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=120)
        assert not t.is_alive()
    for o in out:
        np.testing.assert_array_equal(o[0], want[0])
        np.testing.assert_array_equal(o[1], want[1])
        assert o[2] == want[2]


def test_closed_loop_oracles_agree(orc, opy, gen):
    n, seed = 1500, 5
    cols = gen.fill(55, seed, 0, n, T0, orc.load().orc_classify)
    recs = [opy.Record(**{name: int(cols[name][i]) for name, _ in orc.COLUMNS}) for i in range(n)]
    for k in range(70):
        due_py, st_py = opy.sweep(recs, T0 - 5 + k, mode=1, seed=seed)
        idx, act, st_c = orc.sweep(cols, T0 - 5 + k, mode=1, seed=seed)
        assert [(int(a), int(b)) for a, b in zip(idx, act)] == due_py, k
        assert st_c["n_result_ok"] == st_py.n_result_ok and st_c["n_remedy_fail"] == st_py.n_remedy_fail
    for i in range(n):
        for name, _ in orc.COLUMNS:
            assert int(cols[name][i]) == getattr(recs[i], name), (i, name)
