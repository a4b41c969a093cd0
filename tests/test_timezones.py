"""Named time zones of cron schedules ("CRON_TZ=Europe/Paris 0 9 * * *"; robfig parser.go →
time.LoadLocation, call site hcc.go:253).  Three implementations that share nothing:

  product   csrc/tz.cpp reads the system's TZif files itself (transition table + POSIX TZ footer);
            the per-tick zone table is computed on the device from a flattened copy (tz_eval.h)
  C oracle  libc: setenv("TZ") + tzset() + localtime_r / mktime
  Py oracle the standard library's zoneinfo

CPU tier: offsets, activation (matches) and Next() agree between the three on many zones and
instants, daylight-saving gaps and overlaps included.  The device path is covered by the `-m gpu`
tests below (they also run on the CPU emulator through tests/test_c_abi_on_emulator.py)."""
import ctypes as C
import datetime as dt
import random
import zoneinfo

import numpy as np
import pytest

ZONES = ["America/New_York", "Europe/Paris", "Asia/Kolkata", "Asia/Kathmandu", "Australia/Lord_Howe",
         "America/St_Johns", "America/Sao_Paulo", "Pacific/Chatham", "Africa/Casablanca", "Europe/Dublin",
         "Antarctica/Troll", "Asia/Tehran", "Pacific/Apia", "America/Nuuk", "Europe/London", "Asia/Tokyo"]


def utc(y, m, d, hh=0, mm=0, ss=0):
    return int(dt.datetime(y, m, d, hh, mm, ss, tzinfo=dt.timezone.utc).timestamp())


def _instants(seed):
    rng = random.Random(seed)
    ts = [rng.randrange(0, 4102444800) for _ in range(1500)]           # 1970 .. 2100
    ts += [utc(2026, 9, 21) + 86400 * k + 3600 * (k % 24) for k in range(0, 800, 3)]
    for y in range(2024, 2041):                                          # dense around the usual transition months
        for m in (3, 4, 9, 10, 11):
            base = utc(y, m, 1)
            ts += [base + 3600 * k for k in range(0, 24 * 31, 7)]
    return ts


@pytest.mark.parametrize("zone", ZONES)
def test_utc_offsets_product_vs_libc_vs_zoneinfo(am, orc, opy, zone):
    zi = zoneinfo.ZoneInfo(zone)
    pid = am.tz_lookup(zone)
    oid = C.c_int32()
    assert orc.load().orc_tz_lookup(zone.encode(), len(zone), C.byref(oid)) == 0
    yid = opy.tz_lookup(zone)
    assert pid > 0 and oid.value > 0 and yid > 0
    off = C.c_int32()
    for t in _instants(hash(zone) & 0xFFFF):
        want = int(dt.datetime.fromtimestamp(t, zi).utcoffset().total_seconds())
        assert am.tz_offset(pid, t) == want, (zone, t)
        assert opy.tz_offset(yid, t) == want
        # glibc disagrees with Go / zoneinfo about times BEFORE a zone's first recorded transition
        # (local mean time vs the first standard time): the oracle is only consulted from 1970 on
        assert orc.load().orc_tz_offset(oid.value, t, C.byref(off)) == 0 and off.value == want, (zone, t)


def test_zone_names_as_go_validates_them(am):
    assert am.tz_lookup("") == am.tz_lookup("UTC") == am.tz_lookup("Local") == 0
    for bad in ("Nowhere/Land", "../zoneinfo/UTC", "/etc/localtime", "Europe/Paris/..", "europe/paris_"):
        with pytest.raises(am.CronParseError):
            am.tz_lookup(bad)
    assert am.tz_lookup("Europe/Paris") == am.tz_lookup("Europe/Paris") > 0  # registered once


KAT = [  # spec, instant, fires?
    ("CRON_TZ=Asia/Kolkata 30 9 * * *", utc(2026, 9, 21, 4, 0), True),       # 09:30 IST = 04:00 UTC
    ("CRON_TZ=Asia/Kolkata 30 9 * * *", utc(2026, 9, 21, 9, 30), False),
    ("TZ=Asia/Kathmandu 0 0 * * *", utc(2026, 9, 20, 18, 15), True),          # +05:45
    ("CRON_TZ=America/New_York 30 1 * * *", utc(2026, 11, 1, 5, 30), True),    # 01:30 EDT ...
    ("CRON_TZ=America/New_York 30 1 * * *", utc(2026, 11, 1, 6, 30), True),    # ... and 01:30 EST: fires twice
    ("CRON_TZ=America/New_York 0 2 * * *", utc(2026, 11, 1, 6, 0), False),     # 06:00 UTC is 01:00 EST
    ("CRON_TZ=America/New_York 0 2 * * *", utc(2026, 11, 1, 7, 0), True),
    ("CRON_TZ=America/New_York 30 2 * * *", utc(2027, 3, 14, 7, 30), False),   # 02:30 does not exist that day
    ("CRON_TZ=America/New_York 30 3 * * *", utc(2027, 3, 14, 7, 30), True),    # 03:30 EDT
    ("TZ=Australia/Lord_Howe 15 2 * * *", utc(2026, 10, 3, 15, 45), False),    # 02:00 -> 02:30 that night: no 02:15
    ("TZ=Australia/Lord_Howe 45 2 * * *", utc(2026, 10, 3, 15, 45), True),     # 02:45 +11:00 (a 30-minute shift)
    ("TZ=Australia/Lord_Howe 15 2 * * *", utc(2026, 10, 4, 15, 15), True),     # 02:15 +11:00 the day after
    ("CRON_TZ=Europe/Paris @daily", utc(2026, 9, 20, 22, 0), True),            # midnight CEST
    ("CRON_TZ=Europe/Paris 0 0 * * MON", utc(2026, 9, 20, 22, 0), True),       # Monday in Paris, Sunday in UTC
    ("CRON_TZ=Europe/Paris 0 0 * * SUN", utc(2026, 9, 20, 22, 0), False),
    ("CRON_TZ=UTC 15 9 * * *", utc(2026, 9, 21, 9, 15), True),
]


@pytest.mark.parametrize("spec,T,want", KAT)
def test_activation_in_a_zone(am, orc, opy, spec, T, want):
    assert am.cron_parse(spec).matches(T) is want
    rc, c, _ = orc.cron_parse(spec)
    assert rc == 0 and bool(orc.load().orc_cron_matches(C.byref(c), T)) is want
    assert opy.cron_matches(opy.cron_parse(spec), T) is want
    # matches(T) <=> Next(T - 1) == T, in all three
    assert (am.cron_parse(spec).next(T - 1) == T) is want
    assert (orc.load().orc_cron_next(C.byref(c), T - 1) == T) is want
    assert (opy.cron_next(opy.cron_parse(spec), T - 1) == T) is want


def _random_spec(rng):
    def f(lo, hi):
        k = rng.randrange(6)
        if k < 2:
            return "*"
        if k == 2:
            return f"*/{rng.randrange(2, 8)}"
        if k == 3:
            return str(rng.randrange(lo, hi + 1))
        a = rng.randrange(lo, hi + 1)
        b = rng.randrange(a, hi + 1)
        return f"{a}-{b}" if k == 4 else f"{a},{b}"
    return " ".join([f(0, 59), f(0, 23), f(1, 31), f(1, 12), f(0, 6)])


def test_next_in_a_zone_product_vs_robfig_walk_vs_brute_force(am, orc, opy):
    """am_cron_next (first instant whose wall clock matches) == the oracle's restatement of robfig's
    field-increment walk on the zone's wall clock (libc mktime / localtime) == zoneinfo brute force,
    for random specs and instants — including starts inside the spring / autumn transition days."""
    rng = random.Random(7)
    lib = orc.load()
    starts = [utc(2026, 9, 21, 9, 15), utc(2026, 10, 31, 20, 0), utc(2027, 3, 13, 22, 0), utc(2027, 3, 28, 0, 30),
              utc(2026, 12, 31, 23, 59, 59), utc(2028, 2, 28, 12, 0)]
    n = 0
    for zone in ZONES[:8]:
        for _ in range(40):
            spec = f"CRON_TZ={zone} {_random_spec(rng)}"
            pc = am.cron_parse(spec)
            rc, oc, _ = orc.cron_parse(spec)
            yc = opy.cron_parse(spec)
            assert rc == 0
            t = rng.choice(starts) + rng.randrange(0, 86400)
            got = pc.next(t)
            w_c = lib.orc_cron_next(C.byref(oc), t)
            w_c = None if w_c == -(1 << 63) else w_c
            w_py = opy.cron_next(yc, t)
            assert got == w_py, (spec, t, got, w_py)
            # robfig's walk adds ABSOLUTE hours: after a daylight-saving shift that is not a whole hour
            # (Lord Howe: 30 minutes) it sits at hh:30 and returns 14:30 where 14:00 matched — a quirk
            # of Next() when walking across such a day, not of matches(T) (Next(13:59:59) is 14:00:00
            # there too).  The product defines Next as "first instant whose wall clock matches".
            if zone != "Australia/Lord_Howe":
                assert got == w_c, (spec, t, got, w_c)
            if got is not None:
                assert got > t and pc.matches(got)
                n += 1
    assert n > 200


# ---------------------------------------------------------------- device path
T0 = 1789982100


@pytest.mark.gpu
def test_zone_bound_checks_fire_on_their_zones_wall_clock(am, orc):
    """One record per KAT row, ticked at the row's instant: the device evaluates matches(T) against the
    zone offsets the launcher keeps on the device, and equals the oracle (libc)."""
    recs = []
    for spec, _, _ in KAT:
        rc, r = am.classify(cron=spec, finished_at=T0 - 5)
        assert rc == 0 and r["flags"][0] & 7 == am.KIND_CRON_SPEC
        recs.append(r)
    cols = am.records_to_columns(np.concatenate(recs))
    ocols = {k: v.copy() for k, v in cols.items()}
    # the oracle needs the same zone ids: introduce the zones in the product's order
    for spec, _, _ in KAT:
        assert orc.cron_parse(spec)[0] == 0
    for k, (spec, _, _) in enumerate(KAT):
        rc, oc, _ = orc.cron_parse(spec)
        ocols["flags"][k] = (int(ocols["flags"][k]) & 0x00FFFFFF) | (oc.tz_id << 24)
    with am.Sweep(capacity=len(KAT)) as s:
        s.load_range(0, cols)
        for k, (spec, T, want) in enumerate(KAT):
            idx, act, st = s.tick(T)
            wi, wa, ws = orc.sweep(ocols, T)
            fired = k in idx.tolist()
            assert fired is want, (spec, T)
            assert sorted(idx.tolist()) == sorted(wi.tolist()), (spec, T)
        # off the minute nothing bound to a zone with a whole-minute offset can fire (masks skipped)
        idx, _, _ = s.tick(KAT[0][1] + 1)
        assert len(idx) == 0


@pytest.mark.gpu
def test_population_with_zones_equals_oracle_at_local_midnights(am, orc, gen):
    """config 22 (config 2's population with 1.5 % of the 5-field specs bound to one of six zones) at instants
    chosen as local midnights / DST edges of those zones: list, actions, statistics and columns == oracle."""
    n = 60_000
    prod = gen.fill(22, 2, 0, n, T0, am.load().am_healthcheck_classify)
    orac = gen.fill(22, 2, 0, n, T0, orc.load().orc_classify)
    assert ((prod["flags"] >> 24) != 0).sum() > 100
    for name in am.COLUMN_NAMES:
        np.testing.assert_array_equal(prod[name], orac[name], err_msg=name)
    instants = [utc(2026, 9, 21, 18, 30), utc(2026, 9, 21, 18, 15), utc(2026, 9, 21, 22, 0), utc(2026, 9, 22, 4, 0),
                utc(2026, 9, 22, 2, 30), utc(2026, 9, 21, 13, 30), utc(2026, 11, 1, 6, 0), utc(2026, 10, 3, 16, 0),
                utc(2027, 3, 14, 7, 0)]
    with am.Sweep(capacity=n) as s:
        s.load_range(0, prod)
        for T in instants:
            gi, ga, gs = s.tick(T)
            wi, wa, ws = orc.sweep(orac, T)
            assert gs == ws, T
            np.testing.assert_array_equal(gi, wi, err_msg=str(T))
            np.testing.assert_array_equal(ga, wa)
        dev = s.read_range(0, n)
        for name in am.COLUMN_NAMES:
            np.testing.assert_array_equal(dev[name], orac[name], err_msg=name)


@pytest.mark.gpu
def test_consecutive_ticks_across_daylight_transitions(am, orc):
    """The device keeps one UTC offset per zone and the launcher refreshes them only when a tick leaves the
    window in which no zone changes its offset: tick second by second (and in jumps) across the 2026/2027
    transitions of three zones — among them Lord Howe's half-hour shift — with every-minute schedules bound
    to them; each tick's list must equal the oracle's (libc zones), in particular in the second of the
    transition itself and the one before it."""
    zones = ["America/New_York", "Europe/Paris", "Australia/Lord_Howe"]
    specs = [f"CRON_TZ={z} {m} {h} * * *" for z in zones for h in (1, 2, 3) for m in (0, 30)]
    specs += ["0 6 * * *", "0 1 * * *"]  # plain UTC neighbours
    recs = [am.classify(cron=sp, finished_at=T0 - 5)[1] for sp in specs]
    cols = am.records_to_columns(np.concatenate(recs))
    ocols = {k: v.copy() for k, v in cols.items()}
    for k, sp in enumerate(specs):
        rc, oc, _ = orc.cron_parse(sp)
        assert rc == 0
        ocols["flags"][k] = (int(ocols["flags"][k]) & 0x00FFFFFF) | (oc.tz_id << 24)
    edges = [utc(2026, 11, 1, 6, 0), utc(2027, 3, 14, 7, 0),      # New York: EDT -> EST, EST -> EDT
             utc(2026, 10, 25, 1, 0), utc(2027, 3, 28, 1, 0),     # Paris: CEST -> CET, CET -> CEST
             utc(2026, 10, 3, 15, 30), utc(2027, 4, 3, 15, 0)]    # Lord Howe: +10:30 -> +11, +11 -> +10:30
    fired = 0
    with am.Sweep(capacity=len(specs)) as s:
        s.load_range(0, cols)
        for e in edges:
            # hours around the edge on the minute (the schedules can fire there), then the edge second by second
            ticks = [e - 3600, e - 1800, e - 60] + list(range(e - 3, e + 4)) + [e + 60, e + 1800, e + 3600, e - 7200, e]
            for T in ticks:
                gi, ga, _ = s.tick(T, mode=am.SWEEP_FULL_SCAN)
                wi, wa, _ = orc.sweep(ocols, T, am.SWEEP_FULL_SCAN)
                np.testing.assert_array_equal(gi, wi, err_msg=f"edge {e} tick {T}")
                np.testing.assert_array_equal(ga, wa)
                fired += len(gi)
    assert fired > 20


@pytest.mark.gpu
@pytest.mark.parametrize("edge", ["new_york_fall", "lord_howe_spring"])
def test_blocked_run_of_ticks_across_a_daylight_transition(am, orc, edge):
    """am_sweep_run_ticks with AM_SWEEP_BLOCKED holds one UTC offset per zone for a whole block of ticks:
    the launcher must cut the blocks at the instant a zone changes its offset.  Per-tick statistics over
    200 consecutive seconds around a transition == the oracle (libc zones) tick by tick."""
    e = {"new_york_fall": utc(2026, 11, 1, 6, 0), "lord_howe_spring": utc(2026, 10, 3, 15, 30)}[edge]
    zones = ["America/New_York", "Europe/Paris", "Australia/Lord_Howe", "Asia/Kathmandu"]
    specs = [f"CRON_TZ={z} * * * * *" for z in zones] + [f"CRON_TZ={z} {m} {h} * * *" for z in zones for h in (1, 2) for m in (0, 30)]
    specs += ["* * * * *", "0 6 * * *"]
    recs = [am.classify(cron=sp, finished_at=T0 - 5)[1] for sp in specs]
    cols = am.records_to_columns(np.concatenate(recs))
    ocols = {k: v.copy() for k, v in cols.items()}
    for k, sp in enumerate(specs):
        rc, oc, _ = orc.cron_parse(sp)
        assert rc == 0
        ocols["flags"][k] = (int(ocols["flags"][k]) & 0x00FFFFFF) | (oc.tz_id << 24)
    start, nt = e - 130, 260
    with am.Sweep(capacity=len(specs)) as s:
        s.load_range(0, cols)
        stats = s.run_ticks(start, nt, mode=am.SWEEP_BLOCKED)
        total = 0
        for k in range(nt):
            _, _, ws = orc.sweep(ocols, start + k)
            gs = {f: int(stats[f][k]) for f in am.abi.STAT_FIELDS}
            assert gs == ws, f"{edge}: tick {start + k} (edge {e:+d})"
            total += ws["n_emitted"]
        assert total > 20
