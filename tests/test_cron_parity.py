"""Host-side parity (no GPU): the PRODUCT's upsert-time code — am_cron_parse,
am_cron_matches/next, am_healthcheck_classify, am_civil_from_unix — against the
CPU oracle, bit for bit, on fixed vectors, random byte strings and the synthetic
populations at scale.  Mirrors the reference's own unit tests for this path
(healthcheck_controller_unit_test.go:617-660)."""
import ctypes as C
import random

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from test_oracle_golden import (ACTIVATION, APPENDIX_C, EVERY, NEXT, REJECTED, REJECTED_ZONES, T0, TOKENS, UNSUPPORTED,
                                utc)


def both(am, orc, spec):
    raw = spec if isinstance(spec, bytes) else spec.encode("utf-8", "surrogateescape")
    o_rc, o, _ = orc.cron_parse(raw)
    p = am.AmCron()
    err = C.create_string_buffer(200)
    p_rc = am.load().am_cron_parse(raw, len(raw), C.byref(p), err, len(err))
    return o_rc, o, p_rc, p, err.value


def same(o, p):
    return bytes(o) == bytes(p)  # whole struct, byte for byte (layouts are identical)


def test_TestProcessHealthCheck_InvalidCron_ReturnsError(am):
    """reference: healthcheck_controller_unit_test.go:617-634"""
    with pytest.raises(am.CronParseError):
        am.cron_parse("NOT_A_VALID_CRON")
    rc, rec = am.classify(repeat_after_sec=0, cron="NOT_A_VALID_CRON")
    assert rc == 0 and rec["flags"][0] & 7 == am.KIND_PARSE_ERROR


def test_TestProcessHealthCheck_ValidCron_SetsRepeatAfterSec(am):
    """reference: healthcheck_controller_unit_test.go:636-660"""
    c = am.cron_parse("@every 5s")
    assert c.kind == am.AM_CRON_EVERY and c.repeat_after_sec(T0) == 5 > 0
    rc, rec = am.classify(repeat_after_sec=0, cron="@every 5s")
    assert rc == 0 and rec["flags"][0] & 7 == am.KIND_CRON_EVERY and rec["ras"][0] == 5


def test_TestRemedyWorkflow_IsEmpty(am):
    """reference: api/v1alpha1/healthcheck_types_unit_test.go:24-38"""
    assert am.remedy_is_empty("", True, 0, True)
    assert not am.remedy_is_empty("remedy-", True, 0, True)
    assert not am.remedy_is_empty("", False, 0, True)
    assert not am.remedy_is_empty("", True, 30, True)
    assert not am.remedy_is_empty("", True, 0, False)


@pytest.mark.parametrize("spec", [s for s, _ in APPENDIX_C] + REJECTED + REJECTED_ZONES + UNSUPPORTED + [s for s, _ in EVERY])
def test_fixed_vectors_product_equals_oracle(am, orc, spec):
    o_rc, o, p_rc, p, _ = both(am, orc, spec)
    assert o_rc == p_rc, spec
    assert same(o, p), spec


@pytest.mark.parametrize("spec,T,want", ACTIVATION)
def test_activation_product(am, spec, T, want):
    c = am.cron_parse(spec)
    assert c.matches(T) is want
    assert (c.next(T - 1) == T) is want


@pytest.mark.parametrize("spec,T,want", NEXT)
def test_next_product(am, spec, T, want):
    assert am.cron_parse(spec).next(T) == want


# Go's unicode.IsSpace set, spelled with escapes: NBSP, NEL, ideographic space ...
SEPARATORS = [" ", "  ", "\t", " \t ", "\u00a0", "\u0085", "\u3000 ", "\u2003", "\u200b", "\u1680", "\x0b", "\x1f"]


@settings(max_examples=800, deadline=None)
@given(st.lists(st.sampled_from(TOKENS), min_size=4, max_size=6), st.sampled_from(SEPARATORS),
       st.sampled_from(["", "", "", "TZ=UTC ", "CRON_TZ=UTC  ", "TZ=Mars/Base "]))
def test_token_specs(am, orc, toks, sep, prefix):
    spec = prefix + sep.join(toks)
    o_rc, o, p_rc, p, _ = both(am, orc, spec)
    assert o_rc == p_rc and same(o, p), spec


@settings(max_examples=800, deadline=None)
@given(st.binary(min_size=0, max_size=28))
def test_arbitrary_bytes(am, orc, raw):
    """including invalid UTF-8: Go decodes bad bytes as U+FFFD width 1"""
    o_rc, o, p_rc, p, _ = both(am, orc, raw)
    assert o_rc == p_rc and same(o, p), raw


@settings(max_examples=800, deadline=None)
@given(st.text(alphabet="0123456789*/-,? \t@everyjanfmsuTZ=UTC+.hms\u00b5\u0130K\u00a0", min_size=0,
               max_size=26))
def test_random_text(am, orc, spec):
    o_rc, o, p_rc, p, _ = both(am, orc, spec)
    assert o_rc == p_rc and same(o, p), repr(spec)


def test_unicode_name_folding(am, orc, opy):
    """strings.ToLower maps U+0130 to 'i': "FR\\u0130" is Friday for robfig."""
    spec = "0 0 * * FR\u0130"
    o_rc, o, p_rc, p, _ = both(am, orc, spec)
    assert o_rc == p_rc == 0 and same(o, p) and o.dow == 1 << 5
    assert opy.cron_parse(spec).dow == 1 << 5


def test_next_matches_product_equals_oracle_random(am, orc):
    rng = random.Random(99)
    lib, olib = am.load(), orc.load()
    fields = [(0, 59), (0, 23), (1, 31), (1, 12), (0, 6)]
    for _ in range(1500):
        parts = []
        for lo, hi in fields:
            k = rng.random()
            if k < 0.4:
                parts.append(rng.choice(["*", "?", "*/2", "*/5", "*/7"]))
            elif k < 0.7:
                parts.append(str(rng.randint(lo, hi)))
            else:
                a = rng.randint(lo, hi)
                parts.append(f"{a}-{rng.randint(a, hi)}/{rng.randint(1, 9)}")
        spec = " ".join(parts)
        o_rc, o, p_rc, p, _ = both(am, orc, spec)
        assert o_rc == p_rc == 0 and same(o, p)
        for _ in range(3):
            T = rng.randint(utc(1999, 1, 1), utc(2090, 1, 1))
            assert lib.am_cron_matches(C.byref(p), T) == olib.orc_cron_matches(C.byref(o), T)
            T60 = T - T % 60
            assert lib.am_cron_matches(C.byref(p), T60) == olib.orc_cron_matches(C.byref(o), T60)
            assert lib.am_cron_next(C.byref(p), T) == olib.orc_cron_next(C.byref(o), T), (spec, T)
            assert lib.am_cron_repeat_after_sec(C.byref(p), T) == olib.orc_cron_repeat_after_sec(C.byref(o), T)


def test_civil_time_product_equals_oracle(am, orc):
    rng = random.Random(3)
    a, b = (C.c_int32 * 6)(), (C.c_int32 * 6)()
    Ts = [0, -1, 59, 60, 86399, 86400, -86400, -86401, 951782400, 1709164800, T0, 4102444800,
          -2208988800, 253402300799, -62135596800] + [rng.randint(-2**42, 2**42) for _ in range(20000)]
    for T in Ts:
        am.load().am_civil_from_unix(T, C.byref(a))
        orc.load().orc_civil_from_unix(T, C.byref(b))
        assert tuple(a) == tuple(b), T


@pytest.mark.parametrize("config,seed,n", [(1, 1, 1000), (11, 1, 1000), (2, 2, 400_000), (3, 3, 400_000),
                                           (55, 5, 100_000)])
def test_populations_classify_identically(am, orc, gen, config, seed, n):
    """every record of the synthetic populations: product ladder/parser == oracle's"""
    p = gen.fill(config, seed, 0, n, T0, am.load().am_healthcheck_classify)
    o = gen.fill(config, seed, 0, n, T0, orc.load().orc_classify)
    for name in am.COLUMN_NAMES:
        np.testing.assert_array_equal(p[name], o[name], err_msg=name)
    if config == 2:
        kinds = np.bincount(p["flags"] & 7, minlength=7) / n
        assert abs(kinds[am.KIND_INTERVAL] - 0.50) < 0.01 and abs(kinds[am.KIND_CRON_SPEC] - 0.40) < 0.01
        assert abs(kinds[am.KIND_CRON_EVERY] - 0.08) < 0.01 and kinds[am.KIND_STOPPED] > 0.005
        assert kinds[am.KIND_PARSE_ERROR] > 0.003 and kinds[am.KIND_NO_RESOURCE] > 0.003


def test_classify_domain_checks(am, orc):
    big = 1 << 31
    for kw in [dict(repeat_after_sec=big), dict(remedy_runs_limit=big, repeat_after_sec=5),
               dict(success_count=-big - 1, repeat_after_sec=5), dict(finished_at=1 << 55, repeat_after_sec=5),
               dict(remedy_finished_at=0, repeat_after_sec=5), dict(cron="@every 2540400h")]:
        rc, _ = am.classify(**kw)
        assert rc == am.AM_E_RANGE, kw
    rc, rec = am.classify(cron="CRON_TZ=Europe/Paris 0 9 * * *")
    assert rc == 0 and rec["flags"][0] & 7 == am.KIND_CRON_SPEC
    assert rec["flags"][0] >> am.F_TZ_SHIFT == am.tz_lookup("Europe/Paris") > 0
    rc, rec = am.classify(cron="CRON_TZ=Nowhere/Land 0 9 * * *")  # time.LoadLocation fails: a parse error
    assert rc == 0 and rec["flags"][0] & 7 == am.KIND_PARSE_ERROR
    rc, rec = am.classify(repeat_after_sec=-7, cron="")
    assert rc == 0 and rec["flags"][0] & 7 == am.KIND_STOPPED
    rc, rec = am.classify(repeat_after_sec=60, cron="NOT_A_VALID_CRON", has_remedy=True, fail_p8=77)
    assert rec["flags"][0] == am.KIND_INTERVAL | am.F_HAS_REMEDY | am.F_TIMER_ARMED | (77 << 16) and rec["ras"][0] == 60
