"""oracle_py.py — CPU ORACLE #2 (independent Python-stdlib restatement).

TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing in the product).

Second, independently structured restatement of the same reference decisions
as oracle/amsweep_oracle.c, used to cross-check it:

  * robfig/cron v3.0.1 ``ParseStandard`` (go.mod:14; call site
    internal/controllers/healthcheck_controller.go:253) — here fields are
    expanded to Python ``set``s of ints and only then folded into masks;
  * ``SpecSchedule.Next`` (call site hcc.go:262) — here a plain forward scan
    over calendar days with ``datetime`` (NOT the Go field-increment loop), so
    an error in the C port of that loop cannot hide;
  * the ladder hcc.go:225-267, the result/remedy state machine hcc.go:633-724
    and :819-852, and the re-arm rule hcc.go:745-752.

PARITY UNPINNED against the Go binary (no Go toolchain, robfig not vendored):
see the header of oracle/amsweep_oracle.h for exactly which claims the
reference's own tests pin.
"""
from __future__ import annotations

import datetime as _dt
from dataclasses import dataclass, field as _field

STAR_BIT = 1 << 63
U64 = (1 << 64) - 1

CRON_ERROR, CRON_SPEC, CRON_EVERY = 0, 1, 2

KIND_NO_RESOURCE, KIND_STOPPED, KIND_INTERVAL, KIND_CRON_SPEC = 0, 1, 2, 3
KIND_CRON_EVERY, KIND_PARSE_ERROR, KIND_HOST_FALLBACK = 4, 5, 6
F_HAS_REMEDY = 1 << 3
F_PENDING_OK = 1 << 4
F_PENDING_FAIL = 1 << 5
F_REMEDY_PENDING = 1 << 6
F_REMEDY_OUTCOME_OK = 1 << 7
F_TOMBSTONE = 1 << 8
F_STOPPED_REPORTED = 1 << 9
F_TIMER_ARMED = 1 << 10  # r.GetTimerByName(name) != nil, hcc.go:264
F_FAILP_SHIFT = 16

ACT_SUBMIT_HC = 0x01
ACT_RUN_REMEDY = 0x02
ACT_STOPPED = 0x04
ACT_PARSE_ERROR = 0x08
ACT_REMEDY_SKIP = 0x10
ACT_RESET_ON_PASS = 0x20
ACT_RESET_ON_INTERVAL = 0x40
ACT_ANOMALY = 0x80

MODE_CLOSED_LOOP = 0x1

E_RANGE, E_PARSE, E_UNSUPPORTED = -2, -6, -7


class CronError(ValueError):
    """robfig would return an error (hcc.go:254-257)."""


class CronUnsupported(ValueError):
    """Valid for robfig but not evaluated on the device path (named zone)."""


# Go's unicode.IsSpace
_GO_SPACE = {0x09, 0x0A, 0x0B, 0x0C, 0x0D, 0x20, 0x85, 0xA0, 0x1680, 0x2028, 0x2029, 0x202F,
             0x205F, 0x3000} | set(range(0x2000, 0x200B))


def _is_space(ch: str) -> bool:
    return ord(ch) in _GO_SPACE


def go_fields(s: str) -> list[str]:
    out, cur = [], []
    for ch in s:
        if _is_space(ch):
            if cur:
                out.append("".join(cur))
                cur = []
        else:
            cur.append(ch)
    if cur:
        out.append("".join(cur))
    return out


def go_trim_space(s: str) -> str:
    b, e = 0, len(s)
    while b < e and _is_space(s[b]):
        b += 1
    while e > b and _is_space(s[e - 1]):
        e -= 1
    return s[b:e]


def go_atoi(s: str) -> int:
    """strconv.Atoi acceptance: [+-]?[0-9]+ (ASCII digits only), int64 range."""
    body = s[1:] if s[:1] in ("+", "-") else s
    if not body or any(c not in "0123456789" for c in body):
        raise CronError(f'strconv.Atoi: parsing "{s}": invalid syntax')
    v = int(body)
    if s[:1] == "-":
        v = -v
    if not (-(1 << 63) <= v <= (1 << 63) - 1):
        raise CronError(f'strconv.Atoi: parsing "{s}": value out of range')
    return v


_MONTHS = {n: i + 1 for i, n in enumerate(
    ["jan", "feb", "mar", "apr", "may", "jun", "jul", "aug", "sep", "oct", "nov", "dec"])}
_DOWS = {n: i for i, n in enumerate(["sun", "mon", "tue", "wed", "thu", "fri", "sat"])}

# (min, max, names)
_BOUNDS = {
    "minute": (0, 59, None),
    "hour": (0, 23, None),
    "dom": (1, 31, None),
    "month": (1, 12, _MONTHS),
    "dow": (0, 6, _DOWS),
}


def _go_lower(s: str) -> str:
    # strings.ToLower uses simple case mapping per rune; Python's str.lower()
    # applies full mappings (U+0130 -> 'i' + U+0307), so map rune by rune.
    out = []
    for ch in s:
        if ch == "İ":
            out.append("i")
        else:
            low = ch.lower()
            out.append(low if len(low) == 1 else ch)
    return "".join(out)


def _int_or_name(tok: str, names) -> int:
    if names is not None:
        hit = names.get(_go_lower(tok))
        if hit is not None:
            return hit
    v = go_atoi(tok)
    if v < 0:
        raise CronError(f"negative number ({v}) not allowed: {tok}")
    return v


def _expand_range(expr: str, lo_b: int, hi_b: int, names) -> tuple[set[int], bool]:
    """One comma piece -> (values, star?).  Mirrors getRange's quirks:
    '*-5' is '*'; 'N/s' means 'N-max/s'; step>1 drops the star."""
    parts = expr.split("/")
    lh = parts[0].split("-")
    star = False
    if lh[0] in ("*", "?"):
        start, end, star = lo_b, hi_b, True
    else:
        start = _int_or_name(lh[0], names)
        if len(lh) == 1:
            end = start
        elif len(lh) == 2:
            end = _int_or_name(lh[1], names)
        else:
            raise CronError(f"too many hyphens: {expr}")
    if len(parts) == 1:
        step = 1
    elif len(parts) == 2:
        step = go_atoi(parts[1])
        if step < 0:
            raise CronError(f"negative number ({step}) not allowed: {parts[1]}")
        if len(lh) == 1:
            end = hi_b
        if step > 1:
            star = False
    else:
        raise CronError(f"too many slashes: {expr}")
    if start < lo_b:
        raise CronError(f"beginning of range ({start}) below minimum ({lo_b}): {expr}")
    if end > hi_b:
        raise CronError(f"end of range ({end}) above maximum ({hi_b}): {expr}")
    if start > end:
        raise CronError(f"beginning of range ({start}) beyond end of range ({end}): {expr}")
    if step == 0:
        raise CronError(f"step of range should be a positive number: {expr}")
    return set(range(start, end + 1, step)), star


def _field_mask(text: str, which: str) -> int:
    lo_b, hi_b, names = _BOUNDS[which]
    values: set[int] = set()
    star = False
    for piece in text.split(","):
        if piece == "":
            continue  # strings.FieldsFunc drops empty pieces
        v, s = _expand_range(piece, lo_b, hi_b, names)
        values |= v
        star = star or s
    m = 0
    for v in values:
        m |= 1 << v
    return m | (STAR_BIT if star else 0)


def _all(which: str) -> int:
    lo_b, hi_b, _ = _BOUNDS[which]
    return sum(1 << v for v in range(lo_b, hi_b + 1)) | STAR_BIT


_UNITS = {"ns": 1, "us": 10**3, "µs": 10**3, "μs": 10**3, "ms": 10**6, "s": 10**9,
          "m": 60 * 10**9, "h": 3600 * 10**9}


def parse_duration(s: str) -> int:
    """time.ParseDuration -> nanoseconds (raises CronError)."""
    orig = s
    neg = False
    if s and s[0] in "+-":
        neg = s[0] == "-"
        s = s[1:]
    if s == "0":
        return 0
    if s == "":
        raise CronError(f"time: invalid duration {orig!r}")
    total = 0
    while s:
        if not (s[0] == "." or s[0] in "0123456789"):
            raise CronError(f"time: invalid duration {orig!r}")
        i = 0
        v = 0
        while i < len(s) and s[i] in "0123456789":
            if v > (1 << 63) // 10:
                raise CronError(f"time: invalid duration {orig!r}")
            v = v * 10 + int(s[i])
            if v > (1 << 63):
                raise CronError(f"time: invalid duration {orig!r}")
            i += 1
        pre = i > 0
        s = s[i:]
        post = False
        f, scale = 0, 1.0
        if s and s[0] == ".":
            s = s[1:]
            i = 0
            overflow = False
            while i < len(s) and s[i] in "0123456789":
                if not overflow:
                    if f > ((1 << 63) - 1) // 10:
                        overflow = True
                    else:
                        y = f * 10 + int(s[i])
                        if y > (1 << 63):
                            overflow = True
                        else:
                            f = y
                            scale *= 10
                i += 1
            post = i > 0
            s = s[i:]
        if not pre and not post:
            raise CronError(f"time: invalid duration {orig!r}")
        i = 0
        while i < len(s) and not (s[i] == "." or s[i] in "0123456789"):
            i += 1
        if i == 0:
            raise CronError(f"time: missing unit in duration {orig!r}")
        u, s = s[:i], s[i:]
        if u not in _UNITS:
            raise CronError(f"time: unknown unit {u!r} in duration {orig!r}")
        unit = _UNITS[u]
        if v > (1 << 63) // unit:
            raise CronError(f"time: invalid duration {orig!r}")
        v *= unit
        if f > 0:
            v += int(float(f) * (float(unit) / scale))
            if v > (1 << 63):
                raise CronError(f"time: invalid duration {orig!r}")
        total += v
        if total > (1 << 63):
            raise CronError(f"time: invalid duration {orig!r}")
    if neg:
        return -total
    if total > (1 << 63) - 1:
        raise CronError(f"time: invalid duration {orig!r}")
    return total


@dataclass
class Cron:
    kind: int = CRON_ERROR
    minute: int = 0
    hour: int = 0
    dom: int = 0
    month: int = 0
    dow: int = 0
    delay_sec: int = 0
    tz_id: int = 0  # SpecSchedule.Location: 0 = UTC, else an id of tz_lookup()

    def masks(self):
        return (self.minute, self.hour, self.dom, self.month, self.dow)


# ---- named time zones (time.LoadLocation): the standard library's zoneinfo, nothing shared
#      with the product's TZif reader or the C oracle's libc calls; ids in order of first appearance
F_TZ_SHIFT = 24
_TZ_NAMES: list[str] = [""]
_TZ_INFO: list = [None]


def tz_lookup(name: str) -> int:
    if name in ("", "UTC", "Local"):
        return 0
    if name in _TZ_NAMES:
        return _TZ_NAMES.index(name)
    import zoneinfo
    if name.startswith(("/", "\\")) or ".." in name or len(name) > 255:
        raise CronError(f"provided bad location {name}")
    try:
        zi = zoneinfo.ZoneInfo(name)
    except Exception:
        raise CronError(f"provided bad location {name}: unknown time zone {name}") from None
    if len(_TZ_NAMES) > 255:
        raise CronUnsupported(name)
    _TZ_NAMES.append(name)
    _TZ_INFO.append(zi)
    return len(_TZ_NAMES) - 1


def _wall(t: int, tz_id: int) -> _dt.datetime:
    """t's wall clock in the zone (naive datetime)"""
    if tz_id == 0:
        return _utc(t)
    return _dt.datetime.fromtimestamp(t, _TZ_INFO[tz_id]).replace(tzinfo=None)


def tz_offset(tz_id: int, t: int) -> int:
    if tz_id == 0:
        return 0
    return int(_dt.datetime.fromtimestamp(t, _TZ_INFO[tz_id]).utcoffset().total_seconds())


def cron_parse(spec: str) -> Cron:
    """cron.ParseStandard (hcc.go:253).  Raises CronError / CronUnsupported."""
    if spec == "":
        raise CronError("empty spec string")
    tz_id = 0
    if spec.startswith("TZ=") or spec.startswith("CRON_TZ="):
        i = spec.find(" ")
        eq = spec.find("=")
        if i < 0:
            raise CronError("robfig v3.0.1 panics: TZ= without a following space")
        loc = spec[eq + 1:i]
        spec = go_trim_space(spec[i:])
        tz_id = tz_lookup(loc)
    if spec.startswith("@"):
        one_min, one_hr = 1 << 0, 1 << 0
        table = {
            "@yearly": (one_min, one_hr, 1 << 1, 1 << 1, _all("dow")),
            "@annually": (one_min, one_hr, 1 << 1, 1 << 1, _all("dow")),
            "@monthly": (one_min, one_hr, 1 << 1, _all("month"), _all("dow")),
            "@weekly": (one_min, one_hr, _all("dom"), _all("month"), 1 << 0),
            "@daily": (one_min, one_hr, _all("dom"), _all("month"), _all("dow")),
            "@midnight": (one_min, one_hr, _all("dom"), _all("month"), _all("dow")),
            "@hourly": (one_min, _all("hour"), _all("dom"), _all("month"), _all("dow")),
        }
        if spec in table:
            mi, hr, dm, mo, dw = table[spec]
            return Cron(CRON_SPEC, mi, hr, dm, mo, dw, tz_id=tz_id)
        if spec.startswith("@every "):
            try:
                ns = parse_duration(spec[len("@every "):])
            except CronError as e:
                raise CronError(f"failed to parse duration {spec}: {e}") from None
            if ns < 10**9:
                ns = 10**9
            return Cron(CRON_EVERY, delay_sec=(ns - ns % 10**9) // 10**9)
        raise CronError(f"unrecognized descriptor: {spec}")
    fields = go_fields(spec)
    if len(fields) != 5:
        raise CronError(f"expected exactly 5 fields, found {len(fields)}: {fields}")
    mi = _field_mask(fields[0], "minute")
    hr = _field_mask(fields[1], "hour")
    dm = _field_mask(fields[2], "dom")
    mo = _field_mask(fields[3], "month")
    dw = _field_mask(fields[4], "dow")
    return Cron(CRON_SPEC, mi, hr, dm, mo, dw, tz_id=tz_id)


_EPOCH = _dt.datetime(1970, 1, 1)


def _utc(t: int) -> _dt.datetime:
    return _EPOCH + _dt.timedelta(seconds=t)


def _go_weekday(d: _dt.date) -> int:
    return (d.weekday() + 1) % 7  # Python Monday=0 -> Go Sunday=0


def _day_ok(c: Cron, d: _dt.date) -> bool:
    dom_ok = bool(c.dom >> d.day & 1)
    dow_ok = bool(c.dow >> _go_weekday(d) & 1)
    if (c.dom | c.dow) & STAR_BIT:
        return dom_ok and dow_ok
    return dom_ok or dow_ok


def civil_from_unix(t: int):
    d = _utc(t)
    return (d.second, d.minute, d.hour, d.day, d.month, _go_weekday(d.date()))


def cron_matches(c: Cron, t: int) -> bool:
    if c.kind != CRON_SPEC:
        return False
    d = _wall(t, c.tz_id)
    return (d.second == 0 and bool(c.minute >> d.minute & 1) and bool(c.hour >> d.hour & 1)
            and bool(c.month >> d.month & 1) and _day_ok(c, d.date()))


def cron_next(c: Cron, t: int):
    """First activation strictly after whole second t, or None (Go zero time)
    when none up to the end of year(t+1s)+5.  Forward day scan."""
    if c.kind == CRON_EVERY:
        return t + c.delay_sec
    if c.kind != CRON_SPEC:
        return None
    if c.tz_id:
        # zone-bound: the first instant after t whose wall clock in the zone matches, by brute force
        # over whole minutes (days whose local date cannot match are skipped a local day at a time)
        y0 = _wall(t + 1, c.tz_id).year
        cand = t + 1
        cand += (-(cand + tz_offset(c.tz_id, cand))) % 60
        while True:
            w = _wall(cand, c.tz_id)
            if w.year > y0 + 5:
                return None
            if not ((c.month >> w.month & 1) and _day_ok(c, w.date())):
                # the first instant whose wall clock shows the next local date (minute steps; a
                # transition may sit anywhere in between, so no arithmetic on offsets here)
                nxt = w.date() + _dt.timedelta(days=1)
                cand += 86400 - (w.hour * 3600 + w.minute * 60 + w.second) - 4 * 3600
                while _wall(cand, c.tz_id).date() < nxt:
                    cand += 60
                continue
            if w.second == 0 and (c.hour >> w.hour & 1) and (c.minute >> w.minute & 1):
                return cand
            cand += 60 - w.second
    start = _utc(t + 1)
    year_limit = start.year + 5
    mins = [m for m in range(60) if c.minute >> m & 1]
    hrs = [h for h in range(24) if c.hour >> h & 1]
    if not mins or not hrs:
        return None
    day = start.date()
    first = True
    while day.year <= year_limit:
        if (c.month >> day.month & 1) and _day_ok(c, day):
            for h in hrs:
                for m in mins:
                    cand = _dt.datetime(day.year, day.month, day.day, h, m, 0)
                    if not first or cand >= start:
                        return int((cand - _EPOCH).total_seconds())
        day += _dt.timedelta(days=1)
        first = False
    return None


def cron_repeat_after_sec(c: Cron, t: int) -> int:
    nx = cron_next(c, t)
    if nx is None:
        return -9223372036 + 1
    return nx - t


# ---------------------------------------------------------------------------
@dataclass
class HealthCheck:
    repeat_after_sec: int = 0
    cron: str = ""
    has_resource: bool = True
    has_remedy: bool = False
    remedy_runs_limit: int = 0
    remedy_reset_interval: int = 0
    finished_at: int | None = None
    remedy_finished_at: int | None = None
    success_count: int = 0
    failed_count: int = 0
    remedy_success_count: int = 0
    remedy_failed_count: int = 0
    remedy_total_runs: int = 0
    fail_p8: int = 0
    timer_armed: bool = True  # RepeatTimersByName has a timer for the check (hcc.go:264)


@dataclass
class Record:
    minute: int = 0
    hour: int = 0
    dom: int = 0
    month: int = 0
    dow: int = 0
    ras: int = 0
    flags: int = 0
    finished_at: int = 0
    runs_limit: int = 0
    reset_interval: int = 0
    success: int = 0
    failed: int = 0
    remedy_success: int = 0
    remedy_failed: int = 0
    remedy_total: int = 0
    remedy_finished_at: int = 0

    @property
    def kind(self):
        return self.flags & 7


def _i32(v):
    return -(1 << 31) <= v <= (1 << 31) - 1


def _wrap32(v):
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v >= 1 << 31 else v


def remedy_is_empty(generate_name: str, resource_is_nil: bool, timeout: int,
                    rbac_rules_is_nil: bool) -> bool:
    return generate_name == "" and resource_is_nil and timeout == 0 and rbac_rules_is_nil


def classify(hc: HealthCheck) -> tuple[int, Record]:
    """Ladder order hcc.go:227 -> :238 -> :251 -> :264."""
    lim = 1 << 55
    for v in (hc.remedy_runs_limit, hc.remedy_reset_interval, hc.success_count, hc.failed_count,
              hc.remedy_success_count, hc.remedy_failed_count, hc.remedy_total_runs):
        if not _i32(v):
            return E_RANGE, Record()
    if hc.finished_at is not None and not (-lim < hc.finished_at < lim):
        return E_RANGE, Record()
    if hc.remedy_finished_at is not None and (
            not (-lim < hc.remedy_finished_at < lim) or hc.remedy_finished_at == 0):
        return E_RANGE, Record()
    if not 0 <= hc.fail_p8 <= 255:
        return E_RANGE, Record()
    r = Record()
    rc = 0
    if not hc.has_resource:
        kind = KIND_NO_RESOURCE
    elif hc.repeat_after_sec <= 0 and hc.cron == "":
        kind = KIND_STOPPED
    elif hc.repeat_after_sec <= 0:
        try:
            c = cron_parse(hc.cron)
        except CronUnsupported:
            kind, rc = KIND_HOST_FALLBACK, E_UNSUPPORTED
        except CronError:
            kind = KIND_PARSE_ERROR
        else:
            if c.kind == CRON_EVERY:
                if not _i32(c.delay_sec):
                    return E_RANGE, Record()
                kind, r.ras = KIND_CRON_EVERY, c.delay_sec
            else:
                kind = KIND_CRON_SPEC | (c.tz_id << F_TZ_SHIFT)
                r.minute, r.hour, r.dom, r.month, r.dow = c.masks()
    else:
        if not _i32(hc.repeat_after_sec):
            return E_RANGE, Record()
        kind, r.ras = KIND_INTERVAL, hc.repeat_after_sec
    r.flags = (kind | (F_HAS_REMEDY if hc.has_remedy else 0) | (hc.fail_p8 << F_FAILP_SHIFT)
               | (F_TIMER_ARMED if hc.timer_armed else 0))
    r.finished_at = hc.finished_at if hc.finished_at is not None else 0
    r.remedy_finished_at = hc.remedy_finished_at if hc.remedy_finished_at is not None else 0
    r.runs_limit, r.reset_interval = hc.remedy_runs_limit, hc.remedy_reset_interval
    r.success, r.failed = hc.success_count, hc.failed_count
    r.remedy_success, r.remedy_failed = hc.remedy_success_count, hc.remedy_failed_count
    r.remedy_total = hc.remedy_total_runs
    return rc, r


def _sm64(z):
    z = (z + 0x9E3779B97F4A7C15) & U64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & U64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & U64
    return z ^ (z >> 31)


def key(seed, i, f):
    return _sm64((_sm64((seed ^ _sm64(i & U64)) & U64) + f) & U64)


@dataclass
class Stats:
    n_records: int = 0
    n_emitted: int = 0
    n_submit_hc: int = 0
    n_run_remedy: int = 0
    n_stopped: int = 0
    n_parse_error: int = 0
    n_remedy_skip: int = 0
    n_reset_on_pass: int = 0
    n_reset_on_interval: int = 0
    n_anomaly: int = 0
    n_result_ok: int = 0
    n_result_fail: int = 0
    n_remedy_ok: int = 0
    n_remedy_fail: int = 0
    idx_xor: int = 0
    idx_sum: int = 0
    extra: dict = _field(default_factory=dict)


def _remedy_result(r: Record, t: int, ok: bool, st: Stats):  # hcc.go:821-851
    if ok:
        r.remedy_success = _wrap32(r.remedy_success + 1)
        st.n_remedy_ok += 1
    else:
        r.remedy_failed = _wrap32(r.remedy_failed + 1)
        st.n_remedy_fail += 1
    r.remedy_total = _wrap32(r.remedy_success + r.remedy_failed)
    r.remedy_finished_at = t


def _reset_remedy(r: Record):
    r.remedy_total = r.remedy_success = r.remedy_failed = 0
    r.remedy_finished_at = 0


def _apply_result(r: Record, t: int, st: Stats) -> int:
    act = 0
    f = r.flags
    if f & F_PENDING_OK:  # hcc.go:635-661
        r.success = _wrap32(r.success + 1)
        r.finished_at = t
        st.n_result_ok += 1
        if f & F_HAS_REMEDY and r.remedy_total >= 1:
            _reset_remedy(r)
            act |= ACT_RESET_ON_PASS
    elif f & F_PENDING_FAIL:  # hcc.go:662-722
        r.failed = _wrap32(r.failed + 1)
        r.finished_at = t
        st.n_result_fail += 1
        if f & F_HAS_REMEDY:
            run = False
            if r.runs_limit != 0 and r.reset_interval != 0:
                if r.runs_limit > r.remedy_total:
                    run = True
                elif r.remedy_finished_at == 0:
                    act |= ACT_ANOMALY
                else:
                    d = max(-9223372036, min(9223372036, t - r.remedy_finished_at))
                    if r.reset_interval >= d:
                        act |= ACT_REMEDY_SKIP
                    else:
                        _reset_remedy(r)
                        act |= ACT_RESET_ON_INTERVAL
                        run = True
            else:
                run = True
            if run:
                act |= ACT_RUN_REMEDY
                if f & F_REMEDY_PENDING:
                    _remedy_result(r, t, bool(f & F_REMEDY_OUTCOME_OK), st)
    elif f & F_REMEDY_PENDING:
        _remedy_result(r, t, bool(f & F_REMEDY_OUTCOME_OK), st)
    r.flags = f & ~(F_PENDING_OK | F_PENDING_FAIL | F_REMEDY_PENDING | F_REMEDY_OUTCOME_OK)
    if f & (F_PENDING_OK | F_PENDING_FAIL):  # hcc.go:745-752: the repeat timer is re-armed after either outcome
        r.flags |= F_TIMER_ARMED
    return act


def tick_record(r: Record, t: int, mode: int = 0, seed: int = 0, gidx: int = 0,
                st: Stats | None = None) -> int:
    """SURVEY Appendix B.3 for one record; returns the action byte."""
    st = st if st is not None else Stats()
    if r.flags & F_TOMBSTONE:
        return 0
    kind = r.kind
    if kind in (KIND_NO_RESOURCE, KIND_HOST_FALLBACK) or kind > KIND_HOST_FALLBACK:
        return 0
    act = _apply_result(r, t, st)
    due = False
    if kind == KIND_STOPPED:
        if not r.flags & F_STOPPED_REPORTED:
            act |= ACT_STOPPED
            r.finished_at = t
            r.flags |= F_STOPPED_REPORTED
    elif kind == KIND_PARSE_ERROR:
        act |= ACT_PARSE_ERROR
    elif kind in (KIND_INTERVAL, KIND_CRON_EVERY):
        # hcc.go:264: skipped iff elapsed < RepeatAfterSec AND a timer exists for the check
        due = not ((t - r.finished_at) < r.ras and bool(r.flags & F_TIMER_ARMED))
    elif kind == KIND_CRON_SPEC:
        due = cron_matches(Cron(CRON_SPEC, r.minute, r.hour, r.dom, r.month, r.dow, tz_id=r.flags >> F_TZ_SHIFT), t)
    if due:
        act |= ACT_SUBMIT_HC
        if mode & MODE_CLOSED_LOOP:
            k = key(seed, gidx, t & U64)
            failp = (r.flags >> F_FAILP_SHIFT) & 0xFF
            fail = (k & 0xFF) < failp
            remedy_ok = ((k >> 8) & 0xFF) < 179
            r.flags |= F_PENDING_FAIL if fail else F_PENDING_OK
            r.flags |= F_REMEDY_PENDING | (F_REMEDY_OUTCOME_OK if remedy_ok else 0)
            act |= _apply_result(r, t, st)
    return act


def sweep(records: list[Record], t: int, mode: int = 0, seed: int = 0, shard_base: int = 0):
    """Whole-array tick.  Returns (due list of (global idx, action), Stats)."""
    st = Stats(n_records=len(records))
    due = []
    for i, r in enumerate(records):
        act = tick_record(r, t, mode, seed, shard_base + i, st)
        if act:
            g = shard_base + i
            due.append((g, act))
            st.n_emitted += 1
            st.n_submit_hc += bool(act & ACT_SUBMIT_HC)
            st.n_run_remedy += bool(act & ACT_RUN_REMEDY)
            st.n_stopped += bool(act & ACT_STOPPED)
            st.n_parse_error += bool(act & ACT_PARSE_ERROR)
            st.n_remedy_skip += bool(act & ACT_REMEDY_SKIP)
            st.n_reset_on_pass += bool(act & ACT_RESET_ON_PASS)
            st.n_reset_on_interval += bool(act & ACT_RESET_ON_INTERVAL)
            st.n_anomaly += bool(act & ACT_ANOMALY)
            st.idx_xor ^= g
            st.idx_sum = (st.idx_sum + g) & U64
    return due, st
