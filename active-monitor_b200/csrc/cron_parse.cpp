// cron_parse.cpp — host side of libamsweep that runs at UPSERT time, never in
// the tick: cron string -> five 64-bit field masks, HealthCheck -> packed
// record (ladder classification), plus small host helpers.
//
// Replaces, for the shim, the per-reconcile calls
//   cron.ParseStandard(hcSpec.Schedule.Cron)            hcc.go:253
//   schedule.Next(time.Now())                            hcc.go:262
// of github.com/robfig/cron/v3 v3.0.1 (go.mod:14) and the branch order of
// processHealthCheck, hcc.go:227 -> :238 -> :251 -> :264.
//
// Written as a byte-cursor scanner over the spec (no intermediate string
// splitting); it shares no code with oracle/, which is the checker.
#include <climits>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <string_view>
#include <atomic>
#include <thread>
#include <vector>

#include "../../include/amsweep.h"
#include "civil.h"
#include "tz.h"

// the ABI's struct layouts are part of the contract (tests/test_abi.py)
static_assert(sizeof(am_cron_t) == 56, "am_cron_t layout");
static_assert(sizeof(am_healthcheck_t) == 120, "am_healthcheck_t layout");
static_assert(sizeof(am_record_t) == 96, "am_record_t layout");
static_assert(sizeof(am_tick_stats_t) == 128, "am_tick_stats_t layout");
static_assert(sizeof(am_record_cols_t) == 128, "am_record_cols_t layout");

namespace {

using std::string_view;

// ---- error text ---------------------------------------------------------
struct ErrSink {
  char* buf;
  size_t cap;
  void put(const char* fmt, ...) const __attribute__((format(printf, 2, 3))) {
    if (!buf || cap == 0) return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, cap, fmt, ap);
    va_end(ap);
  }
};
#define SV(x) (int)(x).size(), (x).data()

// ---- Go text semantics --------------------------------------------------
// Decode one UTF-8 rune the way Go's utf8.DecodeRune does: any malformed
// sequence is U+FFFD consuming ONE byte.
inline uint32_t rune_at(string_view s, size_t i, size_t& width) {
  auto b = [&](size_t k) { return (unsigned char)s[i + k]; };
  size_t left = s.size() - i;
  unsigned c = b(0);
  width = 1;
  if (c < 0x80) return c;
  if (c < 0xC2 || c > 0xF4) return 0xFFFD;
  auto cont = [&](size_t k) { return k < left && (b(k) & 0xC0) == 0x80; };
  if (c < 0xE0) {
    if (!cont(1)) return 0xFFFD;
    width = 2;
    return ((c & 0x1Fu) << 6) | (b(1) & 0x3Fu);
  }
  if (c < 0xF0) {
    if (!cont(1) || !cont(2)) return 0xFFFD;
    if ((c == 0xE0 && b(1) < 0xA0) || (c == 0xED && b(1) > 0x9F)) return 0xFFFD;
    width = 3;
    return ((c & 0x0Fu) << 12) | ((b(1) & 0x3Fu) << 6) | (b(2) & 0x3Fu);
  }
  if (!cont(1) || !cont(2) || !cont(3)) return 0xFFFD;
  if ((c == 0xF0 && b(1) < 0x90) || (c == 0xF4 && b(1) > 0x8F)) return 0xFFFD;
  width = 4;
  return ((c & 0x07u) << 18) | ((b(1) & 0x3Fu) << 12) | ((b(2) & 0x3Fu) << 6) | (b(3) & 0x3Fu);
}

// unicode.IsSpace (White_Space property)
inline bool white(uint32_t r) {
  if (r <= 0x20) return r == ' ' || (r >= 0x09 && r <= 0x0D);
  if (r < 0x1680) return r == 0x85 || r == 0xA0;
  return r == 0x1680 || (r >= 0x2000 && r <= 0x200A) || r == 0x2028 || r == 0x2029 ||
         r == 0x202F || r == 0x205F || r == 0x3000;
}

// Walk whitespace-separated words (strings.Fields).  Stores up to `keep`
// words; returns how many there are in total.
size_t words(string_view s, string_view* out, size_t keep) {
  size_t n = 0, i = 0;
  while (i < s.size()) {
    size_t w;
    // skip blanks
    while (i < s.size() && white(rune_at(s, i, w))) i += w;
    if (i >= s.size()) break;
    size_t b = i;
    while (i < s.size() && !white(rune_at(s, i, w))) i += w;
    if (n < keep) out[n] = s.substr(b, i - b);
    ++n;
  }
  return n;
}

// strings.TrimSpace
string_view trim(string_view s) {
  size_t b = 0, w;
  while (b < s.size() && white(rune_at(s, b, w))) b += w;
  // find the end of the last non-blank rune by a forward walk from b
  size_t e = b, i = b;
  while (i < s.size()) {
    uint32_t r = rune_at(s, i, w);
    i += w;
    if (!white(r)) e = i;
  }
  return s.substr(b, e - b);
}

// strconv.Atoi acceptance, then robfig's mustParseInt (no negatives).
enum class Num { Ok, Syntax, Range, Negative };
Num go_uint(string_view t, uint64_t& v, int64_t* signed_out = nullptr) {
  size_t i = 0;
  bool neg = false;
  if (!t.empty() && (t[0] == '+' || t[0] == '-')) { neg = t[0] == '-'; i = 1; }
  if (i >= t.size()) return Num::Syntax;
  unsigned __int128 acc = 0;
  bool big = false;
  for (; i < t.size(); ++i) {
    unsigned d = (unsigned char)t[i] - '0';
    if (d > 9) return Num::Syntax;
    if (!big) {
      acc = acc * 10 + d;
      if (acc > ((unsigned __int128)1 << 63)) big = true;
    }
  }
  const unsigned __int128 lim = neg ? ((unsigned __int128)1 << 63) : (((unsigned __int128)1 << 63) - 1);
  if (big || acc > lim) return Num::Range;
  if (neg && acc != 0) {
    if (signed_out) *signed_out = -(int64_t)(uint64_t)(acc - 1) - 1;
    return Num::Negative;
  }
  v = (uint64_t)acc;
  return Num::Ok;
}

bool number(string_view t, uint64_t& v, const ErrSink& es) {
  int64_t sv = 0;
  switch (go_uint(t, v, &sv)) {
    case Num::Ok: return true;
    case Num::Negative:
      es.put("negative number (%lld) not allowed: %.*s", (long long)sv, SV(t));
      return false;
    case Num::Range:
      es.put("failed to parse int from %.*s: strconv.Atoi: parsing \"%.*s\": value out of range",
             SV(t), SV(t));
      return false;
    default:
      es.put("failed to parse int from %.*s: strconv.Atoi: parsing \"%.*s\": invalid syntax", SV(t),
             SV(t));
      return false;
  }
}

// ---- field domains ------------------------------------------------------
struct Domain {
  uint32_t lo, hi;
  const char* names;  // 3-letter names back to back, or nullptr
  uint32_t first;     // value of names[0..3)
};
constexpr Domain kMinute{0, 59, nullptr, 0};
constexpr Domain kHour{0, 23, nullptr, 0};
constexpr Domain kDom{1, 31, nullptr, 0};
constexpr Domain kMonth{1, 12, "janfebmaraprmayjunjulaugsepoctnovdec", 1};
constexpr Domain kDow{0, 6, "sunmontuewedthufrisat", 0};

constexpr uint64_t span_mask(uint32_t lo, uint32_t hi) {  // bits lo..hi inclusive, hi <= 62
  return ((~0ull) >> (63 - hi)) & ((~0ull) << lo);
}
constexpr uint64_t every_value(const Domain& d) { return span_mask(d.lo, d.hi) | AM_STAR_BIT; }

// strings.ToLower(name) == one of the 3-letter names?  Only ASCII letters (and
// the two non-ASCII runes whose simple lower-case is ASCII: U+0130 -> i,
// U+212A -> k) can produce an ASCII name.
bool named_value(string_view t, const Domain& d, uint64_t& v) {
  if (!d.names) return false;
  char low[3];
  size_t n = 0, i = 0;
  while (i < t.size()) {
    size_t w;
    uint32_t r = rune_at(t, i, w);
    i += w;
    char c;
    if (r < 0x80) c = (r >= 'A' && r <= 'Z') ? (char)(r | 0x20) : (char)r;
    else if (r == 0x130) c = 'i';
    else if (r == 0x212A) c = 'k';
    else return false;
    if (n == 3) return false;
    low[n++] = c;
  }
  if (n != 3) return false;
  for (uint32_t k = 0; d.names[3 * k]; ++k) {
    if (memcmp(low, d.names + 3 * k, 3) == 0) { v = d.first + k; return true; }
  }
  return false;
}

bool value_or_name(string_view t, const Domain& d, uint64_t& v, const ErrSink& es) {
  return named_value(t, d, v) || number(t, v, es);
}

// One comma piece:  lo[-hi][/step]   ('*' and '?' stand for the whole domain)
bool piece(string_view expr, const Domain& d, uint64_t& bits, const ErrSink& es) {
  // left of the first '/', and how many '/' there are
  size_t slash = expr.find('/');
  string_view left = expr.substr(0, slash);
  size_t n_slash = 0;
  for (char c : expr) n_slash += (c == '/');
  // left part: up to the first '-', the part after it, and the '-' count
  size_t dash = left.find('-');
  string_view a = left.substr(0, dash);
  size_t n_dash = 0;
  for (char c : left) n_dash += (c == '-');

  uint64_t from = 0, to = 0, step = 1;
  bool star = false;
  if (a == "*" || a == "?") {  // anything after '*-' is ignored by robfig
    from = d.lo;
    to = d.hi;
    star = true;
  } else {
    if (!value_or_name(a, d, from, es)) return false;
    if (n_dash == 0) {
      to = from;
    } else if (n_dash == 1) {
      if (!value_or_name(left.substr(dash + 1), d, to, es)) return false;
    } else {
      es.put("too many hyphens: %.*s", SV(expr));
      return false;
    }
  }
  if (n_slash == 1) {
    if (!number(expr.substr(slash + 1), step, es)) return false;
    if (n_dash == 0) to = d.hi;  // "N/step" means "N-max/step"
    if (step > 1) star = false;
  } else if (n_slash > 1) {
    es.put("too many slashes: %.*s", SV(expr));
    return false;
  }
  if (from < d.lo) {
    es.put("beginning of range (%llu) below minimum (%u): %.*s", (unsigned long long)from, d.lo,
           SV(expr));
    return false;
  }
  if (to > d.hi) {
    es.put("end of range (%llu) above maximum (%u): %.*s", (unsigned long long)to, d.hi, SV(expr));
    return false;
  }
  if (from > to) {
    es.put("beginning of range (%llu) beyond end of range (%llu): %.*s", (unsigned long long)from,
           (unsigned long long)to, SV(expr));
    return false;
  }
  if (step == 0) {
    es.put("step of range should be a positive number: %.*s", SV(expr));
    return false;
  }
  uint64_t m = 0;
  if (step == 1) {
    m = span_mask((uint32_t)from, (uint32_t)to);
  } else {
    for (uint64_t v = from; v <= to; v += step) m |= 1ull << v;
  }
  bits = m | (star ? AM_STAR_BIT : 0);
  return true;
}

bool field(string_view text, const Domain& d, uint64_t& mask, const ErrSink& es) {
  uint64_t acc = 0;
  size_t i = 0;
  while (i <= text.size()) {
    size_t j = text.find(',', i);
    if (j == string_view::npos) j = text.size();
    if (j > i) {  // empty pieces are dropped (strings.FieldsFunc)
      uint64_t b;
      if (!piece(text.substr(i, j - i), d, b, es)) return false;
      acc |= b;
    }
    i = j + 1;
  }
  mask = acc;
  return true;
}

// ---- time.ParseDuration -------------------------------------------------
bool duration_ns(string_view s, int64_t& out) {
  constexpr uint64_t kTop = 1ull << 63;
  bool neg = false;
  if (!s.empty() && (s[0] == '+' || s[0] == '-')) { neg = s[0] == '-'; s.remove_prefix(1); }
  if (s == "0") { out = 0; return true; }
  if (s.empty()) return false;
  uint64_t total = 0;
  auto digit = [](char c) { return c >= '0' && c <= '9'; };
  while (!s.empty()) {
    if (!(s[0] == '.' || digit(s[0]))) return false;
    // integer part
    uint64_t whole = 0;
    size_t i = 0;
    for (; i < s.size() && digit(s[i]); ++i) {
      if (whole > kTop / 10) return false;
      whole = whole * 10 + (uint64_t)(s[i] - '0');
      if (whole > kTop) return false;
    }
    bool had_int = i > 0;
    s.remove_prefix(i);
    // fraction
    uint64_t frac = 0;
    double scale = 1.0;
    bool had_frac = false;
    if (!s.empty() && s[0] == '.') {
      s.remove_prefix(1);
      bool sat = false;
      for (i = 0; i < s.size() && digit(s[i]); ++i) {
        if (sat) continue;
        if (frac > (kTop - 1) / 10) { sat = true; continue; }
        uint64_t y = frac * 10 + (uint64_t)(s[i] - '0');
        if (y > kTop) { sat = true; continue; }
        frac = y;
        scale *= 10.0;
      }
      had_frac = i > 0;
      s.remove_prefix(i);
    }
    if (!had_int && !had_frac) return false;
    // unit: everything up to the next digit or '.'
    for (i = 0; i < s.size() && !(s[i] == '.' || digit(s[i])); ++i) {}
    if (i == 0) return false;
    string_view u = s.substr(0, i);
    s.remove_prefix(i);
    uint64_t unit;
    if (u == "ns") unit = 1ull;
    else if (u == "us" || u == "\xC2\xB5s" || u == "\xCE\xBCs") unit = 1000ull;
    else if (u == "ms") unit = 1000000ull;
    else if (u == "s") unit = 1000000000ull;
    else if (u == "m") unit = 60000000000ull;
    else if (u == "h") unit = 3600000000000ull;
    else return false;
    if (whole > kTop / unit) return false;
    whole *= unit;
    if (frac > 0) {
      whole += (uint64_t)((double)frac * ((double)unit / scale));
      if (whole > kTop) return false;
    }
    total += whole;
    if (total > kTop) return false;
  }
  if (neg) { out = (int64_t)(0 - total); return true; }
  if (total > kTop - 1) return false;
  out = (int64_t)total;
  return true;
}

// ---- descriptors --------------------------------------------------------
struct Named { const char* text; uint64_t mi, hr, dm, mo, dw; };
const Named kNamed[] = {
    {"@yearly", 1, 1, 1ull << 1, 1ull << 1, every_value(kDow)},
    {"@annually", 1, 1, 1ull << 1, 1ull << 1, every_value(kDow)},
    {"@monthly", 1, 1, 1ull << 1, every_value(kMonth), every_value(kDow)},
    {"@weekly", 1, 1, every_value(kDom), every_value(kMonth), 1ull << 0},
    {"@daily", 1, 1, every_value(kDom), every_value(kMonth), every_value(kDow)},
    {"@midnight", 1, 1, every_value(kDom), every_value(kMonth), every_value(kDow)},
    {"@hourly", 1, every_value(kHour), every_value(kDom), every_value(kMonth), every_value(kDow)},
};

bool fits32(int64_t v) { return v >= INT32_MIN && v <= INT32_MAX; }
constexpr int64_t kTimeLimit = 1ll << 55;

}  // namespace

// =========================================================================
extern "C" {

int am_abi_version(void) { return AMSWEEP_ABI_VERSION; }

const char* am_strerror(int code) {
  switch (code) {
    case AM_OK: return "ok";
    case AM_E_INVAL: return "invalid argument";
    case AM_E_RANGE: return "value outside the record column's domain";
    case AM_E_NOSPACE: return "output buffer too small";
    case AM_E_DEVICE: return "CUDA device error";
    case AM_E_NOMEM: return "out of memory";
    case AM_E_PARSE: return "cron spec rejected";
    case AM_E_UNSUPPORTED: return "spec valid but not evaluated on the device path";
    case AM_E_BUSY: return "a tick is already running on this handle";
    default: return "unknown amsweep error";
  }
}

int am_cron_parse(const char* spec_p, size_t len, am_cron_t* out, char* err, size_t errcap) {
  if (!out || (!spec_p && len)) return AM_E_INVAL;
  ErrSink es{err, errcap};
  if (err && errcap) err[0] = 0;
  memset(out, 0, sizeof *out);
  string_view spec(spec_p ? spec_p : "", len);
  if (spec.empty()) { es.put("empty spec string"); return AM_E_PARSE; }
  int32_t tz_id = 0;

  if (spec.rfind("TZ=", 0) == 0 || spec.rfind("CRON_TZ=", 0) == 0) {
    size_t sp = spec.find(' ');
    size_t eq = spec.find('=');
    if (sp == string_view::npos) {
      // robfig v3.0.1 slices spec[eq+1:-1] and panics; the reconcile recovers
      // it (hcc.go:191-195) and the check never runs.
      es.put("provided bad location: no space after the time zone");
      return AM_E_PARSE;
    }
    string_view zone = spec.substr(eq + 1, sp - eq - 1);
    spec = trim(spec.substr(sp));
    // time.LoadLocation(zone): registered once, carried by id (tz.h); "", "UTC", "Local" -> 0
    int32_t id = 0;
    const int trc = amsweep_tz::lookup(zone.data(), zone.size(), &id);
    if (trc == -1) {
      es.put("provided bad location %.*s: unknown time zone %.*s", SV(zone), SV(zone));
      return AM_E_PARSE;
    }
    if (trc != 0) {
      es.put("time zone %.*s: more than %d distinct zones are not evaluated on the device path", SV(zone),
             amsweep_tz::kMaxZones);
      return AM_E_UNSUPPORTED;
    }
    tz_id = id;
  }

  if (!spec.empty() && spec[0] == '@') {
    for (const Named& n : kNamed) {
      if (spec == n.text) {
        out->kind = AM_CRON_SPEC;
        out->minute = n.mi; out->hour = n.hr; out->dom = n.dm; out->month = n.mo; out->dow = n.dw;
        out->tz_id = tz_id;  // SpecSchedule.Location
        return AM_OK;
      }
    }
    constexpr string_view kEvery = "@every ";
    if (spec.substr(0, kEvery.size()) == kEvery) {
      int64_t ns;
      if (!duration_ns(spec.substr(kEvery.size()), ns)) {
        es.put("failed to parse duration %.*s", SV(spec));
        return AM_E_PARSE;
      }
      // cron.Every: below one second rounds up to it; sub-second part dropped
      if (ns < 1000000000ll) ns = 1000000000ll;
      out->kind = AM_CRON_EVERY;
      out->delay_sec = ns / 1000000000ll;
      return AM_OK;
    }
    es.put("unrecognized descriptor: %.*s", SV(spec));
    return AM_E_PARSE;
  }

  string_view f[5];
  size_t nf = words(spec, f, 5);
  if (nf != 5) {
    es.put("expected exactly 5 fields, found %zu: [%.*s]", nf, SV(spec));
    return AM_E_PARSE;
  }
  am_cron_t c{};
  if (!field(f[0], kMinute, c.minute, es) || !field(f[1], kHour, c.hour, es) ||
      !field(f[2], kDom, c.dom, es) || !field(f[3], kMonth, c.month, es) ||
      !field(f[4], kDow, c.dow, es))
    return AM_E_PARSE;
  c.kind = AM_CRON_SPEC;
  c.tz_id = tz_id;  // SpecSchedule.Location
  *out = c;
  return AM_OK;
}

int am_tz_lookup(const char* name, size_t len, int32_t* tz_id) {
  if (!tz_id || (!name && len)) return AM_E_INVAL;
  const int rc = amsweep_tz::lookup(name, len, tz_id);
  return rc == 0 ? AM_OK : (rc == -1 ? AM_E_PARSE : AM_E_UNSUPPORTED);
}

int am_tz_offset(int32_t tz_id, int64_t unix_sec, int32_t* utoff_out) {
  if (!utoff_out) return AM_E_INVAL;
  return amsweep_tz::offset_at(tz_id, unix_sec, utoff_out) ? AM_OK : AM_E_INVAL;
}

void am_civil_from_unix(int64_t unix_sec, int32_t out[6]) {
  amsweep::CivilTime c = amsweep::civil_from_unix(unix_sec);
  out[0] = c.sec; out[1] = c.min; out[2] = c.hour; out[3] = c.dom; out[4] = c.month; out[5] = c.dow;
}

int am_cron_matches(const am_cron_t* c, int64_t unix_sec) {
  if (!c || c->kind != AM_CRON_SPEC) return 0;
  int32_t off = 0;
  if (c->tz_id && !amsweep_tz::offset_at(c->tz_id, unix_sec, &off)) return 0;
  amsweep::TickWords w = amsweep::tick_words_from_unix(unix_sec + off);  // the zone's wall clock
  if (!w.sec0) return 0;
  if (!(c->minute & w.minute) || !(c->hour & w.hour) || !(c->month & w.month)) return 0;
  bool a = (c->dom & w.dom) != 0, b = (c->dow & w.dow) != 0;
  return ((c->dom | c->dow) & AM_STAR_BIT) ? (a && b) : (a || b);
}

int64_t am_cron_next(const am_cron_t* c, int64_t t) {
  if (!c) return INT64_MIN;
  if (c->kind == AM_CRON_EVERY) return t + c->delay_sec;
  if (c->kind != AM_CRON_SPEC) return INT64_MIN;
  if (c->tz_id == 0) return amsweep::cron_next_utc(c->minute, c->hour, c->dom, c->month, c->dow, t);
  // Zone-bound schedule: the first instant after t whose LOCAL wall clock matches.  Days whose local
  // date cannot match are skipped whole (to the next local midnight, re-evaluated with the offset
  // in force there); inside a candidate day, minute by minute — daylight-saving gaps (local times
  // that never occur) and overlaps (that occur twice) fall out of evaluating real instants, as in
  // robfig's Next, which also works on absolute time.  Five-year horizon as upstream.
  const bool star = ((c->dom | c->dow) >> 63) != 0;
  int32_t off = 0;
  if (!amsweep_tz::offset_at(c->tz_id, t + 1, &off)) return INT64_MIN;
  int64_t d0;
  int32_t sod;
  amsweep::split_days(t + 1 + off, d0, sod);
  int64_t y0;
  int32_t m0, dd0;
  amsweep::civil_from_days(d0, y0, m0, dd0);
  int64_t cand = t + 1 + ((60 - (t + 1 + off) % 60) % 60 + 60) % 60;  // next whole local minute (offsets may carry seconds)
  for (;;) {
    if (!amsweep_tz::offset_at(c->tz_id, cand, &off)) return INT64_MIN;
    int64_t days, y;
    int32_t m, d;
    amsweep::split_days(cand + off, days, sod);
    amsweep::civil_from_days(days, y, m, d);
    if (y > y0 + 5) return INT64_MIN;
    const bool a = (c->dom >> d) & 1, b = (c->dow >> amsweep::weekday_from_days(days)) & 1;
    const bool day_ok = ((c->month >> m) & 1) && (star ? (a && b) : (a || b));
    if (!day_ok) {
      // To the first instant of the next local day.  Under the offset in force now that is
      // cand + 86400 - sod; if the offset changes in between (a transition day) the jump may land
      // past local midnight — walk back to the first minute that still carries the new date — or
      // short of it (the loop then simply continues from there).
      int64_t nx = cand + 86400 - sod;
      int32_t o2 = 0;
      if (!amsweep_tz::offset_at(c->tz_id, nx, &o2)) return INT64_MIN;
      int64_t d2;
      int32_t s2;
      amsweep::split_days(nx + o2, d2, s2);
      if (d2 > days && s2 != 0) {
        for (int k = 0; k < 300; ++k) {
          int32_t o3 = 0;
          if (!amsweep_tz::offset_at(c->tz_id, nx - 60, &o3)) return INT64_MIN;
          int64_t d3;
          int32_t s3;
          amsweep::split_days(nx - 60 + o3, d3, s3);
          if (d3 != d2) break;
          nx -= 60;
        }
      }
      cand = nx;
      continue;
    }
    if (sod % 60 == 0 && ((c->hour >> (sod / 3600)) & 1) && ((c->minute >> ((sod % 3600) / 60)) & 1)) return cand;
    cand += 60 - sod % 60;
  }
}

int64_t am_cron_repeat_after_sec(const am_cron_t* c, int64_t unix_sec) {
  return amsweep::repeat_after_from_next(am_cron_next(c, unix_sec), unix_sec);
}

int am_remedy_is_empty(size_t generate_name_len, int resource_is_nil, int64_t timeout,
                       int rbac_rules_is_nil) {
  return generate_name_len == 0 && resource_is_nil != 0 && timeout == 0 && rbac_rules_is_nil != 0;
}

int am_healthcheck_classify(const am_healthcheck_t* hc, am_record_t* out) {
  if (!hc || !out) return AM_E_INVAL;
  memset(out, 0, sizeof *out);
  const int64_t small[] = {hc->remedy_runs_limit,    hc->remedy_reset_interval, hc->success_count,
                           hc->failed_count,         hc->remedy_success_count,  hc->remedy_failed_count,
                           hc->remedy_total_runs};
  for (int64_t v : small)
    if (!fits32(v)) return AM_E_RANGE;
  if (hc->finished_at_set && (hc->finished_at >= kTimeLimit || hc->finished_at <= -kTimeLimit))
    return AM_E_RANGE;
  if (hc->remedy_finished_at_set &&
      (hc->remedy_finished_at >= kTimeLimit || hc->remedy_finished_at <= -kTimeLimit ||
       hc->remedy_finished_at == 0))
    return AM_E_RANGE;
  if (hc->fail_p8 > 255) return AM_E_RANGE;

  int rc = AM_OK;
  uint32_t kind;
  int32_t ras = 0;
  am_cron_t c{};
  const bool has_cron = hc->cron_len != 0;
  if (!hc->has_resource) {
    kind = AM_KIND_NO_RESOURCE;
  } else if (hc->repeat_after_sec > 0) {  // the last two arms of the ladder: cron ignored
    if (!fits32(hc->repeat_after_sec)) return AM_E_RANGE;
    kind = AM_KIND_INTERVAL;
    ras = (int32_t)hc->repeat_after_sec;
  } else if (!has_cron) {
    kind = AM_KIND_STOPPED;
  } else {
    int prc = am_cron_parse(hc->cron, hc->cron_len, &c, nullptr, 0);
    if (prc == AM_E_UNSUPPORTED) {
      kind = AM_KIND_HOST_FALLBACK;
      rc = AM_E_UNSUPPORTED;
    } else if (prc != AM_OK) {
      kind = AM_KIND_PARSE_ERROR;
    } else if (c.kind == AM_CRON_EVERY) {
      if (!fits32(c.delay_sec)) return AM_E_RANGE;
      kind = AM_KIND_CRON_EVERY;
      ras = (int32_t)c.delay_sec;
    } else {
      kind = AM_KIND_CRON_SPEC;
      out->minute = c.minute; out->hour = c.hour; out->dom = c.dom; out->month = c.month;
      out->dow = c.dow;
    }
  }
  if (kind == AM_KIND_CRON_SPEC) kind |= (uint32_t)c.tz_id << AM_F_TZ_SHIFT;  // SpecSchedule.Location -> table index
  out->flags = kind | (hc->has_remedy ? AM_F_HAS_REMEDY : 0u) | (hc->fail_p8 << AM_F_FAILP_SHIFT) |
               (hc->timer_armed ? AM_F_TIMER_ARMED : 0u);  // r.GetTimerByName(name) != nil, hcc.go:264
  out->ras = ras;
  out->finished_at = hc->finished_at_set ? hc->finished_at : 0;
  out->remedy_finished_at = hc->remedy_finished_at_set ? hc->remedy_finished_at : 0;
  out->runs_limit = (int32_t)hc->remedy_runs_limit;
  out->reset_interval = (int32_t)hc->remedy_reset_interval;
  out->success = (int32_t)hc->success_count;
  out->failed = (int32_t)hc->failed_count;
  out->remedy_success = (int32_t)hc->remedy_success_count;
  out->remedy_failed = (int32_t)hc->remedy_failed_count;
  out->remedy_total = (int32_t)hc->remedy_total_runs;
  return rc;
}

int am_healthcheck_classify_batch(const am_healthcheck_t* hcs, uint64_t n, am_record_t* out,
                                  int32_t* rc_out, int n_threads, uint64_t* n_not_ok) {
  if (n && (!hcs || !out)) return AM_E_INVAL;
  if (n_not_ok) *n_not_ok = 0;
  if (n == 0) return AM_OK;
  unsigned nt = n_threads > 0 ? (unsigned)n_threads : std::thread::hardware_concurrency();
  if (nt == 0) nt = 1;
  const uint64_t kMinPerThread = 4096;  // below this a thread costs more than it saves
  if ((uint64_t)nt > (n + kMinPerThread - 1) / kMinPerThread) nt = (unsigned)((n + kMinPerThread - 1) / kMinPerThread);
  std::atomic<uint64_t> bad{0};
  auto work = [&](uint64_t lo, uint64_t hi) {
    uint64_t b = 0;
    for (uint64_t i = lo; i < hi; ++i) {
      const int rc = am_healthcheck_classify(&hcs[i], &out[i]);
      if (rc_out) rc_out[i] = rc;
      b += rc != AM_OK;
    }
    bad.fetch_add(b, std::memory_order_relaxed);
  };
  const uint64_t per = (n + nt - 1) / nt;
  std::vector<std::thread> th;
  uint64_t next_lo = per < n ? per : n;  // [0, per) is the calling thread's share
  try {
    th.reserve(nt - 1);
    while (next_lo < n) {
      const uint64_t hi = next_lo + per < n ? next_lo + per : n;
      th.emplace_back(work, next_lo, hi);
      next_lo = hi;
    }
  } catch (...) {
    // thread or memory exhaustion: the calling thread runs whatever was not started
  }
  work(0, per < n ? per : n);
  if (next_lo < n) work(next_lo, n);
  for (auto& t : th) t.join();
  if (n_not_ok) *n_not_ok = bad.load();
  return AM_OK;
}

}  // extern "C"
