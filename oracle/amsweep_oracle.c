/*
 * amsweep_oracle.c — CPU ORACLE (test infrastructure, see amsweep_oracle.h).
 *
 * Restates, one Go function per C function and in the Go code's own control
 * flow, the reference decisions listed in the header.  "parity unpinned"
 * caveats are in the header; read them before trusting a green test.
 *
 * Clarity first: no SIMD, no tables, libc calendar.  The product
 * (active-monitor_b200/csrc) shares no code with this file.
 */
#define _GNU_SOURCE
#include "amsweep_oracle.h"

#include <limits.h>
#include <pthread.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

_Static_assert(sizeof(orc_cron_t) == 56, "layout");
_Static_assert(sizeof(orc_healthcheck_t) == 120, "layout");
_Static_assert(sizeof(orc_record_t) == 96, "layout");
_Static_assert(sizeof(orc_tick_stats_t) == 128, "layout");

/* ------------------------------------------------------------------------ */
/* constants mirrored from include/amsweep.h (values are the contract)       */
#define KIND_MASK 0x7u
#define KIND_NO_RESOURCE 0u
#define KIND_STOPPED 1u
#define KIND_INTERVAL 2u
#define KIND_CRON_SPEC 3u
#define KIND_CRON_EVERY 4u
#define KIND_PARSE_ERROR 5u
#define KIND_HOST_FALLBACK 6u
#define F_HAS_REMEDY (1u << 3)
#define F_PENDING_OK (1u << 4)
#define F_PENDING_FAIL (1u << 5)
#define F_REMEDY_PENDING (1u << 6)
#define F_REMEDY_OUTCOME_OK (1u << 7)
#define F_TOMBSTONE (1u << 8)
#define F_STOPPED_REPORTED (1u << 9)
#define F_TIMER_ARMED (1u << 10) /* r.GetTimerByName(name) != nil, hcc.go:264 */
#define F_FAILP_SHIFT 16
#define F_TZ_SHIFT 24 /* bits 24..31: time zone id of a 5-field schedule, 0 = UTC */
#define ACT_SUBMIT_HC 0x01u
#define ACT_RUN_REMEDY 0x02u
#define ACT_STOPPED 0x04u
#define ACT_PARSE_ERROR 0x08u
#define ACT_REMEDY_SKIP 0x10u
#define ACT_RESET_ON_PASS 0x20u
#define ACT_RESET_ON_INTERVAL 0x40u
#define ACT_ANOMALY 0x80u
#define MODE_CLOSED_LOOP 0x1u
#define CRON_ERROR 0
#define CRON_SPEC 1
#define CRON_EVERY 2
#define STAR_BIT (1ull << 63)
#define E_INVAL (-1)
#define E_RANGE (-2)
#define E_NOSPACE (-3)
#define E_PARSE (-6)
#define E_UNSUPPORTED (-7)

/* ======================================================================== */
/* Go runtime helpers: utf8, unicode.IsSpace, strings.Fields, strconv.Atoi   */
/* ======================================================================== */

/* utf8.DecodeRune: invalid encodings yield (U+FFFD, width 1). */
static int go_decode_rune(const unsigned char* s, size_t n, uint32_t* r) {
  if (n == 0) { *r = 0xFFFD; return 0; }
  unsigned char c0 = s[0];
  if (c0 < 0x80) { *r = c0; return 1; }
  if (c0 < 0xC2 || c0 > 0xF4) { *r = 0xFFFD; return 1; }
  if (c0 < 0xE0) {
    if (n < 2 || (s[1] & 0xC0) != 0x80) { *r = 0xFFFD; return 1; }
    *r = ((uint32_t)(c0 & 0x1F) << 6) | (s[1] & 0x3F);
    return 2;
  }
  if (c0 < 0xF0) {
    if (n < 3 || (s[1] & 0xC0) != 0x80 || (s[2] & 0xC0) != 0x80) { *r = 0xFFFD; return 1; }
    if (c0 == 0xE0 && s[1] < 0xA0) { *r = 0xFFFD; return 1; } /* overlong  */
    if (c0 == 0xED && s[1] > 0x9F) { *r = 0xFFFD; return 1; } /* surrogate */
    *r = ((uint32_t)(c0 & 0x0F) << 12) | ((uint32_t)(s[1] & 0x3F) << 6) | (s[2] & 0x3F);
    return 3;
  }
  if (n < 4 || (s[1] & 0xC0) != 0x80 || (s[2] & 0xC0) != 0x80 || (s[3] & 0xC0) != 0x80) {
    *r = 0xFFFD; return 1;
  }
  if (c0 == 0xF0 && s[1] < 0x90) { *r = 0xFFFD; return 1; }
  if (c0 == 0xF4 && s[1] > 0x8F) { *r = 0xFFFD; return 1; }
  *r = ((uint32_t)(c0 & 0x07) << 18) | ((uint32_t)(s[1] & 0x3F) << 12) |
       ((uint32_t)(s[2] & 0x3F) << 6) | (s[3] & 0x3F);
  return 4;
}

/* unicode.IsSpace */
static int go_is_space(uint32_t r) {
  switch (r) {
    case '\t': case '\n': case '\v': case '\f': case '\r': case ' ':
    case 0x85: case 0xA0: case 0x1680: case 0x2028: case 0x2029:
    case 0x202F: case 0x205F: case 0x3000:
      return 1;
  }
  return r >= 0x2000 && r <= 0x200A;
}

typedef struct { const char* p; size_t n; } str_t;

/* strings.Fields: split around runs of unicode.IsSpace; returns count, and the
 * first `cap` fields in out[]. */
static size_t go_fields(str_t s, str_t* out, size_t cap) {
  size_t count = 0, i = 0;
  long start = -1;
  while (i < s.n) {
    uint32_t r;
    int w = go_decode_rune((const unsigned char*)s.p + i, s.n - i, &r);
    if (go_is_space(r)) {
      if (start >= 0) {
        if (count < cap) { out[count].p = s.p + start; out[count].n = i - (size_t)start; }
        count++;
        start = -1;
      }
    } else if (start < 0) {
      start = (long)i;
    }
    i += (size_t)w;
  }
  if (start >= 0) {
    if (count < cap) { out[count].p = s.p + start; out[count].n = s.n - (size_t)start; }
    count++;
  }
  return count;
}

/* strings.TrimSpace */
static str_t go_trim_space(str_t s) {
  size_t b = 0;
  while (b < s.n) {
    uint32_t r;
    int w = go_decode_rune((const unsigned char*)s.p + b, s.n - b, &r);
    if (!go_is_space(r)) break;
    b += (size_t)w;
  }
  size_t e = s.n;
  while (e > b) {
    /* utf8.DecodeLastRune: step back over continuation bytes (max 3) */
    size_t k = e - 1;
    size_t lim = (e - b > 4) ? e - 4 : b;
    while (k > lim && ((unsigned char)s.p[k] & 0xC0) == 0x80) k--;
    uint32_t r;
    int w = go_decode_rune((const unsigned char*)s.p + k, e - k, &r);
    if (k + (size_t)w != e) { r = 0xFFFD; k = e - 1; } /* invalid tail: 1 byte */
    if (!go_is_space(r)) break;
    e = k;
  }
  str_t o = {s.p + b, e - b};
  return o;
}

static int str_has_prefix(str_t s, const char* pre) {
  size_t m = strlen(pre);
  return s.n >= m && memcmp(s.p, pre, m) == 0;
}
static int str_eq(str_t s, const char* lit) {
  size_t m = strlen(lit);
  return s.n == m && memcmp(s.p, lit, m) == 0;
}
static long str_index_byte(str_t s, char c) {
  const char* q = s.n ? memchr(s.p, c, s.n) : NULL;
  return q ? (long)(q - s.p) : -1;
}

/* strconv.Atoi: [+-]?[0-9]+ base 10, no underscores, error on int64 overflow.
 * returns 0 ok, 1 syntax error, 2 range error. */
static int go_atoi(str_t s, int64_t* out) {
  size_t i = 0;
  int neg = 0;
  if (s.n == 0) return 1;
  if (s.p[0] == '+' || s.p[0] == '-') { neg = s.p[0] == '-'; i = 1; }
  if (i == s.n) return 1;
  uint64_t v = 0;
  int range = 0;
  for (; i < s.n; i++) {
    char c = s.p[i];
    if (c < '0' || c > '9') return 1;
    if (!range) {
      if (v > (UINT64_MAX - (uint64_t)(c - '0')) / 10) range = 1;
      else v = v * 10 + (uint64_t)(c - '0');
    }
  }
  if (range) return 2;
  if (!neg && v > (uint64_t)INT64_MAX) return 2;
  if (neg && v > (uint64_t)INT64_MAX + 1) return 2;
  *out = neg ? (int64_t)(0 - v) : (int64_t)v;
  return 0;
}

/* ======================================================================== */
/* robfig/cron v3.0.1 parser.go                                              */
/* ======================================================================== */

typedef struct { uint64_t min, max; const char* const* names; int nnames; uint64_t name_base; } bounds_t;

static const char* const MONTH_NAMES[] = {"jan", "feb", "mar", "apr", "may", "jun",
                                          "jul", "aug", "sep", "oct", "nov", "dec"};
static const char* const DOW_NAMES[] = {"sun", "mon", "tue", "wed", "thu", "fri", "sat"};
static const bounds_t B_SECONDS = {0, 59, NULL, 0, 0};
static const bounds_t B_MINUTES = {0, 59, NULL, 0, 0};
static const bounds_t B_HOURS = {0, 23, NULL, 0, 0};
static const bounds_t B_DOM = {1, 31, NULL, 0, 0};
static const bounds_t B_MONTHS = {1, 12, MONTH_NAMES, 12, 1};
static const bounds_t B_DOW = {0, 6, DOW_NAMES, 7, 0};

typedef struct { char* buf; size_t cap; } errbuf_t;
static void set_err(errbuf_t* e, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
#include <stdarg.h>
static void set_err(errbuf_t* e, const char* fmt, ...) {
  if (!e || !e->buf || e->cap == 0) return;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(e->buf, e->cap, fmt, ap);
  va_end(ap);
}

/* strings.ToLower restricted to what can land on an ASCII name: ASCII letters,
 * plus U+0130 (I with dot) whose simple lower-case mapping is 'i' and U+212A
 * (Kelvin) -> 'k'.  Anything else non-ASCII can never equal a name. */
static int lower_name(str_t s, char* out, size_t cap) {
  size_t i = 0, o = 0;
  while (i < s.n) {
    uint32_t r;
    int w = go_decode_rune((const unsigned char*)s.p + i, s.n - i, &r);
    if (r < 0x80) {
      if (o + 1 >= cap) return -1;
      out[o++] = (r >= 'A' && r <= 'Z') ? (char)(r + 32) : (char)r;
    } else if (r == 0x130) {
      if (o + 1 >= cap) return -1;
      out[o++] = 'i';
    } else if (r == 0x212A) {
      if (o + 1 >= cap) return -1;
      out[o++] = 'k';
    } else {
      return -1;
    }
    i += (size_t)w;
  }
  out[o] = 0;
  return (int)o;
}

/* mustParseInt (parser.go): Atoi, then reject negatives */
static int must_parse_int(str_t expr, uint64_t* out, errbuf_t* e) {
  int64_t num;
  int rc = go_atoi(expr, &num);
  if (rc != 0) {
    set_err(e, "failed to parse int from %.*s: strconv.Atoi: parsing \"%.*s\": %s", (int)expr.n,
            expr.p, (int)expr.n, expr.p, rc == 1 ? "invalid syntax" : "value out of range");
    return -1;
  }
  if (num < 0) {
    set_err(e, "negative number (%lld) not allowed: %.*s", (long long)num, (int)expr.n, expr.p);
    return -1;
  }
  *out = (uint64_t)num;
  return 0;
}

/* parseIntOrName (parser.go) */
static int parse_int_or_name(str_t expr, const bounds_t* r, uint64_t* out, errbuf_t* e) {
  if (r->names != NULL) {
    char low[8];
    int n = lower_name(expr, low, sizeof low);
    if (n == 3) {
      for (int i = 0; i < r->nnames; i++) {
        if (memcmp(low, r->names[i], 3) == 0) { *out = r->name_base + (uint64_t)i; return 0; }
      }
    }
  }
  return must_parse_int(expr, out, e);
}

/* getBits (parser.go) */
static uint64_t get_bits(uint64_t min, uint64_t max, uint64_t step) {
  if (step == 1) return ~(UINT64_MAX << (max + 1)) & (UINT64_MAX << min);
  uint64_t bits = 0;
  for (uint64_t i = min; i <= max; i += step) bits |= 1ull << i;
  return bits;
}
static uint64_t all_bits(const bounds_t* r) { return get_bits(r->min, r->max, 1) | STAR_BIT; }

/* strings.Split(s, sep) for a one-byte separator, at most cap pieces kept;
 * returns the true piece count. */
static size_t split_byte(str_t s, char sep, str_t* out, size_t cap) {
  size_t count = 0, start = 0;
  for (size_t i = 0; i <= s.n; i++) {
    if (i == s.n || s.p[i] == sep) {
      if (count < cap) { out[count].p = s.p + start; out[count].n = i - start; }
      count++;
      start = i + 1;
    }
  }
  return count;
}

/* getRange (parser.go): number | number "-" number [ "/" number ] */
static int get_range(str_t expr, const bounds_t* r, uint64_t* bits_out, errbuf_t* e) {
  uint64_t start = 0, end = 0, step = 0;
  str_t range_and_step[2], low_and_high[2];
  size_t n_rs = split_byte(expr, '/', range_and_step, 2);
  size_t n_lh = split_byte(range_and_step[0], '-', low_and_high, 2);
  int single_digit = (n_lh == 1);
  uint64_t extra = 0;

  if (str_eq(low_and_high[0], "*") || str_eq(low_and_high[0], "?")) {
    start = r->min;
    end = r->max;
    extra = STAR_BIT;
  } else {
    if (parse_int_or_name(low_and_high[0], r, &start, e)) return -1;
    switch (n_lh) {
      case 1: end = start; break;
      case 2:
        if (parse_int_or_name(low_and_high[1], r, &end, e)) return -1;
        break;
      default:
        set_err(e, "too many hyphens: %.*s", (int)expr.n, expr.p);
        return -1;
    }
  }

  switch (n_rs) {
    case 1: step = 1; break;
    case 2:
      if (must_parse_int(range_and_step[1], &step, e)) return -1;
      /* Special handling: "N/step" means "N-max/step". */
      if (single_digit) end = r->max;
      if (step > 1) extra = 0;
      break;
    default:
      set_err(e, "too many slashes: %.*s", (int)expr.n, expr.p);
      return -1;
  }

  if (start < r->min) {
    set_err(e, "beginning of range (%llu) below minimum (%llu): %.*s", (unsigned long long)start,
            (unsigned long long)r->min, (int)expr.n, expr.p);
    return -1;
  }
  if (end > r->max) {
    set_err(e, "end of range (%llu) above maximum (%llu): %.*s", (unsigned long long)end,
            (unsigned long long)r->max, (int)expr.n, expr.p);
    return -1;
  }
  if (start > end) {
    set_err(e, "beginning of range (%llu) beyond end of range (%llu): %.*s",
            (unsigned long long)start, (unsigned long long)end, (int)expr.n, expr.p);
    return -1;
  }
  if (step == 0) {
    set_err(e, "step of range should be a positive number: %.*s", (int)expr.n, expr.p);
    return -1;
  }
  *bits_out = get_bits(start, end, step) | extra;
  return 0;
}

/* getField (parser.go): comma-separated ranges, empty pieces dropped
 * (strings.FieldsFunc) */
static int get_field(str_t field, const bounds_t* r, uint64_t* bits_out, errbuf_t* e) {
  uint64_t bits = 0;
  size_t start = 0;
  for (size_t i = 0; i <= field.n; i++) {
    if (i == field.n || field.p[i] == ',') {
      if (i > start) {
        str_t expr = {field.p + start, i - start};
        uint64_t bit;
        if (get_range(expr, r, &bit, e)) return -1;
        bits |= bit;
      }
      start = i + 1;
    }
  }
  *bits_out = bits;
  return 0;
}

/* ---- time.ParseDuration (Go standard library) --------------------------- */
static int leading_int(str_t* s, uint64_t* x) {
  size_t i = 0;
  uint64_t v = 0;
  for (; i < s->n; i++) {
    char c = s->p[i];
    if (c < '0' || c > '9') break;
    if (v > (1ull << 63) / 10) return -1;
    v = v * 10 + (uint64_t)(c - '0');
    if (v > (1ull << 63)) return -1;
  }
  *x = v;
  s->p += i;
  s->n -= i;
  return 0;
}
static void leading_fraction(str_t* s, uint64_t* x, double* scale) {
  size_t i = 0;
  uint64_t v = 0;
  double sc = 1;
  int overflow = 0;
  for (; i < s->n; i++) {
    char c = s->p[i];
    if (c < '0' || c > '9') break;
    if (overflow) continue;
    if (v > ((1ull << 63) - 1) / 10) { overflow = 1; continue; }
    uint64_t y = v * 10 + (uint64_t)(c - '0');
    if (y > (1ull << 63)) { overflow = 1; continue; }
    v = y;
    sc *= 10;
  }
  *x = v;
  *scale = sc;
  s->p += i;
  s->n -= i;
}
static int unit_ns(str_t u, uint64_t* unit) {
  if (str_eq(u, "ns")) { *unit = 1ull; return 0; }
  if (str_eq(u, "us")) { *unit = 1000ull; return 0; }
  if (str_eq(u, "\xC2\xB5s")) { *unit = 1000ull; return 0; } /* U+00B5 micro */
  if (str_eq(u, "\xCE\xBCs")) { *unit = 1000ull; return 0; } /* U+03BC mu    */
  if (str_eq(u, "ms")) { *unit = 1000000ull; return 0; }
  if (str_eq(u, "s")) { *unit = 1000000000ull; return 0; }
  if (str_eq(u, "m")) { *unit = 60ull * 1000000000ull; return 0; }
  if (str_eq(u, "h")) { *unit = 3600ull * 1000000000ull; return 0; }
  return -1;
}
int orc_parse_duration(const char* sp, size_t len, int64_t* ns_out) {
  str_t s = {sp, len};
  uint64_t d = 0;
  int neg = 0;
  if (s.n != 0) {
    char c = s.p[0];
    if (c == '-' || c == '+') { neg = (c == '-'); s.p++; s.n--; }
  }
  if (str_eq(s, "0")) { *ns_out = 0; return 0; }
  if (s.n == 0) return -1;
  while (s.n != 0) {
    uint64_t v = 0, f = 0;
    double scale = 1;
    if (!(s.p[0] == '.' || (s.p[0] >= '0' && s.p[0] <= '9'))) return -1;
    size_t pl = s.n;
    if (leading_int(&s, &v)) return -1;
    int pre = (pl != s.n);
    int post = 0;
    if (s.n != 0 && s.p[0] == '.') {
      s.p++; s.n--;
      size_t pl2 = s.n;
      leading_fraction(&s, &f, &scale);
      post = (pl2 != s.n);
    }
    if (!pre && !post) return -1;
    size_t i = 0;
    for (; i < s.n; i++) {
      char c = s.p[i];
      if (c == '.' || (c >= '0' && c <= '9')) break;
    }
    if (i == 0) return -1; /* missing unit */
    str_t u = {s.p, i};
    s.p += i; s.n -= i;
    uint64_t unit;
    if (unit_ns(u, &unit)) return -1; /* unknown unit */
    if (v > (1ull << 63) / unit) return -1;
    v *= unit;
    if (f > 0) {
      v += (uint64_t)((double)f * ((double)unit / scale));
      if (v > (1ull << 63)) return -1;
    }
    d += v;
    if (d > (1ull << 63)) return -1;
  }
  if (neg) { *ns_out = (int64_t)(0 - d); return 0; }
  if (d > (1ull << 63) - 1) return -1;
  *ns_out = (int64_t)d;
  return 0;
}

/* Every (constantdelay.go): <1s rounds up to 1s; sub-second part truncated */
static int64_t every_delay_sec(int64_t dur_ns) {
  const int64_t second = 1000000000ll;
  if (dur_ns < second) dur_ns = second;
  return (dur_ns - dur_ns % second) / second;
}

/* parseDescriptor (parser.go) */
static int parse_descriptor(str_t d, orc_cron_t* out, errbuf_t* e) {
  out->kind = CRON_SPEC;
  uint64_t s0 = 1ull << B_MINUTES.min, h0 = 1ull << B_HOURS.min;
  if (str_eq(d, "@yearly") || str_eq(d, "@annually")) {
    out->minute = s0; out->hour = h0; out->dom = 1ull << B_DOM.min;
    out->month = 1ull << B_MONTHS.min; out->dow = all_bits(&B_DOW);
    return 0;
  }
  if (str_eq(d, "@monthly")) {
    out->minute = s0; out->hour = h0; out->dom = 1ull << B_DOM.min;
    out->month = all_bits(&B_MONTHS); out->dow = all_bits(&B_DOW);
    return 0;
  }
  if (str_eq(d, "@weekly")) {
    out->minute = s0; out->hour = h0; out->dom = all_bits(&B_DOM);
    out->month = all_bits(&B_MONTHS); out->dow = 1ull << B_DOW.min;
    return 0;
  }
  if (str_eq(d, "@daily") || str_eq(d, "@midnight")) {
    out->minute = s0; out->hour = h0; out->dom = all_bits(&B_DOM);
    out->month = all_bits(&B_MONTHS); out->dow = all_bits(&B_DOW);
    return 0;
  }
  if (str_eq(d, "@hourly")) {
    out->minute = s0; out->hour = all_bits(&B_HOURS); out->dom = all_bits(&B_DOM);
    out->month = all_bits(&B_MONTHS); out->dow = all_bits(&B_DOW);
    return 0;
  }
  if (str_has_prefix(d, "@every ")) {
    int64_t ns;
    if (orc_parse_duration(d.p + 7, d.n - 7, &ns)) {
      out->kind = CRON_ERROR;
      set_err(e, "failed to parse duration %.*s", (int)d.n, d.p);
      return -1;
    }
    out->kind = CRON_EVERY;
    out->delay_sec = every_delay_sec(ns);
    return 0;
  }
  out->kind = CRON_ERROR;
  set_err(e, "unrecognized descriptor: %.*s", (int)d.n, d.p);
  return -1;
}

/* ---- named time zones (time.LoadLocation, robfig parser.go) ----------------
 * Ids are handed out in order of first appearance (1..255), like the product's,
 * but nothing else is shared with it: the product reads TZif files itself
 * (csrc/tz.cpp); the oracle asks libc — setenv("TZ") + tzset() + localtime_r
 * under a lock, the process-global way. */
#define ORC_MAX_ZONES 255
static pthread_mutex_t g_tz_mu = PTHREAD_MUTEX_INITIALIZER;
static char g_tz_names[ORC_MAX_ZONES + 1][256];
static int g_tz_count = 0;

static int tz_file_exists(const char* name) {
  const char* dirs[] = {getenv("ZONEINFO"), "/usr/share/zoneinfo", "/usr/share/lib/zoneinfo", "/usr/lib/locale/TZ",
                        "/etc/zoneinfo"};
  for (size_t k = 0; k < sizeof dirs / sizeof dirs[0]; k++) {
    if (!dirs[k]) continue;
    char path[768];
    snprintf(path, sizeof path, "%s/%s", dirs[k], name);
    FILE* f = fopen(path, "rb");
    if (!f) continue;
    char magic[4] = {0, 0, 0, 0};
    size_t n = fread(magic, 1, 4, f);
    fclose(f);
    if (n == 4 && memcmp(magic, "TZif", 4) == 0) return 1;
  }
  return 0;
}

/* 0 ok (*id set; "", "UTC", "Local" -> 0), -6 unknown zone, -7 table full */
int orc_tz_lookup(const char* name, size_t len, int32_t* id) {
  if (len == 0 || (len == 3 && memcmp(name, "UTC", 3) == 0) || (len == 5 && memcmp(name, "Local", 5) == 0)) {
    *id = 0;
    return 0;
  }
  if (len > 255 || name[0] == '/' || name[0] == '\\') return E_PARSE;
  char buf[256];
  memcpy(buf, name, len);
  buf[len] = 0;
  if (strstr(buf, "..")) return E_PARSE;
  pthread_mutex_lock(&g_tz_mu);
  int rc = 0;
  int found = 0;
  for (int k = 1; k <= g_tz_count; k++)
    if (strcmp(g_tz_names[k], buf) == 0) { *id = k; found = 1; break; }
  if (!found) {
    if (!tz_file_exists(buf)) rc = E_PARSE;
    else if (g_tz_count >= ORC_MAX_ZONES) rc = E_UNSUPPORTED;
    else { strcpy(g_tz_names[++g_tz_count], buf); *id = g_tz_count; }
  }
  pthread_mutex_unlock(&g_tz_mu);
  return rc;
}

/* T's broken-down LOCAL time in every registered zone (entry 0 = UTC), through libc */
static int zone_tms(int64_t T, struct tm* out /* [ORC_MAX_ZONES + 1] */) {
  time_t tt = (time_t)T;
  gmtime_r(&tt, &out[0]);
  pthread_mutex_lock(&g_tz_mu);
  const int n = g_tz_count;
  if (n > 0) {
    char* old = getenv("TZ");
    char saved[300];
    if (old) snprintf(saved, sizeof saved, "%s", old);
    for (int k = 1; k <= n; k++) {
      char v[300];
      snprintf(v, sizeof v, ":%s", g_tz_names[k]);
      setenv("TZ", v, 1);
      tzset();
      localtime_r(&tt, &out[k]);
    }
    if (old) setenv("TZ", saved, 1); else unsetenv("TZ");
    tzset();
  }
  pthread_mutex_unlock(&g_tz_mu);
  return n;
}

int orc_tz_offset(int32_t id, int64_t T, int32_t* utoff) {
  struct tm tms[ORC_MAX_ZONES + 1];
  const int n = zone_tms(T, tms);
  if (id < 0 || id > n) return E_INVAL;
  *utoff = id ? (int32_t)tms[id].tm_gmtoff : 0;
  return 0;
}

/* Parser.Parse with options Minute|Hour|Dom|Month|Dow|Descriptor
 * (= ParseStandard, hcc.go:253) */
int orc_cron_parse(const char* spec_p, size_t len, orc_cron_t* out, char* err, size_t errcap) {
  errbuf_t e = {err, errcap};
  str_t spec = {spec_p, len};
  memset(out, 0, sizeof *out);
  if (err && errcap) err[0] = 0;
  if (spec.n == 0) { set_err(&e, "empty spec string"); return E_PARSE; }
  int32_t tz_id = 0;

  /* Extract timezone if present */
  if (str_has_prefix(spec, "TZ=") || str_has_prefix(spec, "CRON_TZ=")) {
    long i = str_index_byte(spec, ' ');
    long eq = str_index_byte(spec, '=');
    if (i < 0) {
      /* v3.0.1 slices spec[eq+1:-1] here and panics; Reconcile recovers the
       * panic (hcc.go:191-195) and the check never runs: an error to us. */
      set_err(&e, "provided bad location (no space after TZ=): runtime panic in robfig v3.0.1");
      return E_PARSE;
    }
    str_t loc = {spec.p + eq + 1, (size_t)(i - eq - 1)};
    str_t rest = {spec.p + i, spec.n - (size_t)i};
    spec = go_trim_space(rest);
    /* loc, err = time.LoadLocation(...): "provided bad location %s: %v" */
    int trc = orc_tz_lookup(loc.p, loc.n, &tz_id);
    if (trc == E_PARSE) {
      set_err(&e, "provided bad location %.*s: unknown time zone %.*s", (int)loc.n, loc.p, (int)loc.n, loc.p);
      return E_PARSE;
    }
    if (trc != 0) {
      set_err(&e, "time zone %.*s: more than %d distinct zones", (int)loc.n, loc.p, ORC_MAX_ZONES);
      return E_UNSUPPORTED;
    }
  }

  /* Handle named schedules (descriptors) */
  if (str_has_prefix(spec, "@")) {
    if (parse_descriptor(spec, out, &e)) { memset(out, 0, sizeof *out); return E_PARSE; }
    if (out->kind == CRON_SPEC) out->tz_id = tz_id; /* SpecSchedule.Location; a ConstantDelaySchedule has none */
    return 0;
  }

  /* Split on whitespace; normalizeFields: exactly 5, seconds "0" prepended */
  str_t fields[6];
  size_t count = go_fields(spec, fields, 6);
  if (count != 5) {
    set_err(&e, "expected exactly 5 fields, found %zu: [%.*s]", count, (int)spec.n, spec.p);
    return E_PARSE;
  }
  uint64_t second, minute, hour, dom, month, dow;
  str_t zero = {"0", 1};
  if (get_field(zero, &B_SECONDS, &second, &e)) return E_PARSE;
  if (get_field(fields[0], &B_MINUTES, &minute, &e)) return E_PARSE;
  if (get_field(fields[1], &B_HOURS, &hour, &e)) return E_PARSE;
  if (get_field(fields[2], &B_DOM, &dom, &e)) return E_PARSE;
  if (get_field(fields[3], &B_MONTHS, &month, &e)) return E_PARSE;
  if (get_field(fields[4], &B_DOW, &dow, &e)) return E_PARSE;
  (void)second; /* always 1<<0: the reason matches() requires sec == 0 */
  out->kind = CRON_SPEC;
  out->minute = minute; out->hour = hour; out->dom = dom; out->month = month; out->dow = dow;
  out->tz_id = tz_id; /* SpecSchedule.Location */
  return 0;
}

/* ======================================================================== */
/* robfig/cron v3.0.1 spec.go, in UTC                                        */
/* ======================================================================== */

/* Set (under g_tz_mu, with TZ pointing at the zone) while Next() walks a zone-bound schedule:
 * the two calendar helpers then work on the zone's wall clock, as Go's time package does for a
 * time.Time carrying that Location. */
static __thread int tl_in_zone = 0;

static void utc_tm(int64_t t, struct tm* tm) {
  time_t tt = (time_t)t;
  if (tl_in_zone) localtime_r(&tt, tm);
  else gmtime_r(&tt, tm);
}

void orc_civil_from_unix(int64_t unix_sec, int32_t out[6]) {
  struct tm tm;
  utc_tm(unix_sec, &tm);
  out[0] = tm.tm_sec; out[1] = tm.tm_min; out[2] = tm.tm_hour;
  out[3] = tm.tm_mday; out[4] = tm.tm_mon + 1; out[5] = tm.tm_wday;
}

/* dayMatches (spec.go) */
static int day_matches(const orc_cron_t* s, const struct tm* t) {
  int dom_match = ((1ull << (unsigned)t->tm_mday) & s->dom) > 0;
  int dow_match = ((1ull << (unsigned)t->tm_wday) & s->dow) > 0;
  if ((s->dom & STAR_BIT) > 0 || (s->dow & STAR_BIT) > 0) return dom_match && dow_match;
  return dom_match || dow_match;
}

/* the whole-second activation predicate, given T's UTC fields */
static int matches_tm(const orc_cron_t* c, const struct tm* t) {
  if (c->kind != CRON_SPEC) return 0;
  if (t->tm_sec != 0) return 0; /* Second mask is 1<<0 */
  if (((1ull << (unsigned)t->tm_min) & c->minute) == 0) return 0;
  if (((1ull << (unsigned)t->tm_hour) & c->hour) == 0) return 0;
  if (((1ull << (unsigned)(t->tm_mon + 1)) & c->month) == 0) return 0;
  return day_matches(c, t);
}
int orc_cron_matches(const orc_cron_t* c, int64_t T) {
  if (c->tz_id) { /* the schedule's Location: T's wall clock there */
    struct tm tms[ORC_MAX_ZONES + 1];
    const int n = zone_tms(T, tms);
    return c->tz_id <= n ? matches_tm(c, &tms[c->tz_id]) : 0;
  }
  struct tm t;
  utc_tm(T, &t);
  return matches_tm(c, &t);
}

static int64_t tm_date(int year, int mon1, int mday, int hh, int mm, int ss) {
  struct tm tm;
  memset(&tm, 0, sizeof tm);
  tm.tm_year = year - 1900; tm.tm_mon = mon1 - 1; tm.tm_mday = mday;
  tm.tm_hour = hh; tm.tm_min = mm; tm.tm_sec = ss;
  if (tl_in_zone) {
    tm.tm_isdst = -1;
    return (int64_t)mktime(&tm); /* time.Date(..., loc) */
  }
  return (int64_t)timegm(&tm); /* normalises like time.Date / AddDate */
}

/* SpecSchedule.Next (spec.go) for a whole-second t in UTC; the zero time
 * (nothing within five years) is reported as INT64_MIN.
 * ConstantDelaySchedule.Next (constantdelay.go) = t + Delay. */
static int64_t cron_next_in_current_zone(const orc_cron_t* s, int64_t t0);

int64_t orc_cron_next(const orc_cron_t* s, int64_t t0) {
  if (s->kind == CRON_EVERY) return t0 + s->delay_sec;
  if (s->kind != CRON_SPEC) return INT64_MIN;
  if (s->tz_id == 0) return cron_next_in_current_zone(s, t0);
  /* t = t.In(s.Location): the same walk on the zone's wall clock, libc's tz database behind it */
  pthread_mutex_lock(&g_tz_mu);
  int64_t r = INT64_MIN;
  if (s->tz_id <= g_tz_count) {
    char* old = getenv("TZ");
    char saved[300], v[300];
    if (old) snprintf(saved, sizeof saved, "%s", old);
    snprintf(v, sizeof v, ":%s", g_tz_names[s->tz_id]);
    setenv("TZ", v, 1);
    tzset();
    tl_in_zone = 1;
    r = cron_next_in_current_zone(s, t0);
    tl_in_zone = 0;
    if (old) setenv("TZ", saved, 1); else unsetenv("TZ");
    tzset();
  }
  pthread_mutex_unlock(&g_tz_mu);
  return r;
}

static int64_t cron_next_in_current_zone(const orc_cron_t* s, int64_t t0) {
  /* Start at the earliest possible time (the upcoming second). */
  int64_t t = t0 + 1;
  int added = 0;
  struct tm tm;
  utc_tm(t, &tm);
  int year_limit = tm.tm_year + 1900 + 5;
  const uint64_t second_mask = 1ull; /* ParseStandard prepends "0" */

WRAP:
  utc_tm(t, &tm);
  if (tm.tm_year + 1900 > year_limit) return INT64_MIN;

  /* Find the first applicable month. */
  while (((1ull << (unsigned)(tm.tm_mon + 1)) & s->month) == 0) {
    if (!added) {
      added = 1;
      t = tm_date(tm.tm_year + 1900, tm.tm_mon + 1, 1, 0, 0, 0);
      utc_tm(t, &tm);
    }
    t = tm_date(tm.tm_year + 1900, tm.tm_mon + 2, tm.tm_mday, tm.tm_hour, tm.tm_min, tm.tm_sec);
    utc_tm(t, &tm);
    if (tm.tm_mon == 0) goto WRAP; /* wrapped around to January */
  }

  /* Now get a day in that month. */
  while (!day_matches(s, &tm)) {
    if (!added) {
      added = 1;
      t = tm_date(tm.tm_year + 1900, tm.tm_mon + 1, tm.tm_mday, 0, 0, 0);
      utc_tm(t, &tm);
    }
    t = tm_date(tm.tm_year + 1900, tm.tm_mon + 1, tm.tm_mday + 1, tm.tm_hour, tm.tm_min, tm.tm_sec);
    utc_tm(t, &tm);
    /* "Notice if the hour is no longer midnight due to DST.  Add an hour if it's 23, subtract an
     * hour if it's 1." (spec.go) — never in UTC */
    if (tm.tm_hour != 0) {
      if (tm.tm_hour > 12) t += (int64_t)(24 - tm.tm_hour) * 3600;
      else t -= (int64_t)tm.tm_hour * 3600;
      utc_tm(t, &tm);
    }
    if (tm.tm_mday == 1) goto WRAP;
  }

  while (((1ull << (unsigned)tm.tm_hour) & s->hour) == 0) {
    if (!added) {
      added = 1;
      t = tm_date(tm.tm_year + 1900, tm.tm_mon + 1, tm.tm_mday, tm.tm_hour, 0, 0);
    }
    t += 3600;
    utc_tm(t, &tm);
    if (tm.tm_hour == 0) goto WRAP;
  }

  while (((1ull << (unsigned)tm.tm_min) & s->minute) == 0) {
    if (!added) {
      added = 1;
      t -= tm.tm_sec; /* t.Truncate(time.Minute) */
    }
    t += 60;
    utc_tm(t, &tm);
    if (tm.tm_min == 0) goto WRAP;
  }

  while (((1ull << (unsigned)tm.tm_sec) & second_mask) == 0) {
    if (!added) added = 1; /* t.Truncate(time.Second): whole seconds already */
    t += 1;
    utc_tm(t, &tm);
    if (tm.tm_sec == 0) goto WRAP;
  }
  return t;
}

/* hcc.go:262  RepeatAfterSec = int(Next(now).Sub(now')/time.Second) + 1 for a
 * real clock (0 < ns): equals Next(floor(now)) - floor(now).
 * When Next is the zero time, Sub saturates at minDuration and the Go value
 * is int(-9223372036) + 1. */
int64_t orc_cron_repeat_after_sec(const orc_cron_t* c, int64_t unix_sec) {
  int64_t nx = orc_cron_next(c, unix_sec);
  if (nx == INT64_MIN) return -9223372036ll + 1;
  return nx - unix_sec;
}

/* ======================================================================== */
/* hcc.go ladder at upsert and api/v1alpha1 helpers              */
/* ======================================================================== */

int orc_remedy_is_empty(size_t generate_name_len, int resource_is_nil, int64_t timeout,
                        int rbac_rules_is_nil) {
  /* healthcheck_types.go:104-106 reflect.DeepEqual(w, RemedyWorkflow{}) */
  return generate_name_len == 0 && resource_is_nil && timeout == 0 && rbac_rules_is_nil;
}

static int fits_i32(int64_t v) { return v >= INT32_MIN && v <= INT32_MAX; }
#define TIME_LIMIT (1ll << 55)

int orc_classify(const orc_healthcheck_t* hc, orc_record_t* out) {
  memset(out, 0, sizeof *out);
  /* column domain */
  if (!fits_i32(hc->remedy_runs_limit) || !fits_i32(hc->remedy_reset_interval)) return E_RANGE;
  if (!fits_i32(hc->success_count) || !fits_i32(hc->failed_count) ||
      !fits_i32(hc->remedy_success_count) || !fits_i32(hc->remedy_failed_count) ||
      !fits_i32(hc->remedy_total_runs))
    return E_RANGE;
  if (hc->finished_at_set && (hc->finished_at >= TIME_LIMIT || hc->finished_at <= -TIME_LIMIT))
    return E_RANGE;
  if (hc->remedy_finished_at_set &&
      (hc->remedy_finished_at >= TIME_LIMIT || hc->remedy_finished_at <= -TIME_LIMIT ||
       hc->remedy_finished_at == 0 /* collides with the nil sentinel */))
    return E_RANGE;
  if (hc->fail_p8 > 255) return E_RANGE;

  uint32_t flags = 0;
  int32_t ras = 0;
  int rc = 0;
  if (!hc->has_resource) { /* hcc.go:227 */
    flags = KIND_NO_RESOURCE;
  } else if (hc->repeat_after_sec <= 0 && hc->cron_len == 0) { /* hcc.go:238 */
    flags = KIND_STOPPED;
  } else if (hc->repeat_after_sec <= 0 && hc->cron_len != 0) { /* hcc.go:251 */
    orc_cron_t c;
    int prc = orc_cron_parse(hc->cron, hc->cron_len, &c, NULL, 0);
    if (prc == E_UNSUPPORTED) {
      flags = KIND_HOST_FALLBACK;
      rc = E_UNSUPPORTED;
    } else if (prc != 0) { /* hcc.go:254-257 */
      flags = KIND_PARSE_ERROR;
    } else if (c.kind == CRON_EVERY) {
      if (!fits_i32(c.delay_sec)) return E_RANGE;
      flags = KIND_CRON_EVERY;
      ras = (int32_t)c.delay_sec; /* hcc.go:262 with N2 */
    } else {
      flags = KIND_CRON_SPEC | ((uint32_t)c.tz_id << F_TZ_SHIFT); /* SpecSchedule.Location */
      out->minute = c.minute; out->hour = c.hour; out->dom = c.dom;
      out->month = c.month; out->dow = c.dow;
    }
  } else { /* hcc.go:264 and the final else: RepeatAfterSec > 0, cron ignored */
    if (!fits_i32(hc->repeat_after_sec)) return E_RANGE;
    flags = KIND_INTERVAL;
    ras = (int32_t)hc->repeat_after_sec;
  }
  if (hc->has_remedy) flags |= F_HAS_REMEDY;
  flags |= hc->fail_p8 << F_FAILP_SHIFT;
  if (hc->timer_armed) flags |= F_TIMER_ARMED; /* hcc.go:264: "&& r.GetTimerByName(...) != nil" */
  out->flags = flags;
  out->ras = ras;
  out->finished_at = hc->finished_at_set ? hc->finished_at : 0; /* hcc.go:231-235 */
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

/* ======================================================================== */
/* per-tick function                                            */
/* ======================================================================== */

static inline uint64_t sm64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
uint64_t orc_key(uint64_t seed, uint64_t i, uint64_t f) { return sm64(sm64(seed ^ sm64(i)) + f); }

/* int(now.Time.Sub(RemedyFinishedAt.Time).Seconds()) (hcc.go:690): Sub
 * saturates at +-(1<<63-1) ns, Seconds() is float64, int() truncates. */
static int64_t go_sub_seconds(int64_t now, int64_t then) {
  int64_t d = (int64_t)((uint64_t)now - (uint64_t)then); /* |values| < 2^55: exact */
  if (d > 9223372036ll) return 9223372036ll;
  if (d < -9223372036ll) return -9223372036ll;
  return d;
}

/* watchRemedyWorkflow result (hcc.go:821-851) */
static void apply_remedy_result(orc_record_t* r, int64_t T, int ok, orc_tick_stats_t* st) {
  if (ok) { r->remedy_success = (int32_t)((uint32_t)r->remedy_success + 1u); if (st) st->n_remedy_ok++; }
  else    { r->remedy_failed = (int32_t)((uint32_t)r->remedy_failed + 1u); if (st) st->n_remedy_fail++; }
  r->remedy_total = (int32_t)((uint32_t)r->remedy_success + (uint32_t)r->remedy_failed); /* :829/:845 */
  r->remedy_finished_at = T;                                                            /* :825/:840 */
}

static void reset_remedy(orc_record_t* r) { /* hcc.go:651-656 and :695-700 */
  r->remedy_total = 0;
  r->remedy_success = 0;
  r->remedy_failed = 0;
  r->remedy_finished_at = 0; /* nil */
}

/* step 1: apply a posted workflow result (hcc.go:633-724, 819-852) */
static uint32_t apply_result(orc_record_t* r, int64_t T, orc_tick_stats_t* st) {
  uint32_t f = r->flags, act = 0;
  if (f & F_PENDING_OK) { /* hcc.go:635-661 */
    r->success = (int32_t)((uint32_t)r->success + 1u);
    r->finished_at = T;
    if (st) st->n_result_ok++;
    if ((f & F_HAS_REMEDY) && r->remedy_total >= 1) { /* :649 */
      reset_remedy(r);
      act |= ACT_RESET_ON_PASS;
    }
  } else if (f & F_PENDING_FAIL) { /* hcc.go:662-722 */
    r->failed = (int32_t)((uint32_t)r->failed + 1u);
    r->finished_at = T;
    if (st) st->n_result_fail++;
    if (f & F_HAS_REMEDY) { /* :677 */
      int run = 0;
      if (r->runs_limit != 0 && r->reset_interval != 0) { /* :679 */
        if (r->runs_limit > r->remedy_total) { /* :681 */
          run = 1;
        } else if (r->remedy_finished_at == 0) {
          act |= ACT_ANOMALY; /* nil deref at :690 (N3) */
        } else {
          int64_t d = go_sub_seconds(T, r->remedy_finished_at); /* :690 */
          if ((int64_t)r->reset_interval >= d) { /* :692 */
            act |= ACT_REMEDY_SKIP;
          } else { /* :695-704 */
            reset_remedy(r);
            act |= ACT_RESET_ON_INTERVAL;
            run = 1;
          }
        }
      } else { /* :712-719 */
        run = 1;
      }
      if (run) {
        act |= ACT_RUN_REMEDY;
        if (f & F_REMEDY_PENDING) apply_remedy_result(r, T, (f & F_REMEDY_OUTCOME_OK) != 0, st);
      }
    }
  } else if (f & F_REMEDY_PENDING) { /* remedy finished on its own: hcc.go:821-851 */
    apply_remedy_result(r, T, (f & F_REMEDY_OUTCOME_OK) != 0, st);
  }
  r->flags = f & ~(F_PENDING_OK | F_PENDING_FAIL | F_REMEDY_PENDING | F_REMEDY_OUTCOME_OK);
  /* watchWorkflowReschedule re-arms the repeat timer after either outcome: hcc.go:745-752 */
  if (f & (F_PENDING_OK | F_PENDING_FAIL)) r->flags |= F_TIMER_ARMED;
  return act;
}

/* `tmT[z]` = T's broken-down time in zone z (0 = UTC), computed once per tick by the sweeps (glibc's gmtime_r
 * takes a process-wide lock, which would serialise the threaded baseline) */
static uint32_t tick_record_tm(orc_record_t* r, int64_t T, const struct tm* tmT, uint32_t mode,
                               uint64_t seed, uint64_t gidx, orc_tick_stats_t* st) {
  uint32_t kind = r->flags & KIND_MASK;
  if (r->flags & F_TOMBSTONE) return 0;
  if (kind == KIND_NO_RESOURCE || kind == KIND_HOST_FALLBACK || kind > KIND_HOST_FALLBACK)
    return 0; /* hcc.go:227/290 */

  uint32_t act = apply_result(r, T, st);

  int due = 0;
  switch (kind) {
    case KIND_STOPPED: /* hcc.go:238-250 */
      if (!(r->flags & F_STOPPED_REPORTED)) {
        act |= ACT_STOPPED;
        r->finished_at = T;
        r->flags |= F_STOPPED_REPORTED;
      }
      break;
    case KIND_PARSE_ERROR: /* hcc.go:254-257, requeued after 1 s (:204) */
      act |= ACT_PARSE_ERROR;
      break;
    case KIND_INTERVAL:
    case KIND_CRON_EVERY: { /* not(hcc.go:264)  ==  timer of :751 has fired */
      /* hcc.go:264: skipped iff "now - finishedAt < RepeatAfterSec && timer != nil"; with no
       * timer (controller restart: hcc.go:161 starts with an empty map) the check is submitted */
      int64_t elapsed = (int64_t)((uint64_t)T - (uint64_t)r->finished_at);
      due = !(elapsed < (int64_t)r->ras && (r->flags & F_TIMER_ARMED));
      break;
    }
    case KIND_CRON_SPEC: {
      orc_cron_t c = {r->minute, r->hour, r->dom, r->month, r->dow, 0, CRON_SPEC, 0};
      due = matches_tm(&c, &tmT[r->flags >> F_TZ_SHIFT]); /* T's wall clock in the schedule's Location */
      break;
    }
  }
  if (due) {
    act |= ACT_SUBMIT_HC; /* hcc.go:269-288 */
    if (mode & MODE_CLOSED_LOOP) {
      /* harness convention: the run completes at once */
      uint64_t k = orc_key(seed, gidx, (uint64_t)T);
      uint32_t failp = (r->flags >> F_FAILP_SHIFT) & 0xFFu;
      int fail = (uint32_t)(k & 0xFF) < failp;
      int remedy_ok = (uint32_t)((k >> 8) & 0xFF) < 179u;
      r->flags |= fail ? F_PENDING_FAIL : F_PENDING_OK;
      r->flags |= F_REMEDY_PENDING | (remedy_ok ? F_REMEDY_OUTCOME_OK : 0u);
      act |= apply_result(r, T, st);
    }
  }
  return act;
}

uint32_t orc_tick_record(orc_record_t* r, int64_t T, uint32_t mode, uint64_t seed,
                         uint64_t gidx, orc_tick_stats_t* st) {
  struct tm tmT[ORC_MAX_ZONES + 1];
  zone_tms(T, tmT);
  return tick_record_tm(r, T, tmT, mode, seed, gidx, st);
}

/* ---- whole-array sweeps -------------------------------------------------- */
static void gather(const orc_record_cols_t* c, uint64_t i, orc_record_t* r) {
  r->minute = c->minute ? c->minute[i] : 0; r->hour = c->hour ? c->hour[i] : 0;
  r->dom = c->dom ? c->dom[i] : 0; r->month = c->month ? c->month[i] : 0;
  r->dow = c->dow ? c->dow[i] : 0;
  r->ras = c->ras ? c->ras[i] : 0; r->flags = c->flags[i];
  r->finished_at = c->finished_at ? c->finished_at[i] : 0;
  r->runs_limit = c->runs_limit ? c->runs_limit[i] : 0;
  r->reset_interval = c->reset_interval ? c->reset_interval[i] : 0;
  r->success = c->success ? c->success[i] : 0; r->failed = c->failed ? c->failed[i] : 0;
  r->remedy_success = c->remedy_success ? c->remedy_success[i] : 0;
  r->remedy_failed = c->remedy_failed ? c->remedy_failed[i] : 0;
  r->remedy_total = c->remedy_total ? c->remedy_total[i] : 0;
  r->remedy_finished_at = c->remedy_finished_at ? c->remedy_finished_at[i] : 0;
  r->reserved = 0;
}
/* write back only what the tick changed (keeps untouched cache lines clean) */
#define PUT(col, val) do { if (c->col && c->col[i] != (val)) c->col[i] = (val); } while (0)
static void scatter(orc_record_cols_t* c, uint64_t i, const orc_record_t* r) {
  PUT(flags, r->flags);
  PUT(finished_at, r->finished_at);
  PUT(success, r->success);
  PUT(failed, r->failed);
  PUT(remedy_success, r->remedy_success);
  PUT(remedy_failed, r->remedy_failed);
  PUT(remedy_total, r->remedy_total);
  PUT(remedy_finished_at, r->remedy_finished_at);
}
#undef PUT
static void count_action(orc_tick_stats_t* st, uint32_t act, uint64_t gidx) {
  st->n_emitted++;
  st->n_submit_hc += (act & ACT_SUBMIT_HC) != 0;
  st->n_run_remedy += (act & ACT_RUN_REMEDY) != 0;
  st->n_stopped += (act & ACT_STOPPED) != 0;
  st->n_parse_error += (act & ACT_PARSE_ERROR) != 0;
  st->n_remedy_skip += (act & ACT_REMEDY_SKIP) != 0;
  st->n_reset_on_pass += (act & ACT_RESET_ON_PASS) != 0;
  st->n_reset_on_interval += (act & ACT_RESET_ON_INTERVAL) != 0;
  st->n_anomaly += (act & ACT_ANOMALY) != 0;
  st->idx_xor ^= gidx;
  st->idx_sum += gidx;
}

/* One record of one tick over the SoA columns.  A record without a posted result, in open
 * loop, can only change `flags` and `finished_at` (B.3 step 2: the "Stopped" report), so its
 * counter / remedy columns are neither read nor compared: the CPU sweep then touches the same
 * 56 B per record the north star counts, not all 92. */
static inline uint32_t sweep_one(orc_record_cols_t* c, uint64_t i, int64_t T, const struct tm* tmT,
                                 uint32_t mode, uint64_t seed, uint64_t gidx, orc_tick_stats_t* st) {
  orc_record_t r;
  const uint32_t f = c->flags[i];
  if (!(mode & MODE_CLOSED_LOOP) && !(f & (F_PENDING_OK | F_PENDING_FAIL | F_REMEDY_PENDING))) {
    memset(&r, 0, sizeof r);
    r.minute = c->minute ? c->minute[i] : 0; r.hour = c->hour ? c->hour[i] : 0;
    r.dom = c->dom ? c->dom[i] : 0; r.month = c->month ? c->month[i] : 0;
    r.dow = c->dow ? c->dow[i] : 0;
    r.ras = c->ras ? c->ras[i] : 0;
    r.flags = f;
    const int64_t fin = c->finished_at ? c->finished_at[i] : 0;
    r.finished_at = fin;
    const uint32_t a = tick_record_tm(&r, T, tmT, mode, seed, gidx, st);
    if (r.flags != f) c->flags[i] = r.flags;
    if (c->finished_at && r.finished_at != fin) c->finished_at[i] = r.finished_at;
    return a;
  }
  gather(c, i, &r);
  const uint32_t a = tick_record_tm(&r, T, tmT, mode, seed, gidx, st);
  scatter(c, i, &r);
  return a;
}

int orc_sweep(orc_record_cols_t* cols, uint64_t n, uint64_t shard_base, int64_t T, uint32_t mode,
              uint64_t seed, uint64_t* due_idx, uint32_t* due_action, uint64_t cap,
              uint64_t* n_out, orc_tick_stats_t* stats) {
  orc_tick_stats_t st;
  memset(&st, 0, sizeof st);
  st.n_records = n;
  struct tm tmT[ORC_MAX_ZONES + 1];
  zone_tms(T, tmT);
  for (uint64_t i = 0; i < n; i++) {
    uint32_t act = sweep_one(cols, i, T, tmT, mode, seed, shard_base + i, &st);
    if (act) {
      if (st.n_emitted < cap) {
        if (due_idx) due_idx[st.n_emitted] = shard_base + i;
        if (due_action) due_action[st.n_emitted] = act;
      }
      count_action(&st, act, shard_base + i);
    }
  }
  if (n_out) *n_out = st.n_emitted;
  if (stats) *stats = st;
  return st.n_emitted > cap ? E_NOSPACE : 0;
}

/* ---- multi-threaded sweep (CPU baseline B2) -------------------
 * Records are independent, so the array is cut into one contiguous chunk per
 * thread.  Two passes per chunk, separated by a barrier: (1) evaluate every
 * record, keep its action byte in a scratch array, count; (2) with the
 * exclusive prefix of the chunk counts known, write the chunk's entries at
 * their final positions of the global ascending list.  No serial tail.  The
 * worker threads are created once and parked between calls. */
typedef struct {
  uint64_t n;          /* entries emitted by this chunk */
  orc_tick_stats_t st;
  char pad[128];       /* neighbouring slots are written by different threads */
} mt_slot_t;

typedef struct {
  orc_record_cols_t* cols;
  uint64_t n, shard_base;
  int64_t T;
  uint32_t mode;
  uint64_t seed;
  uint64_t* due_idx;
  uint32_t* due_action;
  uint64_t cap;
  uint8_t* act8; /* scratch: the action byte of every record of this tick */
  mt_slot_t* slot;
  int nthreads;
  const struct tm* tms; /* T in every registered zone */
  pthread_barrier_t bar;
} mt_ctx_t;

static void mt_run_chunk(mt_ctx_t* x, int t) {
  const uint64_t chunk = (x->n + (uint64_t)x->nthreads - 1) / (uint64_t)x->nthreads;
  const uint64_t lo = (uint64_t)t * chunk < x->n ? (uint64_t)t * chunk : x->n;
  const uint64_t hi = lo + chunk < x->n ? lo + chunk : x->n;
  /* thread-private state on this thread's stack */
  orc_tick_stats_t st;
  memset(&st, 0, sizeof st);
  orc_record_cols_t cols = *x->cols;
  const uint64_t base = x->shard_base, seed = x->seed;
  const int64_t T = x->T;
  const uint32_t mode = x->mode;
  uint8_t* act8 = x->act8;
  const struct tm* tmT = x->tms; /* computed once per tick by the caller: setenv/tzset are process-global */
  uint64_t n = 0;
  for (uint64_t i = lo; i < hi; i++) {
    uint32_t a = sweep_one(&cols, i, T, tmT, mode, seed, base + i, &st);
    act8[i] = (uint8_t)a; /* every action bit is below 0x100 */
    if (a) {
      n++;
      count_action(&st, a, base + i);
    }
  }
  x->slot[t].n = n;
  x->slot[t].st = st;
  pthread_barrier_wait(&x->bar);
  uint64_t pos = 0;
  for (int u = 0; u < t; u++) pos += x->slot[u].n;
  uint64_t* di = x->due_idx;
  uint32_t* da = x->due_action;
  const uint64_t cap = x->cap;
  if ((!di && !da) || pos >= cap) return;
  for (uint64_t i = lo; i < hi; i++) {
    const uint8_t a = act8[i];
    if (!a) continue;
    if (pos >= cap) break;
    if (di) di[pos] = base + i;
    if (da) da[pos] = a;
    pos++;
  }
}

/* parked workers: index k runs chunk k+1 of every job that uses more than k+1 threads */
static struct {
  pthread_mutex_t call;  /* one MT sweep at a time */
  pthread_mutex_t mu;
  pthread_cond_t start, done;
  pthread_t* th;
  int n_workers;
  uint64_t gen;
  int remaining;
  mt_ctx_t* ctx;
} g_pool = {PTHREAD_MUTEX_INITIALIZER, PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER,
            PTHREAD_COND_INITIALIZER, NULL, 0, 0, 0, NULL};

typedef struct { int k; uint64_t seen; } mt_worker_arg_t;

/* Pin a parked worker to the (k+1)-th CPU of the affinity mask it inherited (the caller's
 * thread keeps chunk 0 wherever the scheduler puts it): one chunk per core, no migration
 * between the two passes — the CPU baseline varied 5.8x between two boxes without it. */
static void mt_pin_worker(int k) {
  cpu_set_t have, want;
  if (sched_getaffinity(0, sizeof have, &have) != 0) return;
  const int ncpu = CPU_COUNT(&have);
  if (ncpu <= 1) return;
  int target = (k + 1) % ncpu, seen = 0;
  for (int c = 0; c < CPU_SETSIZE; c++) {
    if (!CPU_ISSET(c, &have)) continue;
    if (seen++ == target) {
      CPU_ZERO(&want);
      CPU_SET(c, &want);
      (void)pthread_setaffinity_np(pthread_self(), sizeof want, &want);
      return;
    }
  }
}

static void* mt_pool_worker(void* p) {
  mt_worker_arg_t arg = *(mt_worker_arg_t*)p;
  free(p);
  mt_pin_worker(arg.k);
  uint64_t seen = arg.seen;
  for (;;) {
    pthread_mutex_lock(&g_pool.mu);
    while (g_pool.gen == seen) pthread_cond_wait(&g_pool.start, &g_pool.mu);
    seen = g_pool.gen;
    mt_ctx_t* x = g_pool.ctx;
    pthread_mutex_unlock(&g_pool.mu);
    const int mine = arg.k + 1 < x->nthreads;
    if (mine) mt_run_chunk(x, arg.k + 1);
    pthread_mutex_lock(&g_pool.mu);
    if (--g_pool.remaining == 0) pthread_cond_signal(&g_pool.done);
    pthread_mutex_unlock(&g_pool.mu);
  }
  return NULL;
}

/* make sure `want` workers exist; returns how many do (called with g_pool.call held, no job posted) */
static int mt_pool_grow(int want) {
  if (want <= g_pool.n_workers) return g_pool.n_workers;
  pthread_t* nt = (pthread_t*)realloc(g_pool.th, (size_t)want * sizeof *nt);
  if (!nt) return g_pool.n_workers;
  g_pool.th = nt;
  while (g_pool.n_workers < want) {
    mt_worker_arg_t* a = (mt_worker_arg_t*)malloc(sizeof *a);
    if (!a) break;
    a->k = g_pool.n_workers;
    a->seen = g_pool.gen;
    if (pthread_create(&g_pool.th[g_pool.n_workers], NULL, mt_pool_worker, a) != 0) { free(a); break; }
    pthread_detach(g_pool.th[g_pool.n_workers]);
    g_pool.n_workers++;
  }
  return g_pool.n_workers;
}

int orc_sweep_mt(orc_record_cols_t* cols, uint64_t n, uint64_t shard_base, int64_t T,
                 uint32_t mode, uint64_t seed, uint64_t* due_idx, uint32_t* due_action,
                 uint64_t cap, uint64_t* n_out, orc_tick_stats_t* stats, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 1024) nthreads = 1024;
  if (n < (uint64_t)nthreads) nthreads = n ? (int)n : 1;
  pthread_mutex_lock(&g_pool.call);
  const int have = 1 + mt_pool_grow(nthreads - 1); /* the caller's thread + parked workers */
  if (have < nthreads) nthreads = have;
  mt_ctx_t x;
  memset(&x, 0, sizeof x);
  x.cols = cols; x.n = n; x.shard_base = shard_base; x.T = T; x.mode = mode; x.seed = seed;
  x.due_idx = due_idx; x.due_action = due_action; x.cap = cap; x.nthreads = nthreads;
  struct tm tms[ORC_MAX_ZONES + 1];
  zone_tms(T, tms);
  x.tms = tms;
  x.act8 = (uint8_t*)malloc(n ? n : 1);
  x.slot = (mt_slot_t*)calloc((size_t)nthreads, sizeof *x.slot);
  if (!x.act8 || !x.slot || pthread_barrier_init(&x.bar, NULL, (unsigned)nthreads) != 0) {
    free(x.act8); free(x.slot);
    pthread_mutex_unlock(&g_pool.call);
    return -5;
  }
  /* post the job: every parked worker wakes up, those beyond nthreads only acknowledge */
  pthread_mutex_lock(&g_pool.mu);
  g_pool.ctx = &x;
  g_pool.remaining = g_pool.n_workers;
  g_pool.gen++;
  pthread_cond_broadcast(&g_pool.start);
  pthread_mutex_unlock(&g_pool.mu);
  mt_run_chunk(&x, 0); /* chunk 0 on the caller's thread */
  pthread_mutex_lock(&g_pool.mu);
  while (g_pool.remaining != 0) pthread_cond_wait(&g_pool.done, &g_pool.mu);
  g_pool.ctx = NULL;
  pthread_mutex_unlock(&g_pool.mu);
  pthread_barrier_destroy(&x.bar);

  orc_tick_stats_t st;
  memset(&st, 0, sizeof st);
  uint64_t pos = 0;
  for (int t = 0; t < nthreads; t++) {
    pos += x.slot[t].n;
    const uint64_t* a = (const uint64_t*)&x.slot[t].st;
    uint64_t* d = (uint64_t*)&st;
    for (size_t q = 1; q < sizeof st / sizeof(uint64_t); q++) {
      if (q == 14) d[q] ^= a[q]; /* idx_xor */
      else d[q] += a[q];
    }
  }
  st.n_records = n;
  free(x.act8);
  free(x.slot);
  pthread_mutex_unlock(&g_pool.call);
  if (n_out) *n_out = pos;
  if (stats) *stats = st;
  return pos > cap ? E_NOSPACE : 0;
}

/* CPU baseline B1: what the Go controller pays per evaluation — re-parse the
 * cron string and walk Next() on every pass (hcc.go:253-264). */
uint64_t orc_faithful_eval(const orc_healthcheck_t* hcs, uint64_t n, int64_t T) {
  uint64_t submits = 0;
  for (uint64_t i = 0; i < n; i++) {
    const orc_healthcheck_t* hc = &hcs[i];
    if (!hc->has_resource) continue;                         /* :227 */
    int64_t finished = hc->finished_at_set ? hc->finished_at : 0; /* :231-235 */
    int64_t ras = hc->repeat_after_sec;
    if (ras <= 0 && hc->cron_len == 0) continue;             /* :238 Stopped */
    if (ras <= 0 && hc->cron_len != 0) {                     /* :251 */
      orc_cron_t c;
      if (orc_cron_parse(hc->cron, hc->cron_len, &c, NULL, 0) != 0) continue; /* :254 */
      ras = orc_cron_repeat_after_sec(&c, T);                /* :262 */
      (void)ras;
    } else if ((T - finished) < ras) {                       /* :264 (timer armed) */
      continue;
    }
    submits++;                                               /* :269-288 */
  }
  return submits;
}
