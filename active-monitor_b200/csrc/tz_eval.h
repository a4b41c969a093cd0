// tz_eval.h — UTC offset of a time zone at an instant, from a flattened view of its TZif data
// (transition table, then the POSIX rule of the footer).  Host/device-neutral (AM_HD); used by the
// registry (tz.cpp), which hands the sweep one offset per zone and the window in which they hold.
#pragma once
#include <stdint.h>

#include "civil.h"

namespace amsweep_tz {

struct Rule {      // one side of a POSIX TZ daylight rule
  int32_t kind;    // 0: Jn (1..365, Feb 29 never counted)  1: n (0..365)  2: Mm.w.d   -1: none
  int32_t a, b, c;
  int32_t time;    // seconds after local midnight (default 02:00:00; may be negative or > 24 h)
};

struct ZoneDesc {          // fixed-size part of a zone
  uint32_t trans_begin;    // first entry of the zone in the flattened transition arrays
  uint32_t trans_count;
  int32_t off_first;       // before the first transition
  int32_t has_footer, has_dst;
  int32_t std_off, dst_off;  // seconds EAST of UTC
  Rule start, end;
};

AM_HD bool tz_is_leap(int64_t y) { return (y % 4 == 0 && y % 100 != 0) || y % 400 == 0; }

// seconds since the epoch of a rule's instant in year y, on the rule's own local clock
AM_HD int64_t tz_rule_local_seconds(const Rule& r, int64_t y) {
  int64_t day;
  if (r.kind == 0) {
    int32_t d = r.a;
    if (tz_is_leap(y) && d >= 60) d += 1;
    day = amsweep::days_from_civil(y, 1, 1) + d - 1;
  } else if (r.kind == 1) {
    day = amsweep::days_from_civil(y, 1, 1) + r.a;
  } else {
    const int64_t first = amsweep::days_from_civil(y, r.a, 1);
    const int32_t wd = amsweep::weekday_from_days(first);
    int32_t d = 1 + (r.c - wd + 7) % 7 + 7 * (r.b - 1);
    const int32_t len = (r.a == 2) ? (tz_is_leap(y) ? 29 : 28) : ((r.a == 4 || r.a == 6 || r.a == 9 || r.a == 11) ? 30 : 31);
    while (d > len) d -= 7;  // week 5 = the last one
    day = first + d - 1;
  }
  return day * 86400 + r.time;
}

AM_HD int32_t tz_footer_offset(const ZoneDesc& z, int64_t utc) {
  if (!z.has_dst) return z.std_off;
  int64_t days, y;
  int32_t sod, m, d;
  amsweep::split_days(utc + z.std_off, days, sod);
  amsweep::civil_from_days(days, y, m, d);
  // the rule instants of year y as UTC: the start is given in standard time, the end in daylight time
  const int64_t s = tz_rule_local_seconds(z.start, y) - z.std_off;
  const int64_t e = tz_rule_local_seconds(z.end, y) - z.dst_off;
  const bool dst = s < e ? (utc >= s && utc < e) : !(utc >= e && utc < s);
  return dst ? z.dst_off : z.std_off;
}

AM_HD int32_t tz_zone_offset(const ZoneDesc& z, const int64_t* trans, const int32_t* off, int64_t utc) {
  const int64_t* t = trans + z.trans_begin;
  const int32_t* o = off + z.trans_begin;
  if (z.trans_count == 0) return z.has_footer ? tz_footer_offset(z, utc) : z.off_first;
  if (utc < t[0]) return z.off_first;
  if (utc >= t[z.trans_count - 1] && z.has_footer) return tz_footer_offset(z, utc);
  uint32_t lo = 0, hi = z.trans_count;  // last transition <= utc
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) / 2;
    if (t[mid] <= utc) lo = mid; else hi = mid;
  }
  return o[lo];
}

}  // namespace amsweep_tz
